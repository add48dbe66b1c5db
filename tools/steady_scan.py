#!/usr/bin/env python3
"""Steady-state throughput of bl_evolve_kernel at N = 16384 (the bench's regime: millions of collisions into the run):
1e6 warm-up collisions, then 2 x 1e6 timed.  Knobs from the environment as in tools/tune_scan.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import billiards, workloads
from billiards import _lib
fma = _lib.probe_dot_fma()
h = _lib.get_handle(0, fma)
if os.environ.get("BL_TARGET"):
    h.check(h.lib.bl_set_batch_target(h.h, int(os.environ["BL_TARGET"])))
side = int(os.environ.get("BL_SIDE", "128"))
cfg = workloads.dense_gas(side)
b = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
b.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
b.evolve_events(1000000)
ms = 0.0; rounds = 0; redo = 0
for _ in range(2):
    b.evolve_events(1000000); ms += h.last_kernel_ms()
    p = [int(x) for x in b.last_result.prof]; rounds += p[5]; redo += p[7]
print(f"winctl {os.environ.get('BL_WINCTL', 'default'):24s} horizon {os.environ.get('BL_REPAIR_HORIZON', 'default'):8s} target {os.environ.get('BL_TARGET', 'auto'):5s} "
      f"N={cfg.n}: {2e6 / ms * 1e3:,.0f}/s ({2e6 / rounds:.1f} ev/round, {100.0 * redo / rounds:.1f} % redone)", flush=True)
