#!/usr/bin/env python3
"""Throughput of bl_evolve_kernel on long launches (3 x 100000 collisions after a warm-up) for N = 1024, 4096, 16384 --
the tool behind profiles/r1_target_scan.log.  Knobs come from the environment:
    BL_TARGET=<keys per round>   BL_WINCTL=<violation factor,growth,below fraction,shrink,above fraction>
    BL_REPAIR_HORIZON=<windows>
Run on the GPU box: for t in 28 36 48; do BL_TARGET=$t python tools/tune_scan.py; done"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import torch, billiards, workloads
from billiards import _lib
fma = _lib.probe_dot_fma()
h = _lib.get_handle(0, fma)
if os.environ.get("BL_TARGET"):
    h.check(h.lib.bl_set_batch_target(h.h, int(os.environ["BL_TARGET"])))
out = []
for side in (32, 64, 128):
    cfg = workloads.dense_gas(side)
    b = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
    b.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    b.evolve_events(100000)
    ms = 0.0; rounds = 0; redo = 0
    for _ in range(3):
        b.evolve_events(100000); ms += h.last_kernel_ms()
        p = [int(x) for x in b.last_result.prof]; rounds += p[5]; redo += p[7]
    out.append(f"N={cfg.n}: {300000/ms*1e3:,.0f}/s ({300000/rounds:.1f} ev/round, {redo} redo)")
print(os.environ.get("BL_WINCTL", "default"), "target", os.environ.get("BL_TARGET", "auto"), " | ".join(out), flush=True)
