#!/usr/bin/env python3
"""One launch of the one-CTA kernel (ncu target): dense gas of side*side balls, `events` collisions."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 12
events = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
fma = _lib.probe_dot_fma()
h = _lib.get_handle(0, fma)
cfg = workloads.dense_gas(side)
bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
c = bld.evolve_events(events)
print(cfg.n, c, h.last_kernel_ms(), "ms", "rows re-scanned per collision:", int(bld.last_result.prof[6]) / max(sum(c), 1),
      "cycles per collision (argmin, resolve, pairs, obstacles, barrier, rescan):", [int(x) // max(sum(c), 1) for x in bld.last_result.prof[8:14]])
