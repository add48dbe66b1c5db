#!/usr/bin/env python3
"""One Galperin run (314159 collisions, one launch of the one-thread kernel) -- the target of an ncu capture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402

cfg = workloads.galperin() if len(sys.argv) < 2 or sys.argv[1] == "galperin" else workloads.pool()
bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles))
bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
print(bld.evolve(cfg.end_time), _lib.get_handle().last_kernel_ms(), "ms")
