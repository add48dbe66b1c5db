#!/usr/bin/env python3
"""Calls per second of the next-collision query (bl_next: one launch, the result block to the host, one wait)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import billiards  # noqa: E402
import workloads  # noqa: E402

for cfg in (workloads.pool(), workloads.ideal_gas(32)):
    bld = cfg.build(billiards)
    bld.evolve(1.0)
    for _ in range(50):
        bld._next = None
        bld.next_collision
    t0 = time.perf_counter()
    for _ in range(2000):
        bld._next = None  # drop the cached answer: every call is a launch
        bld.next_collision
    dt = time.perf_counter() - t0
    print(f"{cfg.name}: {2000 / dt:,.0f} next_collision queries/s ({dt / 2000 * 1e6:.1f} us each)")
