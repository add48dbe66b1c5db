#!/usr/bin/env python3
"""Where a frame of the animate() pattern spends its host time: cProfile of evolve(t + dt) + balls_position / balls_velocity
reads for one of the examples (python tools/frame_profile.py pool_first_shot)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200"), os.path.join(ROOT, "examples")):
    sys.path.insert(0, p)
import headless  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pool_first_shot"
headless.run(name)  # warm
build, end_time, dt = headless.EXAMPLES[name]
bld = build()
n_frames = int(round(end_time / dt))
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for f in range(1, n_frames + 1):
    bld.evolve(f * dt)
    pos, vel = bld.balls_position, bld.balls_velocity
pr.disable()
wall = time.perf_counter() - t0
print(f"{name}: {n_frames} frames in {wall * 1e3:.1f} ms under the profiler ({wall / n_frames * 1e6:.1f} us per frame)")
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
