#!/usr/bin/env python3
"""The seven example set-ups of the reference (examples/*.py there), run headless through this package.

The reference's examples hand their billiard to ``visualize.animate(bld, end_time, dt=1/30)``, which calls
``bld.evolve(t)`` once per frame and reads ``bld.balls_position`` / ``balls_velocity`` for the artists.  Plotting is out of
scope here (no matplotlib in the image), so each example below is advanced in exactly that access pattern and prints
one JSON line: balls, collisions, frames, frames per second, collisions per second.  Set-ups come from ``workloads.py``
(the generators the parity fixtures use: same random streams as the reference's scripts where those seed numpy).

    python examples/headless.py                 # all seven
    python examples/headless.py pool collapse   # a selection
    python examples/headless.py --dt 0          # one evolve(end_time) call per example instead of frames
"""
import argparse
import json
import os
import sys
import time
from math import cos, pi, sin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402


def from_config(cfg):
    bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles))
    for k in range(cfg.n):  # one add_ball per ball, like the scripts do
        bld.add_ball(tuple(cfg.pos[k]), tuple(cfg.vel[k]), float(cfg.radius[k]), float(cfg.mass[k]))
    return bld


def ideal_gas():
    """examples/ideal_gas.py:18-41 of the reference: 1000 small balls in the square, equal speeds."""
    rs = np.random.RandomState(0)
    bld = billiards.Billiard(workloads.make_obstacles(billiards, workloads.box_walls(-1, -1, 1, 1)))
    for _ in range(1000):
        pos = rs.uniform((-0.99, -0.99), (0.99, 0.99))
        angle = rs.uniform(0, 2 * pi)
        bld.add_ball(pos, np.asarray([cos(angle), sin(angle)]) / 5, radius=0.01)
    return bld


def sinai_billiard():
    """examples/sinai_billiard.py:17-41: 300 point particles in ONE billiard, slowed down in place afterwards."""
    rs = np.random.RandomState(0)
    bld = billiards.Billiard(workloads.make_obstacles(billiards, workloads.SINAI_OBSTACLES))
    for _ in range(300):
        pos = rs.uniform((-1, -1), (1, 1))
        angle = rs.uniform(0, 2 * pi)
        bld.add_ball(pos, [cos(angle), sin(angle)], radius=0)
    bld.balls_velocity /= 5
    bld.recompute_toi()
    return bld


def newtons_cradle():
    """examples/newtons_cradle.py:14-27 (the striker moves at speed 3 there; the test fixture uses 1)."""
    cfg = workloads.newtons_cradle(5, True)
    cfg.obstacles[1] = ("wall", (12, -2), (12, 2), "left")
    cfg.vel[0] = (3.0, 0.0)
    return from_config(cfg)


EXAMPLES = {
    # name: (builder, end_time of the script's animate() call, dt of that call)
    "collapse": (lambda: from_config(workloads.collapse()), 15.0, 1 / 30),
    "ideal_gas": (ideal_gas, 10.0, 1 / 30),
    "maze": (lambda: from_config(workloads.maze()), 10.0, 1 / 30),
    "newtons_cradle": (newtons_cradle, 12.0, 1 / 30),
    "pi_with_pool": (lambda: from_config(workloads.galperin()), 16.0, 1 / 60),
    "pool_first_shot": (lambda: from_config(workloads.pool()), 30.0, 1 / 30),
    "sinai_billiard": (sinai_billiard, 10.0, 1 / 30),
}


def run(name, dt_override=None):
    build, end_time, dt = EXAMPLES[name]
    if dt_override is not None:
        dt = dt_override
    t0 = time.perf_counter()
    bld = build()
    built = time.perf_counter() - t0
    bb = bo = frames = 0
    t0 = time.perf_counter()
    if dt > 0:
        start = bld.time
        n_frames = int(round((end_time - start) / dt))
        for f in range(1, n_frames + 1):
            c = bld.evolve(start + f * dt if f < n_frames else end_time)
            bb += c[0]
            bo += c[1]
            pos, vel = bld.balls_position, bld.balls_velocity  # what the artists read every frame
            frames += 1
    else:
        bb, bo = bld.evolve(end_time)
        pos, vel = bld.balls_position, bld.balls_velocity
        frames = 1
    ran = time.perf_counter() - t0
    speeds = np.linalg.norm(np.asarray(vel), axis=1)
    return {"example": name, "balls": bld.count, "end_time": bld.time, "ball_ball": int(bb), "ball_obstacle": int(bo),
            "frames": frames, "build_s": round(built, 4), "run_s": round(ran, 4),
            "frames_per_s": round(frames / ran, 1), "collisions_per_s": round((bb + bo) / ran, 1),
            "mean_speed": float(speeds.mean()), "kinetic_energy": float((np.asarray(bld.balls_mass)[np.isfinite(bld.balls_mass)]
                                                                         * speeds[np.isfinite(bld.balls_mass)] ** 2).sum() / 2),
            "position_checksum": float(np.asarray(pos).sum())}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("names", nargs="*", default=[], help="examples to run (default: all)")
    ap.add_argument("--dt", type=float, default=None, help="frame step; 0 = a single evolve(end_time)")
    args = ap.parse_args()
    # the first call of each entry point loads its kernels and allocates the pinned staging buffers: keep that out of
    # the first example's numbers
    warm = from_config(workloads.newtons_cradle(3, True))
    warm.evolve(1.0)
    _ = warm.balls_position, warm.balls_velocity
    warm = from_config(workloads.collapse(70))
    warm.evolve(12.0)
    _ = warm.balls_position, warm.balls_velocity
    for name in args.names or list(EXAMPLES):
        print(json.dumps(run(name, args.dt)), flush=True)


if __name__ == "__main__":
    main()
