#!/usr/bin/env python3
"""Galperin's billiard (README.md:58-63, 114-135 of the reference: the number of collisions spells pi) at six mass
ratios in ONE launch: a multi-ball ensemble whose systems have their own masses (``mass`` of shape (S, n)).

    python examples/pi_digits_ensemble.py      ->  3, 31, 314, 3141, 31415, 314159 collisions
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import billiards  # noqa: E402

digits = 6
wall = [billiards.InfiniteWall((0, -1), (0, 1), exterior="right")]
pos = np.repeat(np.asarray([[(3.0, 0.0), (6.0, 0.0)]]), digits, axis=0)     # light ball at rest, heavy ball behind it
vel = np.repeat(np.asarray([[(0.0, 0.0), (-1.0, 0.0)]]), digits, axis=0)    # ... moving towards the wall
mass = np.asarray([[1.0, 100.0 ** k] for k in range(digits)])
radius = np.repeat(np.asarray([[0.2, 1.0]]), digits, axis=0)
ens = billiards.BilliardEnsemble(wall, pos, vel, radius=radius, mass=mass)
t0 = time.perf_counter()
bb, bo = ens.evolve(float("inf"), max_collisions=10 ** 7)
dt = time.perf_counter() - t0
for k in range(digits):
    print(f"mass ratio 100^{k}: {int(bb[k] + bo[k])} collisions ({int(bb[k])} ball-ball, {int(bo[k])} ball-wall)")
print(f"one launch, {dt * 1e3:.1f} ms")
