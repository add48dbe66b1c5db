#!/usr/bin/env python3
"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402

fma = _lib.probe_dot_fma()
what = sys.argv[1:] or ["pair", "tiny", "lanes", "multi", "sinai", "gas"]
if "pair" in what:
    g = workloads.galperin()
    b = billiards.Billiard(workloads.make_obstacles(billiards, g.obstacles), dot_fma=fma)
    b.add_balls(g.pos, g.vel, g.radius, g.mass)
    print("pair", b.evolve_events(3000, log_capacity=3000, payload=True)[0], b.next_collision)
if "tiny" in what:
    c = workloads.newtons_cradle(4, True)
    b = billiards.Billiard(workloads.make_obstacles(billiards, c.obstacles), dot_fma=fma)
    b.add_balls(c.pos, c.vel, c.radius, c.mass)
    print("tiny", b.evolve(30.0), b.next_collision[0])
if "lanes" in what:
    p = workloads.pool()
    b = billiards.Billiard(workloads.make_obstacles(billiards, p.obstacles), dot_fma=fma)
    b.add_balls(p.pos, p.vel, p.radius, p.mass)
    print("lanes", b.evolve(100.0), len(b.toi_table))
if "multi" in what:
    p = workloads.pool()
    S = 70
    pos = np.repeat(p.pos[None], S, axis=0)
    vel = np.repeat(p.vel[None], S, axis=0)
    vel[:, 15, 0] *= np.linspace(0.8, 1.2, S)
    e = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, p.obstacles), pos, vel, radius=p.radius, mass=p.mass,
                                   dot_fma=fma)
    bb, bo = e.evolve(40.0)
    print("multi", int(bb.sum()), int(bo.sum()))
    for n in (3, 9, 32):
        rng = np.random.default_rng(n)
        side = int(np.ceil(np.sqrt(n)))
        cells = np.asarray([(x + 0.5, y + 0.5) for x in range(side) for y in range(side)])[:n]
        pos = np.repeat(cells[None], 20, axis=0) + rng.uniform(-0.1, 0.1, (20, n, 2))
        vel = rng.uniform(-1, 1, (20, n, 2))
        e = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, workloads.box_walls(0.0, 0.0, side, side)), pos, vel,
                                       radius=0.2, mass=1.0, dot_fma=fma)
        bb, bo = e.evolve(3.0)
        print("multi", n, int(bb.sum()), int(bo.sum()))
if "sinai" in what:
    obstacles, pos, vel = workloads.sinai(4096, seed=3)
    for mode in (0, 2, 1):
        h = _lib.get_handle(None, fma)
        h.check(h.lib.bl_set_ensemble_generic(h.h, mode))
        e = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, dot_fma=fma, count_hits=True)
        e.run(60)
        print("sinai", mode, int(e.counts.sum()))
    h.check(h.lib.bl_set_ensemble_generic(h.h, 0))
if "gas" in what:
    gcfg = workloads.dense_gas(20, seed=1)
    b = billiards.Billiard(workloads.make_obstacles(billiards, gcfg.obstacles), dot_fma=fma)
    b.add_balls(gcfg.pos, gcfg.vel, gcfg.radius, gcfg.mass)
    print("gas", b.evolve_events(600), b.balls_position.shape, len(b._balls_toi))
