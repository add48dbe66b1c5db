#!/usr/bin/env python3
"""Small fixed workload for ncu captures: python tools/ncu_target.py gas <side> <events> | sinai <systems> <collisions>
| galperin.  Launch order for `gas`: table build (2 kernels), warm-up bl_evolve_kernel, measured bl_evolve_kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402

fma = _lib.probe_dot_fma()
what = sys.argv[1] if len(sys.argv) > 1 else "gas"
if what == "gas":
    side, events = int(sys.argv[2]), int(sys.argv[3])
    cfg = workloads.dense_gas(side)
    bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
    bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    print("warm-up", bld.evolve_events(events))
    print("measured", bld.evolve_events(events), "kernel ms", _lib.get_handle(0, fma).last_kernel_ms())
elif what == "pool_ensemble":
    import numpy as np
    S = int(sys.argv[2])
    cfg = workloads.pool()
    rng = np.random.default_rng(2000)
    pos = np.ascontiguousarray(np.broadcast_to(cfg.pos[None], (S,) + cfg.pos.shape))
    vel = np.ascontiguousarray(np.broadcast_to(cfg.vel[None], (S,) + cfg.vel.shape))
    speed = cfg.vel[15, 0] * rng.uniform(0.8, 1.2, S)
    ang = rng.uniform(-0.02, 0.02, S)
    vel[:, 15, 0] = speed * np.cos(ang)
    vel[:, 15, 1] = speed * np.sin(ang)
    obs = workloads.make_obstacles(billiards, cfg.obstacles)
    for _ in range(2):
        ens = billiards.BilliardEnsemble(obs, pos, vel, radius=cfg.radius, mass=cfg.mass, dot_fma=fma)
        ens.run(-1, end_time=100.0)
    print("pool ensemble", int(ens.counts.sum()), "collisions, kernel ms", _lib.get_handle(0, fma).last_kernel_ms())
elif what == "galperin":
    b = workloads.galperin().build(billiards, dot_fma=fma)
    print(b.evolve(16))
else:
    n_sys, n_coll = int(sys.argv[2]), int(sys.argv[3])
    obstacles, pos, vel = workloads.sinai(n_sys, seed=1)
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, dot_fma=fma)
    ens.run(n_coll)
    ens.run(n_coll)
    print("sinai", int(ens.counts.sum()), "kernel ms", _lib.get_handle(0, fma).last_kernel_ms())
torch.cuda.synchronize()
