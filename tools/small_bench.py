#!/usr/bin/env python3
"""Timing of the small-system paths on the GPU box: Galperin (2 balls), pool (16 balls), and the multi-ball ensemble
(S pool tables).  Prints one JSON line per measurement; kernel time = CUDA events inside the library."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402


def single(cfg, end_time, reps=3):
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    best = None
    for _ in range(reps):
        bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
        bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
        bld._flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c = bld.evolve(end_time)
        wall = time.perf_counter() - t0
        ms = h.last_kernel_ms()
        if best is None or ms < best[1]:
            best = (c, ms, wall)
    c, ms, wall = best
    print(json.dumps({"workload": cfg.name, "counts": c, "kernel_ms": ms, "wall_ms": wall * 1e3,
                      "collisions_per_s": sum(c) / (ms * 1e-3)}), flush=True)


def pool_ensemble(S, reps=3):
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg = workloads.pool()
    pos = np.ascontiguousarray(np.broadcast_to(cfg.pos[None], (S,) + cfg.pos.shape))
    vel = np.ascontiguousarray(np.broadcast_to(cfg.vel[None], (S,) + cfg.vel.shape))
    rng = np.random.default_rng(0)
    vel[:, 15, 0] *= rng.uniform(0.8, 1.2, S)
    vel[:, 15, 1] = vel[:, 15, 0] * rng.uniform(-0.02, 0.02, S)
    obs = workloads.make_obstacles(billiards, cfg.obstacles)
    best = None
    for _ in range(reps):
        ens = billiards.BilliardEnsemble(obs, pos, vel, radius=cfg.radius, mass=cfg.mass, dot_fma=fma)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ens.run(-1, end_time=100.0)
        wall = time.perf_counter() - t0
        ms = h.last_kernel_ms()
        total = int(ens.stats()[0].item())
        if best is None or ms < best[1]:
            best = (total, ms, wall)
    total, ms, wall = best
    print(json.dumps({"workload": f"pool_ensemble_{S}", "collisions": total, "kernel_ms": ms, "wall_ms": wall * 1e3,
                      "collisions_per_s": total / (ms * 1e-3)}), flush=True)


def pool_with_callbacks(reps=5):
    """f1: the whole pool run (415 collisions) with a callback on every ball, every wall and the clock"""
    fma = _lib.probe_dot_fma()
    cfg = workloads.pool()
    best = None
    for mode in ("replay", "step"):
        for _ in range(reps if mode == "replay" else 1):
            bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
            bld.callback_mode = mode
            bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
            bld._flush()
            calls = [0]

            def cb(*a):
                calls[0] += 1
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c = bld.evolve(100.0, time_callback=cb, ball_callbacks={i: cb for i in range(16)},
                           obstacle_callbacks={o: cb for o in bld.obstacles})
            wall = time.perf_counter() - t0
            if best is None or wall < best:
                best = wall
        print(json.dumps({"workload": f"pool_callbacks_{mode}", "counts": c, "callbacks": calls[0], "wall_ms": best * 1e3}),
              flush=True)
        best = None


def frame_loop(frames=600):
    """what visualize_matplotlib.animate / the pyglet window do: many short evolve calls + a read of the positions"""
    fma = _lib.probe_dot_fma()
    cfg = workloads.ideal_gas()
    bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
    bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    bld.evolve(0.05)
    _ = bld.balls_position
    torch.cuda.synchronize()
    dt = 1.0 / 60.0
    t0 = time.perf_counter()
    n = 0
    for k in range(frames):
        c = bld.evolve(0.05 + (k + 1) * dt)
        pos = bld.balls_position
        n += sum(c)
    wall = time.perf_counter() - t0
    print(json.dumps({"workload": "frame_loop_ideal_gas_1024", "frames": frames, "collisions": n,
                      "frames_per_s": frames / wall, "ms_per_frame": wall / frames * 1e3}), flush=True)


if __name__ == "__main__":
    pool_with_callbacks()
    frame_loop()
    single(workloads.galperin(), 16.0)
    single(workloads.pool(), 100.0)
    single(workloads.newtons_cradle(5, True), 50.0)
    for S in (1 << 10, 1 << 14, 1 << 16, 1 << 18):
        pool_ensemble(S)
