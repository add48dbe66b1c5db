#!/usr/bin/env python3
"""Randomised parity sweep of the small-system kernels and the ensemble front ends (run on the GPU box):
  * multi-ball ensembles of 1 ... 224 balls (one thread per system, lane groups, one CTA per system from 33 balls, and the lane groups forced for <= 4
    balls), random radii / masses / obstacle sets (walls, a disk, sometimes a line segment), stop / resume in random
    places, against the CPU oracle run per system: clocks, counts, state and the whole time-of-impact table;
  * one-ball ensembles on the randomised geometries of tests/test_gpu_ensembles.py with fresh seeds.
    python tools/fuzz_small.py [n_cases] [seed]
prints one line per case and exits non-zero on the first mismatch."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402
from oracle import oracle  # noqa: E402

spec = importlib.util.spec_from_file_location("tge", os.path.join(ROOT, "tests", "test_gpu_ensembles.py"))
tge = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tge)
INF = float("inf")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same(a, b):
    return np.array_equal(bits(a), bits(b))


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    for case in range(n_cases):
        n = int(rng.choice([1, 2, 2, 3, 4, 4, 5, 7, 8, 12, 16, 16, 20, 27, 32, 33, 41, 64, 65, 97, 130, 224]))
        kernel = 2 if (n <= 4 and rng.random() < 0.4) else 1
        S = int(rng.integers(5, 80)) if n <= 32 else int(rng.integers(2, 12))  # (from 33 balls: one CTA per system)
        obstacles, pos, vel, radius, mass = tge._small_gas(rng, n, S)
        box = float(obstacles[1][1][0])
        kind = rng.random()
        if kind < 0.3:
            obstacles = obstacles[:4]                       # walls only
        elif kind < 0.5:
            c = box / 2  # a line segment instead of the disk, inside the ball-free zone around the centre
            obstacles = obstacles[:4] + [("segment", (c - 0.3, c - 0.25), (c + 0.3, c + 0.2))]
            if n >= 4:
                # no point particle here: on a LineSegment the reference's t_eps shift makes a radius-0 ball collide with
                # the line again and again at one and the same time (it never returns from evolve; neither would the kernel)
                radius = radius.copy()
                radius[3] = radius[0]
        per_system = rng.random() < 0.35   # every system its own radii and masses (bl_multi.per_system)
        if per_system:
            radii = np.repeat(np.asarray(radius)[None], S, axis=0)
            masses = np.repeat(np.asarray(mass)[None], S, axis=0)
            for s in range(1, S):
                radii[s] = np.round(radii[s] * rng.uniform(0.5, 1.1, n) * 1024.0) / 1024.0
                masses[s] = masses[s] * rng.uniform(0.3, 3.0, n)
            if n >= 4 and len(obstacles) == 5 and obstacles[4][0] != "segment":
                radii[:, 3] = 0.0
            elif n >= 4:
                radii[:, 3] = np.maximum(radii[:, 3], 1.0 / 64)  # (no point particle with a segment, see above)
        else:
            radii, masses = radius, mass
        T = float(rng.uniform(2.0, 9.0))
        cuts = sorted(rng.uniform(0.0, T, int(rng.integers(0, 3))).tolist()) + [T]
        extra = int(rng.integers(0, 40))
        ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, radius=radii, mass=masses,
                                         dot_fma=fma)
        h.check(h.lib.bl_set_small_kernel(h.h, kernel))
        try:
            for t in cuts:
                ens.run(-1, end_time=t)
            if extra:
                ens.run(extra)
        finally:
            h.check(h.lib.bl_set_small_kernel(h.h, 1))
        got = (ens.time, ens.balls_velocity, ens.balls_initial_position, ens.balls_position, ens.counts_ball_ball,
               ens.counts_ball_obstacle, ens.toi_tables)
        total = 0
        for s in range(S):
            o = oracle.OracleBilliard(workloads.oracle_obstacles(obstacles), dot_fma=fma)
            o.add_balls(pos[s], vel[s], radii[s] if per_system else radius, masses[s] if per_system else mass)
            cb = co = 0
            for t in cuts:
                a = o.evolve(t, max_events=200000)
                cb += a[0]
                co += a[1]
            if extra:
                a = o.evolve(INF, max_events=extra)
                cb += a[0]
                co += a[1]
            st = o.state()
            rows = o.toi_table
            ok = got[0][s] == o.time and same(got[1][s], st["v"]) and same(got[2][s], st["p0"]) and same(got[3][s], st["pos"]) \
                and (int(got[4][s]), int(got[5][s])) == (cb, co) \
                and all(same(got[6][s, i, :i], rows[i]) and same(got[6][s, :i, i], rows[i]) for i in range(n))
            if not ok:
                print(f"MISMATCH multi case {case}: n={n} kernel={kernel} S={S} system {s} seed={seed}", flush=True)
                sys.exit(1)
            total += cb + co
        print(f"multi case {case}: n={n:2d} kernel={kernel} S={S:2d} obstacles={len(obstacles)} cuts={len(cuts)} extra={extra:2d} per_system={int(per_system)} "
              f"collisions={total} ok", flush=True)
        # a one-ball ensemble on a random geometry
        nw, nd = int(rng.integers(0, 5)), int(rng.integers(0, 2))
        if nw + nd == 0:
            nw = 4
        ob, p1, v1, r1 = tge._fuzz_case(rng, nw, nd, case)
        res = {}
        for mode in (0, 1):
            h.check(h.lib.bl_set_ensemble_generic(h.h, mode))
            e1 = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, ob), p1, v1, radius=r1, dot_fma=fma, count_hits=True)
            e1.run(80, sync=False)
            res[mode] = (e1.time, e1.balls_velocity, e1.balls_initial_position, e1.counts, e1.hits, _lib.to_host(e1._status))
        h.check(h.lib.bl_set_ensemble_generic(h.h, 0))
        want = oracle.ensemble_run(workloads.oracle_obstacles(ob), p1, v1, 80, radius=r1, dot_fma=fma, count_hits=True,
                                   per_system_status=True)
        ref = (want["t"], want["v"], want["p0"], want["done"], want["hits"], want["status"])
        for mode in (0, 1):
            g = res[mode]
            if not (same(g[0], ref[0]) and same(g[1], ref[1]) and same(g[2], ref[2]) and g[3].tolist() == ref[3].tolist()
                    and g[4].tolist() == ref[4].tolist() and g[5].tolist() == ref[5].tolist()):
                print(f"MISMATCH one-ball case {case}: <{nw},{nd}> mode {mode} seed={seed}", flush=True)
                sys.exit(1)
        print(f"one-ball case {case}: <{nw},{nd}> collisions={int(want['done'].sum())} ok", flush=True)
    print("all", n_cases, "cases bit-identical")


if __name__ == "__main__":
    main()
