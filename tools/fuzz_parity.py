#!/usr/bin/env python3
"""Randomised parity sweep (run on the GPU box): seeded gases of many sizes, densities, batch targets and obstacle sets,
event log + state + table compared bit for bit with the CPU oracle.  Not part of the test suite (minutes of CPU time);
    python tools/fuzz_parity.py [n_cases] [seed]
    FUZZ_MID=1 FUZZ_SMALL_KERNEL=3 python tools/fuzz_parity.py 60 5    # 36 ... 196 balls through the one-CTA kernel
prints one line per case and exits non-zero on the first mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402
from oracle import oracle  # noqa: E402

INF = float("inf")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    for case in range(n_cases):
        sides = [3, 5, 6, 9, 13, 20, 31, 33, 46, 52, 70]
        if os.environ.get("FUZZ_MID"):  # 33 ... 224 balls only: the one-CTA kernel (with FUZZ_SMALL_KERNEL=3: at every size)
            sides = [6, 7, 8, 9, 10, 11, 12, 13, 14]
        side = int(rng.choice(sides))
        phi = float(rng.choice([0.05, 0.2, 0.4, 0.55, 0.7]))
        target = int(rng.choice([0, 0, 1, 2, 3, 7, 20, 40]))
        hetero = side <= 46 and rng.random() < 0.4
        if hetero:
            cfg = workloads.hetero_gas(side=max(side, 6), seed=int(rng.integers(1 << 30)))
        else:
            cfg = workloads.dense_gas(side, seed=int(rng.integers(1 << 30)), phi=phi)
        if rng.random() < 0.3:  # a segment through the box, balls near it removed
            box = float(side if not hetero else max(side, 6))
            a, b = np.asarray([0.21 * box, 0.18 * box]), np.asarray([0.63 * box, 0.71 * box])
            pos = np.asarray(cfg.pos)
            u = np.clip(((pos - a) @ (b - a)) / ((b - a) @ (b - a)), 0.0, 1.0)
            keep = np.hypot(*(pos - (a + u[:, None] * (b - a))).T) > np.asarray(cfg.radius) * np.ones(len(pos)) + 0.05
            cfg = workloads.Config(cfg.name + "_seg", cfg.obstacles + [("segment", tuple(a), tuple(b))], pos[keep],
                                   np.asarray(cfg.vel)[keep], (np.asarray(cfg.radius) * np.ones(len(pos)))[keep],
                                   (np.asarray(cfg.mass) * np.ones(len(pos)))[keep])
        n_events = int(min(6000, max(300, 40 * cfg.n))) if cfg.n < 3000 else 2500
        chunks = int(rng.choice([1, 1, 3, 17]))
        ref = cfg.build_oracle(oracle, dot_fma=fma, flags=oracle.FAST_ROWMIN)
        ref.enable_log(n_events)
        ref.evolve(INF, max_events=n_events)
        rt, ri, rj = ref.log()
        h.check(h.lib.bl_set_batch_target(h.h, target))
        h.check(h.lib.bl_set_small_kernel(h.h, int(os.environ.get("FUZZ_SMALL_KERNEL", "1"))))
        try:
            bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
            bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
            ts, ii, jj = [], [], []
            left = len(rt)
            while left > 0:
                k = max(1, min(left, -(-len(rt) // chunks)))
                _, (t, i, j) = bld.evolve_events(k, log_capacity=k)
                ts.append(t); ii.append(i); jj.append(j)
                left -= k
                if len(t) < k:
                    break
        finally:
            h.check(h.lib.bl_set_batch_target(h.h, 0))
            h.check(h.lib.bl_set_small_kernel(h.h, 1))
        t, i, j = np.concatenate(ts), np.concatenate(ii), np.concatenate(jj)
        ok = (len(t) == len(rt) and np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(bits(t), bits(rt))
              and np.array_equal(bits(bld.balls_velocity), bits(ref.state()["v"]))
              and np.array_equal(bits(np.concatenate(bld.toi_table)), bits(np.concatenate(ref.toi_table))))
        print(f"case {case}: {cfg.name} N={cfg.n} events={len(rt)} target={target} chunks={chunks} "
              f"rounds={int(bld.last_result.prof[5])} -> {'ok' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            bad = np.flatnonzero((i[:len(ri)] != ri[:len(i)]) | (j[:len(rj)] != rj[:len(j)]) | (bits(t)[:len(rt)] != bits(rt)[:len(t)]))
            print("first difference at event", bad[:3], file=sys.stderr)
            sys.exit(1)
    print("all cases ok")


if __name__ == "__main__":
    main()
