#!/usr/bin/env python3
"""Kernel time of the full-size Sinai job (2^20 systems x 10^4 collisions) for the three one-ball ensemble kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n_coll = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 2, 1]
fma = _lib.probe_dot_fma()
h = _lib.get_handle(None, fma)
obstacles, pos, vel = workloads.sinai(n, seed=1000)
obs = workloads.make_obstacles(billiards, obstacles)
ref = None
for mode in modes:
    h.check(h.lib.bl_set_ensemble_generic(h.h, mode))
    best = None
    for rep in range(3):
        ens = billiards.BilliardEnsemble(obs, pos, vel, dot_fma=fma, count_hits=True)
        ens.run(n_coll)
        ms = h.last_kernel_ms()
        best = ms if best is None or ms < best else best
    got = (ens.time, ens.balls_velocity, ens.balls_initial_position, ens.hits)
    same = "-" if ref is None else all(np.array_equal(a.view(np.uint64) if a.dtype == np.float64 else a,
                                                        b.view(np.uint64) if b.dtype == np.float64 else b) for a, b in zip(got, ref))
    ref = got if ref is None else ref
    print({0: "box", 1: "generic", 2: "shaped"}[mode], f"{best:.2f} ms", f"{n * n_coll / best / 1e6:.2f} G collisions/s",
          "bit-identical to first:", same, flush=True)
h.check(h.lib.bl_set_ensemble_generic(h.h, 0))
