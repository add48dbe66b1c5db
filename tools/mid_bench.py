#!/usr/bin/env python3
"""Systems of 33 ... 224 balls: the one-CTA kernel (bl_evolve_cta.cu) against the grid-wide event kernel on the same
box, same inputs (bl_set_small_kernel(h, 0) sends every size through the grid kernel, (h, 3) every size up to 224 balls through the one-CTA kernel); logs must agree bit for bit."""
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
h = _lib.get_handle(0, fma)
events = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
cases = [(f"dense gas {s * s}", workloads.dense_gas(s)) for s in (6, 8, 10, 12, 14)]
cases += [("collapse 200", workloads.collapse()), ("maze 200", workloads.maze()), ("hetero", workloads.hetero_gas())]
for name, cfg in cases:
    out = {}
    for mode, label in ((3, "cta"), (0, "grid")):
        h.check(h.lib.bl_set_small_kernel(h.h, mode))
        best = None
        for rep in range(2):
            bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
            bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
            ev = events if "collapse" not in name else 2000
            c, (t, i, j) = bld.evolve_events(ev, log_capacity=ev)
            ms = h.last_kernel_ms()
            best = ms if best is None or ms < best else best
        out[label] = (best, sum(c), t, i, j, bld.balls_velocity.copy())
    h.check(h.lib.bl_set_small_kernel(h.h, 1))
    a, b = out["cta"], out["grid"]
    same = np.array_equal(a[2].view(np.uint64), b[2].view(np.uint64)) and np.array_equal(a[3], b[3]) and \
        np.array_equal(a[4], b[4]) and np.array_equal(a[5].view(np.uint64), b[5].view(np.uint64))
    print(f"{name}: N={cfg.n} {a[1]} collisions: one CTA {a[0]:.2f} ms = {a[1] / a[0] / 1e3:.3f} M/s, grid kernel {b[0]:.2f} ms = "
          f"{b[1] / b[0] / 1e3:.3f} M/s, x{b[0] / a[0]:.2f}; bit-identical: {same}", flush=True)
