#!/usr/bin/env python3
"""Heterogeneous gases (random radii and masses, one infinite and one zero mass, walls + disk): the grid-wide kernel
against the one-CTA kernel at several sizes, same box, same inputs."""
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
events = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for side in (8, 10, 12, 15, 20, 32):
    cfg = workloads.hetero_gas(side=side, seed=3)
    out = {}
    for mode, label in ((3, "cta"), (0, "grid")):
        if label == "cta" and cfg.n > 224:
            continue
        h.check(h.lib.bl_set_small_kernel(h.h, mode))
        bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
        bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
        c = bld.evolve_events(events)
        r = bld.last_result
        out[label] = (h.last_kernel_ms(), sum(c), int(r.prof[5]), int(r.prof[7]))
    h.check(h.lib.bl_set_small_kernel(h.h, 1))
    print(cfg.name, "N =", cfg.n, {k: f"{v[1] / v[0] / 1e3:.3f} M/s ({v[2]} rounds, {v[3]} redone)" for k, v in out.items()}, flush=True)
