#!/usr/bin/env python3
"""Where do two schedules of the same dense-gas run first differ?  (diagnostic for the full-size parity test)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import billiards  # noqa: E402
import workloads  # noqa: E402
from billiards import _lib  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_events = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**7
variants = [(0, 10**6), (9, 777777), (0, 777777), (9, 10**6)]
fma = _lib.probe_dot_fma()
h = _lib.get_handle(None, fma)
cfg = workloads.dense_gas(side)
logs = []
for target, chunk in variants:
    h.check(h.lib.bl_set_batch_target(h.h, target))
    bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
    bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    T, I, J = [], [], []
    done = 0
    while done < n_events:
        k = min(chunk, n_events - done)
        counts, (t, i, j) = bld.evolve_events(k, log_capacity=k)
        done += sum(counts)
        T.append(t); I.append(i); J.append(j)
    h.check(h.lib.bl_set_batch_target(h.h, 0))
    logs.append((np.concatenate(T), np.concatenate(I), np.concatenate(J)))
    print("variant", target, chunk, "done", done, "time", bld.time, flush=True)
base = logs[0]
for v, lg in zip(variants[1:], logs[1:]):
    n = min(len(base[0]), len(lg[0]))
    bad = np.flatnonzero((base[0][:n].view(np.uint64) != lg[0][:n].view(np.uint64)) | (base[1][:n] != lg[1][:n]) | (base[2][:n] != lg[2][:n]))
    if bad.size == 0:
        print(v, "identical to", variants[0])
    else:
        e = int(bad[0])
        print(v, "first difference at event", e, "of", n, ":", len(bad), "differing records")
        for q in range(max(0, e - 3), min(n, e + 4)):
            print("   ", q, repr(base[0][q]), base[1][q], base[2][q], "|", repr(lg[0][q]), lg[1][q], lg[2][q])
