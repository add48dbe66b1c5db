#!/usr/bin/env python3
"""Per-phase cycle breakdown of bl_evolve_kernel (clock64 counters of CTA 0, returned in bl_result.prof) for a few
system sizes, plus Galperin / pool wall times.  Run on the GPU box: python tools/phase_profile.py [sides...]"""
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

fma = _lib.probe_dot_fma()
h = _lib.get_handle(0, fma)
if os.environ.get("BL_TARGET"):
    h.check(h.lib.bl_set_batch_target(h.h, int(os.environ["BL_TARGET"])))
NAMES = ("post+collect+sort", "resolve", "solve+violation", "repair")


def run(cfg, events, label):
    bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
    bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    bld.evolve_events(min(events, 2000))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = bld.evolve_events(events)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r = bld.last_result
    n = sum(c)
    ms = h.last_kernel_ms()
    prof = [int(x) for x in r.prof]
    rounds = max(prof[5], 1)
    per = {k: round(prof[q] / rounds) for k, q in (("post", 8), ("barrier1", 9), ("collect", 11), ("..resolved", 13), ("..rank only", 12), ("resolve+rank", 2), ("sort", 14), ("..cut+copy", 15), ("..row stores done", 0),
                                                    ("solve", 10), ("optimistic commit", 6), ("repair", 3))}
    print(f"{label}: N={cfg.n} events={n} kernel={ms:.2f} ms -> {n / ms * 1e3:,.0f} coll/s ({ms * 1e3 / max(n,1):.2f} us/event); "
          f"rounds {prof[5]} ({n / rounds:.1f} events/round, {prof[7]} re-posted); cycles/round {per}; "
          f"repair phases {prof[4]} rows {int(r.n_rescans)}; window {r.window:.3g}", flush=True)


if __name__ == "__main__":
    sides = [int(x) for x in sys.argv[1:]] or [8, 16, 32, 45, 64, 90, 128]
    g = workloads.galperin()
    b = g.build(billiards, dot_fma=fma)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = b.evolve(16)
    dt = time.perf_counter() - t0
    print(f"galperin: {c} in {dt * 1e3:.1f} ms wall, kernel {h.last_kernel_ms():.1f} ms -> {sum(c) / dt:,.0f} coll/s; "
          f"rounds {b.last_result.prof[5]} cycles/round {[int(x) // max(b.last_result.prof[5], 1) for x in b.last_result.prof[:4]]}", flush=True)
    p = workloads.pool().build(billiards, dot_fma=fma)
    t0 = time.perf_counter()
    c = p.evolve(100)
    print(f"pool: {c} in {(time.perf_counter() - t0) * 1e3:.2f} ms wall, kernel {h.last_kernel_ms():.3f} ms", flush=True)
    for s in sides:
        run(workloads.dense_gas(s), 20000 if s >= 64 else 50000, f"dense side {s}")
    run(workloads.ideal_gas(32), 50000, "ideal 1024")
