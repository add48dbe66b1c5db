#!/usr/bin/env python3
"""Stress test of the two schedules of bl_evolve_kernel (run on the GPU box):
    python tools/stress_schedules.py [n_events] [repetitions]           # batched rounds, repeated
    DBG_TARGET=1 python tools/stress_schedules.py [n_events] [reps]     # one collision per round, repeated
Every repetition first runs many short launches on another system (different CTA count, stops inside batches), then the
16384-ball gas for n_events collisions in one launch, and compares event log bit for bit with the first one-per-round
run.  The short rounds of the one-per-round schedule are what exposed the missing readers' gate (a CTA committing -- and
overwriting the global mirrors -- while a slower CTA was still resolving from them)."""
import os, sys, time, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import billiards, workloads
from billiards import _lib
fma = _lib.probe_dot_fma()
h = _lib.get_handle(0, fma)
INF = float("inf")

def build(cfg):
    b = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
    b.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    return b

def perturb():
    cfg = workloads.dense_gas(47, seed=11)
    bld = build(cfg)
    chunks = [1, 2, 3, 5, 7, 11, 13, 64, 1, 100, 257]
    done = 0; i = 0
    while done < 3000:
        k = min(chunks[i % len(chunks)], 3000 - done); i += 1
        bld.evolve_events(k, log_capacity=k); done += k
    a = build(cfg)
    for f in range(1, 101):
        a.evolve(0.05 * f / 100)

def long_run(target, n_events):
    h.check(h.lib.bl_set_batch_target(h.h, target))
    try:
        bld = build(workloads.dense_gas(int(os.environ.get("DBG_SIDE", "128"))))
        t0 = time.perf_counter()
        counts, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        h.check(h.lib.bl_set_batch_target(h.h, 0))
    return counts, t, i, j, int(bld.last_result.prof[5]), int(bld.last_result.prof[7]), dt

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ref = None
for rep in range(reps):
    perturb()
    print("rep", rep, "perturbed", flush=True)
    c, t, i, j, rounds, redo, dt = long_run(int(os.environ.get("DBG_TARGET", "0")), n_events)
    d = hashlib.sha256(t.tobytes() + i.tobytes() + j.tobytes()).hexdigest()[:16]
    print("rep", rep, "batched", c, d, "rounds", rounds, "redo", redo, f"{dt:.2f}s", flush=True)
    if ref is None:
        c1, t1, i1, j1, r1, _, dt1 = long_run(1, n_events)
        ref = (c1, t1, i1, j1)
        print("reference (one per round)", c1, hashlib.sha256(t1.tobytes() + i1.tobytes() + j1.tobytes()).hexdigest()[:16], f"{dt1:.2f}s", flush=True)
    ok = c == ref[0] and np.array_equal(t.view(np.uint64), ref[1].view(np.uint64)) and np.array_equal(i, ref[2]) and np.array_equal(j, ref[3])
    if not ok:
        m = min(len(t), len(ref[1]))
        bad = np.flatnonzero((t[:m].view(np.uint64) != ref[1][:m].view(np.uint64)) | (i[:m] != ref[2][:m]) | (j[:m] != ref[3][:m]))
        print("MISMATCH first at", bad[:5], "of", m, flush=True)
        if len(bad):
            b0 = int(bad[0])
            for q in range(max(0, b0 - 3), min(m, b0 + 4)):
                print(q, "got", repr(float(t[q])), int(i[q]), int(j[q]), " ref", repr(float(ref[1][q])), int(ref[2][q]), int(ref[3][q]), flush=True)
        sys.exit(1)
print("all ok")
