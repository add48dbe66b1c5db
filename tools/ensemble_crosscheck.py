#!/usr/bin/env python3
"""Cross-check of the two ensemble kernels (run on the GPU box):
    python tools/ensemble_crosscheck.py run out.npz [systems] [collisions]     # honours BL_ENSEMBLE_GENERIC=1
    python tools/ensemble_crosscheck.py compare a.npz b.npz
The shaped kernel (walls-then-disk lists: only the winning wall is divided for) must leave exactly the bits the generic
kernel (every obstacle evaluated as written in the reference) leaves, for every system."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

if sys.argv[1] == "run":
    import billiards  # noqa: E402
    import workloads  # noqa: E402
    from billiards import _lib  # noqa: E402
    n_sys = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    n_coll = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
    fma = _lib.probe_dot_fma()
    obstacles, pos, vel = workloads.sinai(n_sys, seed=77)
    # a few systems aimed exactly at corners and along walls: ties and near-ties between walls
    pos[:8] = [(0.5, 0.5), (-0.5, 0.5), (0.25, -0.75), (0.9, 0.9), (0.5, 0.5), (0.0, 0.75), (0.75, 0.0), (-0.6, -0.6)]
    vel[:8] = [(1, 1), (-1, 1), (1, -1), (0.6, 0.8), (0, 1), (1, 0), (0, -1), (-0.8, -0.6)]
    vel[:8] /= np.linalg.norm(vel[:8], axis=1)[:, None]
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, dot_fma=fma, count_hits=True)
    ens.run(n_coll)
    np.savez(sys.argv[2], t=ens.time, v=ens.balls_velocity, p0=ens.balls_initial_position, hits=ens.hits,
             ms=_lib.get_handle(None, fma).last_kernel_ms())
    print("ran", n_sys, "systems x", n_coll, "collisions; generic =", os.environ.get("BL_ENSEMBLE_GENERIC", "0"),
          "kernel ms", _lib.get_handle(None, fma).last_kernel_ms())
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    ok = True
    for key in ("t", "v", "p0"):
        same = np.array_equal(a[key].view(np.uint64), b[key].view(np.uint64))
        ok = ok and same
        print(key, "identical" if same else f"DIFFERENT in {int((a[key].view(np.uint64) != b[key].view(np.uint64)).sum())} words")
    same = np.array_equal(a["hits"], b["hits"])
    print("hits", "identical" if same else "DIFFERENT")
    print("kernel ms:", float(a["ms"]), "vs", float(b["ms"]))
    sys.exit(0 if ok and same else 1)
