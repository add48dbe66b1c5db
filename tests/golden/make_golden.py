#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [--big]

The unmodified reference package is imported from /root/reference/src under the alias
``billiards_ref`` and driven through its public API (Billiard/add_ball/bounce_*/evolve).
Every fixture stores inputs implicitly (workloads.py generators + seed) and outputs explicitly
(event logs (t, i, j|-1-obstacle), end states, known-answer values) as raw float64/int64 bits.

``--big`` additionally runs the N=16384 dense gas (388 s of add_ball + ~20 s of events).

The fixtures were generated on a host whose numpy 2-vector dot is fma(a1*b1 + fl(a0*b0))
(recorded in each file as ``dot_fma``); tests pass that mode explicitly to the oracle and to
the CUDA path.
"""

import argparse
import ctypes
import importlib.util
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402

REF_SRC = "/root/reference/src/billiards"


def import_reference():
    spec = importlib.util.spec_from_file_location(
        "billiards_ref", os.path.join(REF_SRC, "__init__.py"), submodule_search_locations=[REF_SRC])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["billiards_ref"] = mod
    spec.loader.exec_module(mod)
    return mod


def probe_dot_fma():
    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(12345)
    a, b = rng.standard_normal((4000, 2)), rng.standard_normal((4000, 2))
    fused = all(x.dot(y) == libm.fma(x[1], y[1], x[0] * y[0]) for x, y in zip(a, b))
    plain = all(x.dot(y) == x[0] * y[0] + x[1] * y[1] for x, y in zip(a, b))
    assert fused != plain, "numpy dot matches neither candidate formula"
    return fused


def run_logged(bld, end_time=None, n_events=None, payload=False):
    """Step the reference event by event, recording (t, i, j) with j = -1-obstacle_index for walls/disks."""
    ts, ii, jj, pay = [], [], [], []
    obs_index = {id(o): q for q, o in enumerate(bld.obstacles)}

    while True:
        if n_events is not None and len(ts) >= n_events:
            break
        bb, bo = bld._next_ball_ball_collision, bld._next_ball_obstacle_collision
        tn = min(bb[0], bo[0])
        if end_time is not None and not tn <= end_time:
            break
        if not np.isfinite(tn):
            break
        rec = {}

        def cb_factory(idx):
            def cb(t, p, u, v, other):
                rec[idx] = (p.tolist(), u.tolist(), v.tolist())
            return cb

        if bb[0] <= bo[0]:
            i, j = int(bb[1]), int(bb[2])
            bld.bounce_ball_ball({i: cb_factory(i), j: cb_factory(j)} if payload else None)
            ts.append(float(bld.time)); ii.append(i); jj.append(j)
            if payload:
                pay.append(sum(rec[i], []) + sum(rec[j], []))
        else:
            i = int(bo[1])
            obs, args = bo[2]
            bld.bounce_ball_obstacle({i: cb_factory(i)} if payload else None)
            ts.append(float(bld.time)); ii.append(i); jj.append(-1 - obs_index[id(obs)])
            if payload:
                pay.append(sum(rec[i], []) + [float(args[0]) if args else 0.0, 0, 0, 0, 0, 0])
    if end_time is not None:
        bld._move(end_time)
    out = {"log_t": np.asarray(ts, dtype=np.float64), "log_i": np.asarray(ii, dtype=np.int64),
           "log_j": np.asarray(jj, dtype=np.int64)}
    if payload:
        out["log_payload"] = np.asarray(pay, dtype=np.float64).reshape(len(ts), 12)
    return out


def end_state(bld):
    bb = bld.next_ball_ball_collision
    bo = bld._next_ball_obstacle_collision
    obs_index = {id(o): q for q, o in enumerate(bld.obstacles)}
    bo_obs = -1 if bo[2] is None else obs_index[id(bo[2][0])]
    bo_arg = float(bo[2][1][0]) if (bo[2] is not None and bo[2][1]) else 0.0
    n = bld.count
    tri = np.concatenate([np.array(r, dtype=np.float64) for r in bld.toi_table]) if n > 1 else np.empty(0)
    return {
        "time": np.float64(bld.time),
        "pos": np.array(bld.balls_position, dtype=np.float64),
        "vel": np.array(bld.balls_velocity, dtype=np.float64),
        "t0": np.array(bld.balls_initial_time[:, 0], dtype=np.float64),
        "p0": np.array(bld.balls_initial_position, dtype=np.float64),
        "next_bb": np.asarray([bb[0], bb[1], bb[2]], dtype=np.float64),
        "next_bo": np.asarray([bo[0], int(bo[1]), bo_obs, bo_arg], dtype=np.float64),
        "toi_tri": tri,
        "balls_toi": np.array(bld._balls_toi, dtype=np.float64),
        "balls_idx": np.asarray([int(x) for x in bld._balls_idx], dtype=np.int64),
        "obs_toi": np.array(bld._obstacles_toi, dtype=np.float64),
        "obs_obs": np.asarray([-1 if o is None else obs_index[id(o[0])] for o in bld._obstacles_obs], dtype=np.int64),
    }


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.1f} KiB)")


def logged_case(ref, dot_fma, cfg, name, **kw):
    t0 = time.time()
    bld = cfg.build(ref)
    built = time.time() - t0
    start = end_state(bld)
    t0 = time.time()
    log = run_logged(bld, **kw)
    ran = time.time() - t0
    n_ev = len(log["log_t"])
    print(f"{name}: N={cfg.n} build {built:.2f}s, {n_ev} events in {ran:.2f}s "
          f"({n_ev / max(ran, 1e-9):.1f} coll/s), hash {workloads.event_log_hash(log['log_t'], log['log_i'], log['log_j'])}")
    out = {"dot_fma": np.int64(dot_fma), "ref_seconds": np.float64(ran), "ref_build_seconds": np.float64(built)}
    out.update(log)
    out.update({"end_" + k: v for k, v in end_state(bld).items() if k != "toi_tri" or cfg.n <= 256})
    out.update({"start_" + k: v for k, v in start.items() if k in ("next_bb", "next_bo", "balls_toi", "balls_idx",
                                                                   "obs_toi", "obs_obs") or (k == "toi_tri" and cfg.n <= 256)})
    save(name, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()

    ref = import_reference()
    dot_fma = probe_dot_fma()
    print("reference:", ref.__file__, "numpy", np.__version__, "dot_fma =", dot_fma)
    want = lambda n: args.only is None or args.only == n  # noqa: E731

    # ---- G: Galperin, README.md:114-135 / docs/quickstart.rst:50-131 ------------------------------
    if want("galperin"):
        cfg = workloads.galperin()
        bld = cfg.build(ref)
        first = bld.next_collision
        cum, tot = [], 0
        for t in (1, 2, 3, 4, 5):
            c = bld.evolve(t)
            tot += sum(c)
            cum.append(tot)
        c16 = bld.evolve(16)
        total = tot + sum(c16)
        v, m = bld.balls_velocity, np.asarray(bld.balls_mass)
        ke = float((m * (v * v).sum(axis=1)).sum() / 2)
        bld2 = cfg.build(ref)
        t0 = time.time()
        counts = bld2.evolve(16)
        dt = time.time() - t0
        print(f"galperin: {counts} in {dt:.2f}s, cum {cum}, total {total}, KE {ke!r}")
        bld3 = cfg.build(ref)
        log = run_logged(bld3, n_events=20000)
        save("galperin", dot_fma=np.int64(dot_fma), first=np.asarray(first[:3], dtype=np.float64),
             cumulative=np.asarray(cum, dtype=np.int64), counts16=np.asarray(counts, dtype=np.int64),
             total=np.int64(total), kinetic_energy=np.float64(ke), ref_seconds=np.float64(dt),
             log20k_state_pos=bld3.balls_position, log20k_state_vel=bld3.balls_velocity,
             **log, **{"end_" + k: v for k, v in end_state(bld2).items()})

    # ---- P: pool ----------------------------------------------------------------------------------
    if want("pool"):
        logged_case(ref, dot_fma, workloads.pool(), "pool", end_time=100.0, payload=True)

    # ---- H: heterogeneous gas ---------------------------------------------------------------------
    if want("hetero"):
        cfg = workloads.hetero_gas()
        # make the fixture independent of libm's pow(): all (r_a + r_b)**2 must equal the plain product
        r = cfg.radius
        s = (r[:, None] + r[None, :]).ravel().tolist() + (0.8 + r).tolist()
        assert all(x ** 2 == x * x for x in s), "radius set hits a pow() != s*s case; change the seed"
        logged_case(ref, dot_fma, cfg, "hetero_gas", n_events=4000)

    # ---- D small: dense gases ---------------------------------------------------------------------
    if want("dense256"):
        logged_case(ref, dot_fma, workloads.dense_gas(16), "dense_gas_256", n_events=2000)
    if want("dense1024"):
        logged_case(ref, dot_fma, workloads.dense_gas(32), "dense_gas_1024", n_events=1000)
    if want("ideal1024"):
        logged_case(ref, dot_fma, workloads.ideal_gas(32), "ideal_gas_1024", n_events=1000)
    if args.big and want("dense16384"):
        logged_case(ref, dot_fma, workloads.dense_gas(128), "dense_gas_16384", n_events=100)

    # ---- X: the collapsing cloud (examples/collapse.py): 200 balls, no obstacles --------------------------
    if want("collapse"):
        logged_case(ref, dot_fma, workloads.collapse(), "collapse", end_time=15.0)

    # ---- C: Newton's cradle (docs/usage.rst:80-158) --------------------------------------------------
    if want("cradle"):
        cfg = workloads.newtons_cradle(5, True)
        bld = cfg.build(ref)
        nb, no = bld.next_ball_ball_collision, bld.next_ball_obstacle_collision
        log = run_logged(bld, end_time=4.0)
        counts4 = (int((log["log_j"] >= 0).sum()), int((log["log_j"] < 0).sum()))
        bld.balls_position[2] += (0, 1e-10)
        bld.recompute_toi(2)
        after_nudge = end_state(bld)
        c30 = bld.evolve(30)
        print(f"cradle: next {nb} {no[:2]} counts(4)={counts4} after nudge evolve(30)={c30}")
        save("cradle", dot_fma=np.int64(dot_fma), next_bb=np.asarray(nb, dtype=np.float64),
             next_bo=np.asarray([no[0], no[1], 1, float(no[2][1][0])]), counts4=np.asarray(counts4),
             counts30=np.asarray(c30), **log, **{"nudge_" + k: v for k, v in after_nudge.items()},
             **{"end_" + k: v for k, v in end_state(bld).items()})

    # ---- step consistency + recompute_toi scenario (tests/test_simulation.py:461-583) ----------------
    if want("four"):
        obstacles = workloads.box_walls(-1, -1, 1, 1)
        pos = [(0, 0.2), (0, 0.5), (0, -0.8), (0.7, -0.8)]
        vel = [(1, 2), (-1, 2), (0, -1), (0, 0)]
        cfg = workloads.Config("four", obstacles, pos, vel, 0.2, 1.0)
        bld = cfg.build(ref)
        log = run_logged(bld, end_time=10.0)
        once = end_state(bld)
        b2 = cfg.build(ref)
        b2.evolve(1)
        b2.balls_position[0] = (0, 0)
        b2.recompute_toi(0)
        s1 = end_state(b2)
        b2.evolve(2)
        b2.balls_velocity[1] = (0, 0)
        b2.balls_radius[2] = 0.1
        b2.recompute_toi([1, 2])
        s2 = end_state(b2)
        b2.evolve(3)
        b2.recompute_toi()
        s3 = end_state(b2)
        c5 = b2.evolve(5)
        s5 = end_state(b2)
        print(f"four: {len(log['log_t'])} events to t=10; recompute scenario evolve(5) -> {c5}")
        save("four_balls", dot_fma=np.int64(dot_fma), counts5=np.asarray(c5), **log,
             **{"once_" + k: v for k, v in once.items()}, **{"s1_" + k: v for k, v in s1.items()},
             **{"s2_" + k: v for k, v in s2.items()}, **{"s3_" + k: v for k, v in s3.items()},
             **{"s5_" + k: v for k, v in s5.items()})

    # ---- S: Sinai systems -------------------------------------------------------------------------------
    if want("sinai"):
        obstacles, pos, vel = workloads.sinai(64, seed=0)
        n_coll = 2000
        out_p0, out_v, out_t = np.empty((64, 2)), np.empty((64, 2)), np.empty(64)
        hits = np.zeros((64, 5), dtype=np.int64)
        t0 = time.time()
        for s in range(64):
            bld = ref.Billiard(workloads.make_obstacles(ref, obstacles))
            bld.add_ball(tuple(pos[s].tolist()), tuple(vel[s].tolist()), radius=0)
            idx = {id(o): q for q, o in enumerate(bld.obstacles)}
            for _ in range(n_coll):
                hits[s, idx[id(bld._next_ball_obstacle_collision[2][0])]] += 1
                bld.bounce_ball_obstacle()
            out_p0[s], out_v[s], out_t[s] = bld.balls_position[0], bld.balls_velocity[0], bld.time
        dt = time.time() - t0
        print(f"sinai: 64 x {n_coll} collisions in {dt:.2f}s ({64 * n_coll / dt:.0f} coll/s); disk share "
              f"{hits[:, 4].sum() / hits.sum():.3f}")
        # one long system (SURVEY.md appendix B last row)
        bld = ref.Billiard(workloads.make_obstacles(ref, obstacles))
        bld.add_ball(tuple(pos[0].tolist()), tuple(vel[0].tolist()), radius=0)
        for _ in range(10000):
            bld.bounce_ball_obstacle()
        save("sinai", dot_fma=np.int64(dot_fma), n_coll=np.int64(n_coll), p0=out_p0, v=out_v, t=out_t, hits=hits,
             ref_seconds=np.float64(dt), long_t=np.float64(bld.time), long_pos=bld.balls_position[0],
             long_vel=bld.balls_velocity[0])

    # ---- unit vectors for the scalar kernels -----------------------------------------------------------------
    if want("unit"):
        from billiards_ref import physics as rp
        from billiards_ref import obstacles as ro
        rng = np.random.default_rng(7)
        n = 2000
        # clustered so that a good share reaches the sqrt branch
        p1, p2 = rng.uniform(-2, 2, (n, 2)), rng.uniform(-2, 2, (n, 2))
        v1, v2 = rng.standard_normal((n, 2)), rng.standard_normal((n, 2))
        aim = rng.random(n) < 0.6
        v2[aim] = v1[aim] + (p1[aim] - p2[aim]) * rng.uniform(0.2, 2, (int(aim.sum()), 1)) \
            + 0.2 * rng.standard_normal((int(aim.sum()), 2))
        r1, r2 = rng.uniform(0, 0.5, n), rng.uniform(0, 0.5, n)
        r1[::7] = 0.25
        r2[::7] = 0.25
        s = r1 + r2
        ok = np.asarray([x ** 2 == x * x for x in s.tolist()])
        r2[~ok] = r1[~ok]  # 2*r1 squared is exact enough: pow(2r,2) == (2r)*(2r) for these
        s = r1 + r2
        assert all(x ** 2 == x * x for x in s.tolist())
        toi = np.asarray([rp.toi_ball_ball(p1[k], v1[k], float(r1[k]), p2[k], v2[k], float(r2[k])) for k in range(n)])
        print(f"unit: toi finite share {np.isfinite(toi).mean():.3f}")
        # elastic collisions: approaching pairs in contact
        ang = rng.uniform(0, 2 * np.pi, n)
        q1 = rng.uniform(-3, 3, (n, 2))
        q2 = q1 + (r1 + r2 + 0.1)[:, None] * np.stack([np.cos(ang), np.sin(ang)], axis=1)
        u1, u2 = rng.standard_normal((n, 2)), rng.standard_normal((n, 2))
        m1, m2 = rng.uniform(0.1, 10, n), rng.uniform(0.1, 10, n)
        m1[::11] = 0.0
        w1, w2, sep = np.zeros((n, 2)), np.zeros((n, 2)), np.zeros(n, dtype=np.int64)
        for k in range(n):
            try:
                a, b = rp.elastic_collision(q1[k], u1[k], float(m1[k]), q2[k], u2[k], float(m2[k]))
                w1[k], w2[k] = a, b
            except ValueError:
                sep[k] = 1
        # wall and disk
        wall = ro.InfiniteWall((-1.3, -0.7), (2.1, 1.9), exterior="right")
        disk = ro.Disk((0.3, -0.2), 0.75)
        wp = rng.uniform(-3, 3, (n, 2))
        wv = rng.standard_normal((n, 2))
        wr = rng.uniform(0, 0.4, n)
        ok = np.asarray([(0.75 + x) ** 2 == (0.75 + x) * (0.75 + x) for x in wr.tolist()])
        wr[~ok] = 0.25
        wt, wh, wres = np.zeros(n), np.zeros(n), np.zeros((n, 2))
        dtm, dres, dsep = np.zeros(n), np.zeros((n, 2)), np.zeros(n, dtype=np.int64)
        for k in range(n):
            t, a = wall.detect_collision(wp[k], wv[k], float(wr[k]))
            wt[k] = t
            if a:
                wh[k] = a[0]
                wres[k] = wall.resolve_collision(wp[k], wv[k], float(wr[k]), *a)
            t, a = disk.detect_collision(wp[k], wv[k], float(wr[k]))
            dtm[k] = t
            try:
                dres[k] = disk.resolve_collision(wp[k], wv[k], float(wr[k]))
            except ValueError:
                dsep[k] = 1
        save("unit_vectors", dot_fma=np.int64(dot_fma), p1=p1, v1=v1, r1=r1, p2=p2, v2=v2, r2=r2, toi=toi,
             q1=q1, u1=u1, m1=m1, q2=q2, u2=u2, m2=m2, w1=w1, w2=w2, sep=sep,
             wall_start=np.asarray(wall.start_point, dtype=np.float64), wall_normal=wall._normal,
             disk=np.asarray([0.3, -0.2, 0.75]), wp=wp, wv=wv, wr=wr, wall_t=wt, wall_headway=wh, wall_res=wres,
             disk_t=dtm, disk_res=dres, disk_sep=dsep)

    # ---- M: LineSegment obstacles (examples/maze.py) and the segment / point kernels -------------------
    if want("maze"):
        logged_case(ref, dot_fma, workloads.maze(), "maze", n_events=3000, payload=True)
        logged_case(ref, dot_fma, workloads.maze_wide(), "maze_wide", n_events=4000, payload=True)

    if want("segment"):
        import importlib
        rp = importlib.import_module("billiards_ref.physics")
        ro = importlib.import_module("billiards_ref.obstacles")
        rng = np.random.default_rng(2024)
        n = 3000
        seg_pts = [((-0.3, -1.0), (-0.3, 0.5)), ((-0.2, -0.4), (0.2, 0.4)), ((1.25, 0.5), (-0.75, 2.0)), ((0, 0), (2, 0))]
        which = rng.integers(0, len(seg_pts), n)
        pos = rng.uniform(-2.5, 2.5, (n, 2))
        vel = rng.standard_normal((n, 2))
        rad = rng.uniform(0, 0.3, n)
        rad[::5] = 0.0
        # aim a third of the balls at the segment they are tested against, another part at its end points
        segs = [ro.LineSegment(a, b) for a, b in seg_pts]
        for k in range(n):
            a, b = np.asarray(seg_pts[which[k]], dtype=np.float64)
            if k % 3 == 0:
                target = a + rng.uniform(-0.1, 1.1) * (b - a)
                vel[k] = (target - pos[k]) * rng.uniform(0.3, 2.0)
            elif k % 3 == 1:
                target = (a if rng.random() < 0.5 else b) + 0.5 * rad[k] * rng.standard_normal(2)
                vel[k] = (target - pos[k]) * rng.uniform(0.3, 2.0)
        vel[7] = (0.0, 0.0)
        vel[8] = (1.0, 0.0)       # parallel to the last segment
        which[8] = 3
        pos[8] = (-1.0, 0.7)
        ok = np.asarray([x ** 2 == x * x for x in rad.tolist()])
        rad[~ok] = 0.125
        seg8 = np.zeros((n, 8))
        pt_t = np.zeros(n); par_t = np.zeros(n); par_u = np.zeros(n); par_code = np.zeros(n, dtype=np.int64)
        det_t = np.zeros(n); det_u = np.full(n, np.nan); det_code = np.zeros(n, dtype=np.int64)
        res = np.zeros((n, 2)); res0 = np.zeros((n, 2)); res1 = np.zeros((n, 2)); sep = np.zeros((n, 3), dtype=np.int64)
        for k in range(n):
            sg = segs[which[k]]
            seg8[k] = [sg.start_point[0], sg.start_point[1], sg._covector[0], sg._covector[1], sg._normal[0],
                       sg._normal[1], sg.end_point[0], sg.end_point[1]]
            pt_t[k] = rp.toi_ball_point(pos[k], vel[k], float(rad[k]), sg.end_point) if vel[k].any() else np.inf
            t, u = rp.toi_and_param_ball_segment(pos[k], vel[k], float(rad[k]), sg.start_point, sg._covector, sg._normal)
            par_t[k] = t
            par_code[k] = 3 if u is None else (0 if isinstance(u, float) else 1 + int(u))
            par_u[k] = float(u) if isinstance(u, float) else 0.0
            if vel[k].any():
                t, (u,) = sg.detect_collision(pos[k], vel[k], float(rad[k]))
            else:
                t, u = np.inf, None
            det_t[k] = t
            det_code[k] = 3 if u is None else (0 if isinstance(u, float) else 1 + int(u))
            det_u[k] = np.nan if u is None else float(u)
            for col, uu, out in ((0, 0.37, res), (1, 0, res0), (2, 1, res1)):
                try:
                    out[k] = sg.resolve_collision(pos[k], vel[k], float(rad[k]), uu)
                except ValueError:
                    sep[k, col] = 1
        print(f"segment: finite detect share {np.isfinite(det_t).mean():.3f}, face hits {(det_code == 0).mean():.3f}, "
              f"end-point tries {((det_code == 1) | (det_code == 2)).mean():.3f}")
        save("segment_vectors", dot_fma=np.int64(dot_fma), seg=seg8, pos=pos, vel=vel, rad=rad, point_t=pt_t,
             param_t=par_t, param_u=par_u, param_code=par_code, detect_t=det_t, detect_u=det_u, detect_code=det_code,
             resolve_face=res, resolve_start=res0, resolve_end=res1, resolve_sep=sep)


if __name__ == "__main__":
    main()
