"""GPU parity tests of the ensemble front ends (run on the B200 box: ``pytest -m gpu``).

* the one-ball ensemble kernel unrolled for "walls, then a disk" lists (``bl_ensemble_kernel_u``) against the generic
  kernel and the CPU oracle on randomised geometries -- every ``<NW, ND>`` instance, oblique walls, scales 1e-6 .. 1e6,
  random disk and ball radii, shots aimed at corners and tangent to the disk, balls starting in contact;
* its shared-reciprocal division against ``__ddiv_rn`` on 2^30 operand triples;
* multi-ball ensembles (one group of lanes per system) against the oracle, against the single-``Billiard`` path and at
  full size (2^16 pool tables).

Reference lines: simulation.py:354-438, :504-554; obstacles.py:70-76, :112-144; physics.py:33-76, :273-282.
"""

import ctypes
from math import pi

import numpy as np
import pytest

import workloads
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu

INF = float("inf")


@pytest.fixture(scope="module")
def billiards():
    import billiards as b
    return b


def _polygon(rng, n_vertices):
    """random convex polygon around the origin, counter-clockwise (interior on the left of every edge)"""
    while True:
        ang = np.sort(rng.uniform(0, 2 * pi, n_vertices))
        gaps = np.diff(np.concatenate([ang, [ang[0] + 2 * pi]]))
        if gaps.max() < 0.9 * pi and gaps.min() > 0.3:
            break
    rad = rng.uniform(1.0, 1.3, n_vertices)
    pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1)
    # keep it convex: a vertex inside the hull of the others is pushed out by using the unit circle instead
    for k in range(n_vertices):
        a, b, c = pts[k - 1], pts[k], pts[(k + 1) % n_vertices]
        if (b[0] - a[0]) * (c[1] - b[1]) - (b[1] - a[1]) * (c[0] - b[0]) <= 0:
            pts = np.stack([np.cos(ang), np.sin(ang)], axis=1) * 1.15
            break
    return pts


def _fuzz_case(rng, nw, nd, case):
    """obstacle list (user terms) + initial conditions for one randomised one-ball ensemble"""
    scale = 10.0 ** rng.uniform(-6, 6)
    shift = rng.uniform(-3, 3, 2) * scale * (case % 3 != 0)
    axis_aligned = case % 4 == 0 or (nw == 4 and case % 2 == 1)  # the box shortcut of the shaped kernel
    if axis_aligned:
        verts = np.asarray([(-1.0, -1.0), (1.0, -1.0), (1.0, 1.0), (-1.0, 1.0)]) * rng.uniform(0.8, 1.3)
    else:
        verts = _polygon(rng, max(nw, 3) if nw != 4 else 4)
    nv = len(verts)
    edges = [(verts[k], verts[(k + 1) % nv]) for k in range(nv)]
    picked = list(rng.permutation(nv)[:nw]) if nw < nv else list(range(nv))
    obstacles = [("wall", tuple((edges[k][0] * scale + shift).tolist()), tuple((edges[k][1] * scale + shift).tolist()),
                  "left") for k in picked]
    R = 0.0
    centre = np.zeros(2)
    if nd:
        R = rng.uniform(0.05, 0.2)
        centre = rng.uniform(-0.1, 0.1, 2)
        obstacles.append(("disk", tuple((centre * scale + shift).tolist()), float(R * scale)))
    radius = 0.0 if case % 2 == 0 else float(rng.uniform(0.0, 0.05))
    # positions inside the full polygon (the walls that were not picked are simply absent), clear of the disk
    n_sys = 1500
    pos = np.empty((0, 2))
    normals = []
    for a, b in edges:
        d = b - a
        nrm = np.asarray([-d[1], d[0]]) / np.hypot(*d)
        normals.append((a, nrm))
    while len(pos) < n_sys:
        p = rng.uniform(-1.3, 1.3, (4 * n_sys, 2))
        ok = np.ones(len(p), dtype=bool)
        for a, nrm in normals:
            ok &= (p - a) @ nrm > radius + 0.01
        if nd:
            ok &= np.hypot(*(p - centre).T) > R + radius + 0.01
        pos = np.concatenate([pos, p[ok]])
    pos = pos[:n_sys]
    theta = rng.uniform(0, 2 * pi, n_sys)
    # special shots: at a vertex of the polygon (two walls within rounding of each other), tangent to the disk
    k = np.arange(n_sys)
    aim = k % 5 == 1
    tv = verts[rng.integers(0, nv, n_sys)]
    theta = np.where(aim, np.arctan2(tv[:, 1] - pos[:, 1], tv[:, 0] - pos[:, 0]), theta)
    if nd:
        tang = k % 5 == 2
        dc = centre - pos
        dist = np.hypot(*dc.T)
        eps = rng.choice([0.0, 1e-15, -1e-15, 1e-12, -1e-12, 1e-9, -1e-9, 1e-6, -1e-6], n_sys)
        off = np.arcsin(np.clip((R + radius) / dist, 0, 1)) * rng.choice([-1.0, 1.0], n_sys) + eps
        theta = np.where(tang, np.arctan2(dc[:, 1], dc[:, 0]) + off, theta)
    speed = 10.0 ** rng.uniform(-2, 2, n_sys)
    vel = np.stack([np.cos(theta), np.sin(theta)], axis=1) * speed[:, None] * scale
    # motion exactly along an axis: one velocity component is a zero of either sign
    for q in range(3, 203, 5):
        vel[q, q % 2] = 0.0 if q % 3 else -0.0
    pos = pos * scale + shift
    # balls that start in contact with a wall: exactly at distance `radius` along its normal (gap = 0 up to rounding)
    if nw:
        a, b = (np.asarray(x) for x in obstacles[0][1:3])
        d = b - a
        nrm = np.asarray([-d[1], d[0]]) / np.hypot(*d)
        for q in range(20):
            u = rng.uniform(0.3, 0.7)
            pos[n_sys - 1 - q] = a + u * d + nrm * (radius * scale) * (1.0 + (q % 3 - 1) * 1e-16)
    return obstacles, pos, vel, float(radius * scale)


@pytest.mark.parametrize("nw,nd", [(1, 0), (2, 0), (3, 0), (4, 0), (0, 1), (1, 1), (2, 1), (3, 1), (4, 1)])
def test_shaped_ensemble_kernel_on_random_geometries(billiards, oracle, nw, nd):
    """bl_ensemble_kernel_u<NW, ND> == generic kernel == oracle, bit for bit, off the Sinai set-up: 24 random obstacle
    lists per instance (216 in all)."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    rng = np.random.default_rng(1000 * nw + 100 * nd + 7)
    n_coll = 60
    total = 0
    for case in range(24):
        obstacles, pos, vel, radius = _fuzz_case(rng, nw, nd, case)
        obs = workloads.make_obstacles(billiards, obstacles)
        got = {}
        for name, generic in (("shaped", 0), ("generic", 1), ("shaped_no_box", 2)):
            h.check(h.lib.bl_set_ensemble_generic(h.h, generic))
            try:
                ens = billiards.BilliardEnsemble(obs, pos, vel, radius=radius, dot_fma=fma, count_hits=True)
                ens.run(n_coll, sync=False)
                got[name] = (ens.time, ens.balls_velocity, ens.balls_initial_position, ens.counts, ens.hits,
                             _lib.to_host(ens._status))
            finally:
                h.check(h.lib.bl_set_ensemble_generic(h.h, 0))
        want = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos, vel, n_coll, radius=radius, dot_fma=fma,
                                   count_hits=True, per_system_status=True)
        ref = (want["t"], want["v"], want["p0"], want["done"], want["hits"], want["status"])
        for name in ("generic", "shaped_no_box", "shaped"):
            what = f"<{nw},{nd}> case {case} {name}"
            for q, label in enumerate(("time", "velocity", "position")):
                assert_bits_equal(got[name][q], ref[q], f"{what}: {label}")
            assert got[name][3].tolist() == ref[3].tolist(), what
            assert got[name][4].tolist() == ref[4].tolist(), what
            assert got[name][5].tolist() == ref[5].tolist(), what
        total += int(want["done"].sum())
    # closed shapes bounce for ever (but for a few axis-parallel shots out of an open box); open ones (fewer than three
    # walls) let most balls escape after a few hits
    assert total > 0.97 * 24 * 1500 * n_coll if nw >= 3 else total > 5000


def test_shared_reciprocal_division_is_ieee(billiards):
    """bl_div2_same_den (the Disk reflection's two quotients over one denominator) against __ddiv_rn: 2^30 operand
    triples = 2^31 quotients, incl. subnormal / huge / non-finite operands, must agree in every bit."""
    from billiards import _lib
    h = _lib.get_handle()
    bad = ctypes.c_int64(-1)
    first = (ctypes.c_double * 3)()
    h.check(h.lib.bl_selftest_div2(h.h, 12345, 1 << 30, ctypes.byref(bad), first, h.stream()))
    assert bad.value == 0, f"{bad.value} quotients differ from __ddiv_rn, e.g. {first[0]!r} / {first[1]!r} -> {first[2]!r}"


# ---------------------------------------------------------------------------------------------------------------------
# multi-ball ensembles
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_systems(oracle, obstacles, pos, vel, radius, mass, fma):
    out = []
    for s in range(pos.shape[0]):
        o = oracle.OracleBilliard(workloads.oracle_obstacles(obstacles), dot_fma=fma)
        o.add_balls(pos[s], vel[s], radius, mass)
        out.append(o)
    return out


def _small_gas(rng, n, S):
    """S systems of n balls on a jittered lattice in a box with a central disk; per-slot radii and masses, one ball
    of infinite mass when there is room.  (No zero-mass ball here: between two approaching heavy balls it gains speed
    with every bounce and collides infinitely often in finite time -- evolve(end_time) would never return, in the
    reference as well; tests/golden/hetero_gas.npz covers zero mass with a count-based stop.)"""
    side = int(np.ceil(np.sqrt(n))) + 2
    box = float(side)
    cells = [(ix + 0.5, iy + 0.5) for ix in range(side) for iy in range(side)
             if np.hypot(ix + 0.5 - box / 2, iy + 0.5 - box / 2) > 1.0]
    assert len(cells) >= n
    radius = np.round(rng.uniform(0.1, 0.3, n) * 1024.0) / 1024.0
    if n >= 4:
        radius[1] = radius[0]  # a shared radius class
        radius[3] = 0.0        # a point particle
    mass = rng.uniform(0.5, 5.0, n)
    if n >= 6:
        mass[2] = INF
    pos = np.empty((S, n, 2))
    vel = np.empty((S, n, 2))
    for s in range(S):
        pick = rng.permutation(len(cells))[:n]
        pos[s] = np.asarray(cells)[pick] + rng.uniform(-0.15, 0.15, (n, 2))
        th = rng.uniform(0, 2 * pi, n)
        vel[s] = np.stack([np.cos(th), np.sin(th)], axis=1) * rng.uniform(0.3, 1.5, n)[:, None]
        if n >= 6:
            vel[s, 2] = 0.0  # the infinite mass rests: a moving one squeezes a ball against a wall without bound
    obstacles = workloads.box_walls(0.0, 0.0, box, box) + [("disk", (box / 2, box / 2), 0.5)]
    return obstacles, pos, vel, radius, mass


@pytest.mark.parametrize("n,kernel", [(1, 1), (2, 1), (3, 1), (4, 1), (1, 2), (2, 2), (3, 2), (4, 2), (5, 1), (8, 1),
                                      (9, 1), (16, 1), (17, 1), (31, 1), (32, 1), (33, 1), (47, 1), (64, 1), (100, 1),
                                      (224, 1)])
def test_multi_ball_ensemble_against_oracle(billiards, oracle, n, kernel):
    """S systems x n balls -- one thread per system up to 4 balls (kernel 1), one group of W lanes per system up to 32
    (kernel 2 forces it for n <= 4 too), one CTA per system from 33 to 224 balls: evolve(T/2) + evolve(T) + run(count) == the oracle's evolve on each system,
    bit for bit (state, clocks, counts, the whole time-of-impact table)."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    rng = np.random.default_rng(50 + n)
    S = 37 if n > 4 else 150
    if n > 32:
        S = 11 if n <= 100 else 4
    obstacles, pos, vel, radius, mass = _small_gas(rng, n, S)
    if n <= 4 and kernel == 1:
        obstacles = obstacles[:4]  # walls only: the one-thread kernel's two-balls-per-wall loop
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, radius=radius, mass=mass,
                                     dot_fma=fma)
    T = 12.0
    h.check(h.lib.bl_set_small_kernel(h.h, kernel))
    try:
        c1 = ens.evolve(T / 2)
        c2 = ens.evolve(T)
        ens.run(25)
    finally:
        h.check(h.lib.bl_set_small_kernel(h.h, 1))
    refs = _oracle_systems(oracle, obstacles, pos, vel, radius, mass, fma)
    got_t, got_v, got_p0, got_pos = ens.time, ens.balls_velocity, ens.balls_initial_position, ens.balls_position
    got_bb, got_bo, tables = ens.counts_ball_ball, ens.counts_ball_obstacle, ens.toi_tables
    for s, o in enumerate(refs):
        a = o.evolve(T / 2)
        b = o.evolve(T)
        c = o.evolve(INF, max_events=25)
        assert (int(c1[0][s]), int(c1[1][s])) == a and (int(c2[0][s]), int(c2[1][s])) == b, f"system {s}"
        assert (int(got_bb[s]), int(got_bo[s])) == (a[0] + b[0] + c[0], a[1] + b[1] + c[1]), f"system {s}"
        st = o.state()
        assert got_t[s] == o.time, f"system {s}"
        assert_bits_equal(got_v[s], st["v"], f"system {s}: velocities")
        assert_bits_equal(got_p0[s], st["p0"], f"system {s}: initial positions")
        assert_bits_equal(got_pos[s], st["pos"], f"system {s}: positions")
        rows = o.toi_table
        for i in range(n):
            assert_bits_equal(tables[s, i, :i], rows[i], f"system {s}: table row {i}")
            assert_bits_equal(tables[s, :i, i], rows[i], f"system {s}: table column {i}")


def test_pool_ensemble_matches_single_path_and_oracle(billiards, oracle):
    """Pool tables with different break shots: ensemble == oracle per system == a single Billiard built from the same
    numbers (the kernel is the same, the launch shape differs)."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    cfg = workloads.pool()
    S = 48
    rng = np.random.default_rng(3)
    pos = np.repeat(cfg.pos[None], S, axis=0)
    vel = np.repeat(cfg.vel[None], S, axis=0)
    ang = rng.uniform(-0.05, 0.05, S)
    ang[0] = 0.0
    speed = cfg.vel[15, 0] * rng.uniform(0.5, 1.5, S)
    speed[0] = cfg.vel[15, 0]
    vel[:, 15, 0] = speed * np.cos(ang)
    vel[:, 15, 1] = speed * np.sin(ang)
    obs = workloads.make_obstacles(billiards, cfg.obstacles)
    ens = billiards.BilliardEnsemble(obs, pos, vel, radius=cfg.radius, mass=cfg.mass, dot_fma=fma)
    bb, bo = ens.evolve(100.0)
    assert (int(bb[0]), int(bo[0])) == (167, 248)  # system 0 is examples/pool_first_shot.py itself
    got_v, got_pos = ens.balls_velocity, ens.balls_position
    for s, o in enumerate(_oracle_systems(oracle, cfg.obstacles, pos, vel, cfg.radius, cfg.mass, fma)):
        assert o.evolve(100.0) == (int(bb[s]), int(bo[s])), f"table {s}"
        assert_bits_equal(got_v[s], o.balls_velocity, f"table {s}")
        assert_bits_equal(got_pos[s], o.balls_position, f"table {s}")
    for s in (0, 7, 31):
        one = billiards.Billiard(obs, dot_fma=fma)
        one.add_balls(pos[s], vel[s], cfg.radius, cfg.mass)
        assert one.evolve(100.0) == (int(bb[s]), int(bo[s]))
        assert_bits_equal(one.balls_velocity, got_v[s])
        assert_bits_equal(one.balls_position, got_pos[s])
    assert ens.gather_stats().tolist() == [int(bb.sum() + bo.sum()), int(bb.sum()), int(bo.sum())]


def test_full_size_pool_ensemble(billiards, golden):
    """2^16 pool tables (examples/pool_first_shot.py:17-35) to t = 100: every table reports (167, 248) and ends in the
    state of the single-system path (= the live reference's, tests/golden/pool.npz), bit for bit."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    g = golden("pool")
    cfg = workloads.pool()
    S = 1 << 16
    pos = np.ascontiguousarray(np.broadcast_to(cfg.pos[None], (S,) + cfg.pos.shape))
    vel = np.ascontiguousarray(np.broadcast_to(cfg.vel[None], (S,) + cfg.vel.shape))
    obs = workloads.make_obstacles(billiards, cfg.obstacles)
    ens = billiards.BilliardEnsemble(obs, pos, vel, radius=cfg.radius, mass=cfg.mass, dot_fma=fma)
    bb, bo = ens.evolve(100.0)
    assert (bb == 167).all() and (bo == 248).all()
    one = billiards.Billiard(obs, dot_fma=fma)
    one.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    assert one.evolve(100.0) == (167, 248)
    v, p = ens.balls_velocity, ens.balls_position
    assert (v.view(np.uint64) == one.balls_velocity.view(np.uint64)[None]).all()
    assert (p.view(np.uint64) == one.balls_position.view(np.uint64)[None]).all()
    if bool(g["dot_fma"]) == fma and "end_vel" in g:
        assert_bits_equal(v[S - 1], g["end_vel"])
    assert (ens.time == 100.0).all()
    assert ens.gather_stats().tolist() == [S * 415, S * 167, S * 248]


def test_run_from_host_matches_the_resident_path(billiards):
    """ensemble.run_from_host (blocks of systems pipelined over streams, pinned buffers) == BilliardEnsemble.run on the
    whole set, for one-ball and multi-ball systems: the partition into blocks must not change a bit."""
    from billiards import _lib
    from billiards.ensemble import run_from_host
    fma = _lib.probe_dot_fma()
    obstacles, pos, vel = workloads.sinai(5001, seed=9)
    obs = workloads.make_obstacles(billiards, obstacles)
    ens = billiards.BilliardEnsemble(obs, pos, vel, dot_fma=fma)
    ens.run(300)
    for chunks in (1, 3, 4):
        out = run_from_host(obs, pos, vel, 300, dot_fma=fma, chunks=chunks)
        assert_bits_equal(out["time"], ens.time, f"{chunks} chunks")
        assert_bits_equal(out["velocity"], ens.balls_velocity, f"{chunks} chunks")
        assert_bits_equal(out["position"], ens.balls_initial_position, f"{chunks} chunks")
        assert out["counts"].tolist() == ens.counts.tolist() and int(out["stats"][0]) == 5001 * 300
    cfg = workloads.pool()
    S = 37
    rng = np.random.default_rng(12)
    mpos = np.repeat(cfg.pos[None], S, axis=0)
    mvel = np.repeat(cfg.vel[None], S, axis=0)
    mvel[:, 15, 0] *= rng.uniform(0.7, 1.3, S)
    mvel[:, 15, 1] = mvel[:, 15, 0] * rng.uniform(-0.03, 0.03, S)
    pobs = workloads.make_obstacles(billiards, cfg.obstacles)
    whole = billiards.BilliardEnsemble(pobs, mpos, mvel, radius=cfg.radius, mass=cfg.mass, dot_fma=fma)
    whole.run(-1, end_time=100.0)
    out = run_from_host(pobs, mpos, mvel, -1, end_time=100.0, radius=cfg.radius, mass=cfg.mass, dot_fma=fma, chunks=4)
    assert_bits_equal(out["time"], whole.time)
    assert_bits_equal(out["velocity"], whole.balls_velocity)
    assert_bits_equal(out["position"], whole.balls_position)
    assert out["counts"].tolist() == whole.counts.tolist()
    assert out["stats"].tolist() == whole.gather_stats().tolist()
    # systems of 49 balls: one CTA per system
    gas = workloads.dense_gas(7, seed=2)
    S = 23
    th = rng.uniform(0, 2 * pi, (S, gas.n))
    gpos = np.repeat(gas.pos[None], S, axis=0)
    gvel = np.stack([np.cos(th), np.sin(th)], axis=2)
    gobs = workloads.make_obstacles(billiards, gas.obstacles)
    whole = billiards.BilliardEnsemble(gobs, gpos, gvel, radius=float(gas.radius[0]), mass=1.0, dot_fma=fma)
    whole.run(-1, end_time=1.5)
    out = run_from_host(gobs, gpos, gvel, -1, end_time=1.5, radius=float(gas.radius[0]), mass=1.0, dot_fma=fma, chunks=3)
    assert_bits_equal(out["time"], whole.time)
    assert_bits_equal(out["velocity"], whole.balls_velocity)
    assert_bits_equal(out["position"], whole.balls_position)
    assert out["counts"].tolist() == whole.counts.tolist() and int(out["stats"][0]) > 1000


def test_library_calls_on_a_non_current_device(billiards, oracle):
    """A Billiard and both ensemble kinds on cuda:1 while cuda:0 is torch's current device (every C entry makes the
    handle's device current for the call)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    torch.cuda.set_device(0)
    cfg = workloads.pool()
    obs = workloads.make_obstacles(billiards, cfg.obstacles)
    bld = billiards.Billiard(obs, device=1, dot_fma=fma)
    bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    assert bld.evolve(100.0) == (167, 248)
    assert torch.cuda.current_device() == 0
    # the host mirrors: written by a kernel on cuda:1 into page-locked memory that was allocated with cuda:0 current
    b0 = billiards.Billiard(obs, device=0, dot_fma=fma)
    b0.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    assert b0.evolve(100.0) == (167, 248)
    assert_bits_equal(bld.balls_position, b0.balls_position, "positions read back from cuda:1")
    assert_bits_equal(bld.balls_velocity, b0.balls_velocity, "velocities read back from cuda:1")
    gas = workloads.dense_gas(20, seed=1)
    g = billiards.Billiard(workloads.make_obstacles(billiards, gas.obstacles), device=1, dot_fma=fma)
    g.add_balls(gas.pos, gas.vel, gas.radius, gas.mass)
    _, (t, i, j) = g.evolve_events(800, log_capacity=800)
    ref = gas.build_oracle(oracle, dot_fma=fma, flags=oracle.FAST_ROWMIN)
    ref.enable_log(800)
    ref.evolve(INF, max_events=800)
    assert_bits_equal(t, ref.log()[0])
    obstacles, pos, vel = workloads.sinai(512, seed=4)
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, device=1, dot_fma=fma)
    ens.run(100)
    want = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos, vel, 100, dot_fma=fma)
    assert_bits_equal(ens.time, want["t"])
    multi = billiards.BilliardEnsemble(obs, np.repeat(cfg.pos[None], 8, axis=0), np.repeat(cfg.vel[None], 8, axis=0),
                                       radius=cfg.radius, mass=cfg.mass, device=1, dot_fma=fma)
    bb, bo = multi.evolve(100.0)
    assert (bb == 167).all() and (bo == 248).all()
    assert torch.cuda.current_device() == 0


@pytest.mark.parametrize("n", [2, 3, 4, 8, 16, 32, 40])
def test_per_system_radii_and_masses_against_oracle(billiards, oracle, n):
    """radius / mass of shape (S, n): every system of the ensemble has its own composition (per-system radius classes and
    (r_a + r_b)**2 tables).  Each system == the oracle run on that system's balls, bit for bit; systems that share a
    composition with the slot-wise ensemble give what it gives."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    rng = np.random.default_rng(700 + n)
    S = 21 if n <= 32 else 6
    obstacles, pos, vel, radius, mass = _small_gas(rng, n, S)
    radii = np.repeat(radius[None], S, axis=0)
    masses = np.repeat(mass[None], S, axis=0)
    for s in range(1, S):  # system 0 keeps the shared composition
        radii[s] = np.round(radius * rng.uniform(0.5, 1.1, n) * 1024.0) / 1024.0
        masses[s] = mass * rng.uniform(0.3, 3.0, n)
    if n >= 4:
        radii[:, 3] = 0.0   # the point particle stays one (two of them never collide)
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, radius=radii, mass=masses,
                                     dot_fma=fma)
    T = 9.0
    c1 = ens.evolve(T / 3)
    c2 = ens.evolve(T)
    ens.run(15)
    got_t, got_v, got_pos = ens.time, ens.balls_velocity, ens.balls_position
    got_bb, got_bo, tables = ens.counts_ball_ball, ens.counts_ball_obstacle, ens.toi_tables
    for s in range(S):
        o = oracle.OracleBilliard(workloads.oracle_obstacles(obstacles), dot_fma=fma)
        o.add_balls(pos[s], vel[s], radii[s], masses[s])
        a, b = o.evolve(T / 3), o.evolve(T)
        c = o.evolve(INF, max_events=15)
        assert (int(c1[0][s]), int(c1[1][s])) == a and (int(c2[0][s]), int(c2[1][s])) == b, f"system {s}"
        assert (int(got_bb[s]), int(got_bo[s])) == (a[0] + b[0] + c[0], a[1] + b[1] + c[1]), f"system {s}"
        st = o.state()
        assert got_t[s] == o.time, f"system {s}"
        assert_bits_equal(got_v[s], st["v"], f"system {s}: velocities")
        assert_bits_equal(got_pos[s], st["pos"], f"system {s}: positions")
        rows = o.toi_table
        for i in range(n):
            assert_bits_equal(tables[s, i, :i], rows[i], f"system {s}: table row {i}")
    shared = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos[:1], vel[:1], radius=radii[0],
                                        mass=masses[0], dot_fma=fma)
    shared.evolve(T / 3); shared.evolve(T); shared.run(15)
    assert_bits_equal(shared.balls_velocity[0], got_v[0], "shared composition == per-system composition")


def test_galperin_digits_of_pi_at_six_mass_ratios_in_one_launch(billiards):
    """README.md:58-63, :114-135 of the reference (the number of collisions spells pi) as ONE ensemble: six two-ball
    systems whose heavy ball weighs 100**k, k = 0 .. 5 -- per-system masses."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    cfg = workloads.galperin()
    S = 6
    pos = np.repeat(cfg.pos[None], S, axis=0)
    vel = np.repeat(cfg.vel[None], S, axis=0)
    mass = np.asarray([[1.0, 100.0 ** k] for k in range(S)])
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, cfg.obstacles), pos, vel,
                                     radius=np.repeat(cfg.radius[None], S, axis=0), mass=mass, dot_fma=fma)
    bb, bo = ens.evolve(INF, max_collisions=10**6)
    assert (bb + bo).tolist() == [3, 31, 314, 3141, 31415, 314159]
    single = workloads.galperin().build(billiards, dot_fma=fma)
    assert single.evolve(16) == (int(bb[5]), int(bo[5]))
