"""Parity of the CUDA path (through the C ABI, via the drop-in Python API) with the golden vectors of the live
reference and with the CPU oracle on the same seeded inputs.  Bit-exact for event order, times and state."""

import numpy as np
import pytest

import workloads
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu
INF = float("inf")


@pytest.fixture(scope="module")
def billiards():
    import billiards as pkg
    return pkg


def build(billiards, cfg, fma, one_by_one=False):
    bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), dot_fma=fma)
    if one_by_one:
        for k in range(cfg.n):
            bld.add_ball(tuple(cfg.pos[k]), tuple(cfg.vel[k]), float(cfg.radius[k]), float(cfg.mass[k]))
    else:
        bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    return bld


LOGGED = [
    ("pool", workloads.pool, dict(end_time=100.0)),
    ("hetero_gas", workloads.hetero_gas, dict(n_events=4000)),
    ("dense_gas_256", lambda: workloads.dense_gas(16), dict(n_events=2000)),
    ("dense_gas_1024", lambda: workloads.dense_gas(32), dict(n_events=1000)),
    ("ideal_gas_1024", lambda: workloads.ideal_gas(32), dict(n_events=1000)),
    ("dense_gas_16384", lambda: workloads.dense_gas(128), dict(n_events=100)),
    ("maze", workloads.maze, dict(n_events=3000)),            # LineSegment obstacles (examples/maze.py)
    ("maze_wide", workloads.maze_wide, dict(n_events=4000)),  # all four segments, both faces, end points
    ("collapse", workloads.collapse, dict(end_time=15.0)),    # no obstacles at all (examples/collapse.py)
]


@pytest.mark.parametrize("name,make,kw", LOGGED, ids=[c[0] for c in LOGGED])
def test_event_log_matches_reference(billiards, golden, name, make, kw):
    _check_against_golden_log(billiards, golden, name, make, kw)


MID = [c for c in LOGGED if c[0] in ("hetero_gas", "maze", "maze_wide", "collapse")]  # 60, 200, 150, 200 balls


@pytest.mark.parametrize("name,make,kw", MID, ids=[c[0] for c in MID])
def test_event_log_through_the_one_cta_kernel(billiards, golden, name, make, kw):
    """Single systems of 33 .. 224 balls take the grid-wide kernel by default; the one-CTA-per-system kernel that serves
    ensembles of such systems (csrc/bl_evolve_cta.cu) must reproduce the live reference's logs, end states, row minima
    and whole tables just the same."""
    from billiards import _lib
    h = _lib.get_handle(None, bool(golden(name)["dot_fma"]))
    h.check(h.lib.bl_set_small_kernel(h.h, 3))
    try:
        _check_against_golden_log(billiards, golden, name, make, kw)
    finally:
        h.check(h.lib.bl_set_small_kernel(h.h, 1))


def _check_against_golden_log(billiards, golden, name, make, kw):
    g = golden(name)
    fma = bool(g["dot_fma"])
    bld = build(billiards, make(), fma)
    # table state right after the build
    assert_bits_equal(bld._balls_toi, g["start_balls_toi"], "initial _balls_toi")
    assert [int(x) for x in bld._balls_idx] == g["start_balls_idx"].tolist()
    assert_bits_equal(bld._obstacles_toi, g["start_obs_toi"], "initial _obstacles_toi")
    if "start_toi_tri" in g:
        assert_bits_equal(np.concatenate(bld.toi_table), g["start_toi_tri"], "initial toi_table")
    if "end_time" in kw:
        counts, (t, i, j) = bld.evolve_events(10**9, end_time=kw["end_time"], log_capacity=len(g["log_t"]) + 8)
    else:
        counts, (t, i, j) = bld.evolve_events(kw["n_events"], log_capacity=kw["n_events"])
    assert len(t) == len(g["log_t"])
    assert i.tolist() == g["log_i"].tolist()
    assert j.tolist() == g["log_j"].tolist()
    assert_bits_equal(t, g["log_t"], "event times")
    assert counts == (int((g["log_j"] >= 0).sum()), int((g["log_j"] < 0).sum()))
    if "end_time" in kw:
        assert bld.time == float(g["end_time"])
        assert_bits_equal(bld.balls_position, g["end_pos"], "end positions")
    assert_bits_equal(bld.balls_velocity, g["end_vel"], "end velocities")
    assert_bits_equal(bld.balls_initial_position, g["end_p0"], "end initial positions")
    assert_bits_equal(bld.balls_initial_time[:, 0], g["end_t0"], "end initial times")
    assert_bits_equal(bld._balls_toi, g["end_balls_toi"], "_balls_toi")
    assert [int(x) for x in bld._balls_idx] == g["end_balls_idx"].tolist()
    assert_bits_equal(bld._obstacles_toi, g["end_obs_toi"], "_obstacles_toi")
    nb = bld.next_ball_ball_collision
    assert_bits_equal([nb[0]], g["end_next_bb"][:1])
    assert list(nb[1:]) == g["end_next_bb"][1:].astype(int).tolist()
    no = bld.next_ball_obstacle_collision
    assert_bits_equal([no[0]], g["end_next_bo"][:1])
    assert no[1] == int(g["end_next_bo"][1])
    if "end_toi_tri" in g:
        assert_bits_equal(np.concatenate(bld.toi_table), g["end_toi_tri"], "toi_table")


def test_pool_payload_and_one_by_one_build(billiards, golden):
    g = golden("pool")
    bld = build(billiards, workloads.pool(), bool(g["dot_fma"]), one_by_one=True)
    counts, (t, i, j, p) = bld.evolve_events(10**9, end_time=100.0, log_capacity=512, payload=True)
    assert counts == (167, 248)
    assert_bits_equal(t, g["log_t"])
    assert_bits_equal(p, g["log_payload"], "callback payloads")


def test_galperin_exact(billiards, golden):
    """314159 collisions for mass ratio 100**5 (README.md:134-135, docs/quickstart.rst:50-131 of the reference)."""
    g = golden("galperin")
    fma = bool(g["dot_fma"])
    cfg = workloads.galperin()
    bld = build(billiards, cfg, fma, one_by_one=True)
    assert bld.next_collision == (1.8000000000000005, 0, 1)
    total, cum = 0, []
    for t in (1, 2, 3, 4, 5):
        total += sum(bld.evolve(t))
        cum.append(total)
    assert cum == [0, 1, 1, 4, 314152]
    total += sum(bld.evolve(16))
    assert total == 314159
    assert bld.balls_velocity[0, 0] == float.fromhex("0x1.78217eaf374a7p-1")
    assert_bits_equal(bld.balls_position, g["end_pos"])
    assert_bits_equal(bld.balls_velocity, g["end_vel"])
    v, m = bld.balls_velocity, np.asarray(bld.balls_mass)
    assert float((m * (v * v).sum(axis=1)).sum() / 2) == 4999999999.989935
    assert bld.next_ball_ball_collision == (INF, -1, 0)
    assert bld.next_ball_obstacle_collision == (INF, 0, None)
    one = build(billiards, cfg, fma)
    assert one.evolve(16) == (157080, 157079)
    logged = build(billiards, cfg, fma)
    _, (t, i, j) = logged.evolve_events(20000, log_capacity=20000)
    assert_bits_equal(t, g["log_t"])
    assert i.tolist() == g["log_i"].tolist() and j.tolist() == g["log_j"].tolist()


@pytest.mark.parametrize("n_side,n_events", [(8, 3000), (23, 1500), (40, 600)])
def test_against_oracle_seeded(billiards, oracle, n_side, n_events):
    """Sizes and seeds the golden files do not cover (odd N -> partially filled warps, N > 1024 -> clusters)."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    cfg = workloads.dense_gas(n_side, seed=n_side, phi=0.35)
    bld = build(billiards, cfg, fma)
    ref = cfg.build_oracle(oracle, dot_fma=fma, flags=oracle.FAST_ROWMIN)
    ref.enable_log(n_events)
    ref.evolve(INF, max_events=n_events)
    rt, ri, rj = ref.log()
    _, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
    assert i.tolist() == ri.tolist() and j.tolist() == rj.tolist()
    assert_bits_equal(t, rt, "event times")
    st = ref.state()
    assert_bits_equal(bld.balls_velocity, st["v"])
    assert_bits_equal(bld.balls_initial_position, st["p0"])
    bt, bi, ot, oo, _ = ref.minima()
    assert_bits_equal(bld._balls_toi, bt)
    assert [int(x) for x in bld._balls_idx] == bi.tolist()
    assert_bits_equal(np.concatenate(bld.toi_table), np.concatenate(ref.toi_table), "toi_table")


def _oracle_log(oracle, cfg, fma, n_events):
    ref = cfg.build_oracle(oracle, dot_fma=fma, flags=oracle.FAST_ROWMIN)
    ref.enable_log(n_events)
    ref.evolve(INF, max_events=n_events)
    return ref, ref.log()


@pytest.mark.parametrize("n_side,n_events", [(48, 4000), (64, 2500), (128, 6000)])
def test_grid_mode_long_run_against_oracle(billiards, oracle, n_side, n_events):
    """More than 2048 balls: cooperative grid launch, global-memory barrier, windows of dozens of collisions per
    round, several stale-candidate repair phases -- event by event against the CPU oracle."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    cfg = workloads.dense_gas(n_side)
    ref, (rt, ri, rj) = _oracle_log(oracle, cfg, fma, n_events)
    bld = build(billiards, cfg, fma)
    _, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
    assert i.tolist() == ri.tolist() and j.tolist() == rj.tolist()
    assert_bits_equal(t, rt, "event times")
    assert bld.last_result.prof[5] < n_events / 4, "rounds should commit several collisions each"
    st = ref.state()
    assert_bits_equal(bld.balls_velocity, st["v"])
    assert_bits_equal(bld.balls_initial_position, st["p0"])
    assert_bits_equal(bld.balls_initial_time[:, 0], st["t0"])
    bt, bi, ot, oo, _ = ref.minima()
    assert_bits_equal(bld._balls_toi, bt)          # goes through bl_symmetrize
    assert [int(x) for x in bld._balls_idx] == bi.tolist()
    assert_bits_equal(bld._obstacles_toi, ot)
    if n_side <= 48:
        assert_bits_equal(np.concatenate(bld.toi_table), np.concatenate(ref.toi_table), "toi_table")


def test_heterogeneous_system_in_grid_mode(billiards, oracle):
    """~2400 balls with ~250 distinct radii, random masses incl. one infinite and one zero, four walls, a Disk and two
    LineSegments: radius-class lookups, mass handling and every obstacle kind in the cooperative-grid kernel."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    base = workloads.hetero_gas(side=50, seed=17)
    segs = [((10.2, 10.3), (20.4, 14.1)), ((35.5, 8.2), (30.1, 22.7))]
    pos, keep = np.asarray(base.pos), np.ones(base.n, dtype=bool)
    for a, b in segs:
        a, b = np.asarray(a), np.asarray(b)
        u = np.clip(((pos - a) @ (b - a)) / ((b - a) @ (b - a)), 0.0, 1.0)
        keep &= np.hypot(*(pos - (a + u[:, None] * (b - a))).T) > np.asarray(base.radius) + 0.05
    cfg = workloads.Config("hetero_grid", base.obstacles + [("segment", a, b) for a, b in segs], pos[keep],
                           np.asarray(base.vel)[keep], np.asarray(base.radius)[keep], np.asarray(base.mass)[keep])
    assert cfg.n > 2048
    n_events = 4000
    ref, (rt, ri, rj) = _oracle_log(oracle, cfg, fma, n_events)
    assert (rj <= -6).sum() > 5 and (rj == -5).sum() > 0     # the segments and the disk do get hit
    bld = build(billiards, cfg, fma)
    _, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
    assert i.tolist() == ri.tolist() and j.tolist() == rj.tolist()
    assert_bits_equal(t, rt, "event times")
    assert_bits_equal(bld.balls_velocity, ref.state()["v"])
    assert_bits_equal(np.concatenate(bld.toi_table), np.concatenate(ref.toi_table), "toi_table")
    bt, bi, ot, oo, _ = ref.minima()
    assert_bits_equal(bld._obstacles_toi, ot)


@pytest.mark.parametrize("target", [1, 2, 5, 17, 44])
def test_batch_size_does_not_change_results(billiards, oracle, target):
    """Collisions per synchronisation round is a speed knob only: same log, same table, same state for every value."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    for cfg, n_events in ((workloads.hetero_gas(), 3000), (workloads.dense_gas(34, seed=7, phi=0.45), 2500),
                          (workloads.dense_gas(47, seed=3), 2500)):
        ref, (rt, ri, rj) = _oracle_log(oracle, cfg, fma, n_events)
        h.check(h.lib.bl_set_batch_target(h.h, target))
        try:
            bld = build(billiards, cfg, fma)
            _, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
        finally:
            h.check(h.lib.bl_set_batch_target(h.h, 0))
        assert i.tolist() == ri.tolist() and j.tolist() == rj.tolist()
        assert_bits_equal(t, rt, "event times")
        assert_bits_equal(bld.balls_velocity, ref.state()["v"])
        assert_bits_equal(np.concatenate(bld.toi_table), np.concatenate(ref.toi_table), "toi_table")


def test_stops_inside_a_batch(billiards, oracle):
    """Count-based and time-based stops that fall in the middle of a window: many short launches == one launch."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    cfg = workloads.dense_gas(47, seed=11)
    n_events = 3000
    ref, (rt, ri, rj) = _oracle_log(oracle, cfg, fma, n_events)
    bld = build(billiards, cfg, fma)
    ts, is_, js = [], [], []
    chunks = [1, 2, 3, 5, 7, 11, 13, 64, 1, 100, 257]
    done = 0
    while done < n_events:
        k = min(chunks[len(ts) % len(chunks)], n_events - done)
        c, (t, i, j) = bld.evolve_events(k, log_capacity=k)
        assert sum(c) == k
        ts.append(t); is_.append(i); js.append(j)
        done += k
    assert np.concatenate(is_).tolist() == ri.tolist() and np.concatenate(js).tolist() == rj.tolist()
    assert_bits_equal(np.concatenate(ts), rt, "event times")
    assert_bits_equal(bld.balls_velocity, ref.state()["v"])
    # time-based: frames of a fixed dt, as visualize_matplotlib.animate drives evolve
    a, b = build(billiards, cfg, fma), build(billiards, cfg, fma)
    t_end = float(rt[1500])
    total = (0, 0)
    for f in range(1, 201):
        c = a.evolve(t_end * f / 200)
        total = (total[0] + c[0], total[1] + c[1])
    assert total == b.evolve(t_end)
    assert_bits_equal(a.balls_position, b.balls_position)
    assert_bits_equal(a.balls_velocity, b.balls_velocity)
    assert_bits_equal(np.concatenate(a.toi_table), np.concatenate(b.toi_table), "toi_table")


def test_cta_shape_does_not_change_results(billiards, oracle):
    """Systems of 65..3600 balls run 32 balls x 16 threads per CTA, larger ones 64 x 8: a speed choice only.  Both
    shapes (one radius class and many) give the same log, state and table; the heterogeneous one also equals the
    oracle event by event."""
    from billiards import _lib
    import ctypes
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)

    def shape(n):
        c, t = ctypes.c_int(), ctypes.c_int()
        h.check(h.lib.bl_evolve_config(h.h, n, ctypes.byref(c), ctypes.byref(t)))
        return c.value, t.value

    hetero = workloads.hetero_gas(side=14, seed=11)
    ref, (rt, ri, rj) = _oracle_log(oracle, hetero, fma, 3000)
    for cfg, n_events in ((workloads.dense_gas(40, seed=5), 20000), (hetero, 3000), (workloads.dense_gas(9), 4000)):
        runs = []
        for limit in (-1, 0):
            h.check(h.lib.bl_set_cta_shape_limit(h.h, limit))
            try:
                ctas = shape(cfg.n)[0]
                bld = build(billiards, cfg, fma)
                counts, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
                runs.append((ctas, counts, t, i, j, bld.balls_velocity.copy(), bld.balls_initial_position.copy(),
                             np.concatenate(bld.toi_table)))
            finally:
                h.check(h.lib.bl_set_cta_shape_limit(h.h, -1))
        a, b = runs
        assert a[0] == (cfg.n + 31) // 32 and b[0] == (cfg.n + 63) // 64
        assert a[1] == b[1] and sum(a[1]) == n_events
        assert a[3].tolist() == b[3].tolist() and a[4].tolist() == b[4].tolist()
        assert_bits_equal(a[2], b[2], "event times")
        assert_bits_equal(a[5], b[5], "velocities")
        assert_bits_equal(a[6], b[6], "initial positions")
        assert_bits_equal(a[7], b[7], "toi_table")
        if cfg is hetero:
            assert a[3].tolist() == ri.tolist() and a[4].tolist() == rj.tolist()
            assert_bits_equal(a[2], rt, "event times vs oracle")


def test_batched_rounds_equal_one_at_a_time_over_a_long_run(billiards):
    """120 000 collisions of the 16384-ball gas: windows of ~24 collisions per round (with their rollbacks and repair
    phases) against the same kernel committing ONE collision per round, which is the reference's schedule.  Far beyond
    what the CPU oracle can reach; event log, state and table must be identical bit for bit."""
    from billiards import _lib
    import hashlib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg = workloads.dense_gas(128)
    n_events = 120000
    runs = []
    for target in (0, 1):
        h.check(h.lib.bl_set_batch_target(h.h, target))
        try:
            bld = build(billiards, cfg, fma)
            counts, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
        finally:
            h.check(h.lib.bl_set_batch_target(h.h, 0))
        rounds = int(bld.last_result.prof[5])
        digest = hashlib.sha256(t.tobytes() + i.tobytes() + j.tobytes()).hexdigest()
        runs.append((counts, digest, bld.balls_velocity.copy(), bld.balls_initial_position.copy(),
                     bld._balls_toi.copy(), bld.time, rounds))
    assert runs[0][6] < n_events / 8 and runs[1][6] >= n_events      # batched vs one collision per round
    assert runs[0][0] == runs[1][0] and sum(runs[0][0]) == n_events
    assert runs[0][1] == runs[1][1], "event logs differ"
    assert_bits_equal(runs[0][2], runs[1][2], "velocities")
    assert_bits_equal(runs[0][3], runs[1][3], "initial positions")
    assert_bits_equal(runs[0][4], runs[1][4], "_balls_toi")
    assert runs[0][5] == runs[1][5]
    # physics sanity of the long run: energy conserved (equal masses, elastic), nobody left the box
    v = runs[0][2]
    assert abs(float((v * v).sum()) / cfg.n - 1.0) < 1e-9
    p = runs[0][3]
    r = float(cfg.radius[0])
    assert p.min() >= r - 1e-9 and p.max() <= 128.0 - r + 1e-9


def test_short_rounds_do_not_race_with_slower_ctas(billiards):
    """One collision per round makes the rounds so short that a CTA reaches its commit (which overwrites the global
    mirrors of the colliding balls) while a slower CTA may still be resolving from them -- unless the readers' gate
    holds it back.  Repeated runs of that schedule, with launches of another shape in between, must all give the
    log of the batched schedule."""
    from billiards import _lib
    import hashlib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg, other = workloads.dense_gas(128), workloads.dense_gas(47, seed=11)
    n_events = 40000

    def run(target):
        h.check(h.lib.bl_set_batch_target(h.h, target))
        try:
            bld = build(billiards, cfg, fma)
            counts, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
        finally:
            h.check(h.lib.bl_set_batch_target(h.h, 0))
        return counts, hashlib.sha256(t.tobytes() + i.tobytes() + j.tobytes()).hexdigest(), bld.balls_velocity.copy()

    want = run(0)
    for rep in range(8):
        o = build(billiards, other, fma)
        for k in (1, 2, 3, 5, 64, 257, 11):
            o.evolve_events(k, log_capacity=k)
        got = run(1)
        assert got[0] == want[0] and got[1] == want[1], f"repetition {rep}: one collision per round differs"
        assert_bits_equal(got[2], want[2], "velocities")


def test_beyond_16384_balls(billiards, oracle):
    """32761 balls: more balls than 148 CTAs x 128, so 224 balls per CTA with two helpers each (the run-time shaped
    kernel variant) and an 8.6 GB table.  First 600 collisions against the oracle, then 30 000 batched vs one per round."""
    from billiards import _lib
    import hashlib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg = workloads.dense_gas(181, seed=4, phi=0.45)
    ref, (rt, ri, rj) = _oracle_log(oracle, cfg, fma, 600)
    digests = []
    for target in (0, 1):
        h.check(h.lib.bl_set_batch_target(h.h, target))
        try:
            bld = build(billiards, cfg, fma)
            _, (t, i, j) = bld.evolve_events(600, log_capacity=600)
            assert i.tolist() == ri.tolist() and j.tolist() == rj.tolist()
            assert_bits_equal(t, rt, "event times")
            _, (t, i, j) = bld.evolve_events(30000, log_capacity=30000)
        finally:
            h.check(h.lib.bl_set_batch_target(h.h, 0))
        digests.append((hashlib.sha256(t.tobytes() + i.tobytes() + j.tobytes()).hexdigest(),
                        bld.balls_velocity.copy(), int(bld.last_result.prof[5])))
        del bld
    assert digests[0][0] == digests[1][0]
    assert_bits_equal(digests[0][1], digests[1][1])
    assert digests[0][2] < 30000 / 8 <= digests[1][2]


@pytest.mark.parametrize("side,events", [(190, 4000), (256, 3000)], ids=["36100", "65536"])
def test_largest_systems_trade_threads_and_batch_size_for_balls(billiards, side, events):
    """Beyond 148 x 224 balls a CTA's shared memory no longer holds two helper threads per ball and full-size batches:
    36100 balls run 256 per CTA with one thread each, 65536 balls (a 34 GB table) 448 per CTA and at most 4 collisions per
    round.  Out of the oracle's reach, so: batched rounds vs one collision per round (the reference's schedule) must
    give the same log and the same bits, the energy is conserved and everybody stays in the box."""
    from billiards import _lib
    import hashlib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg = workloads.dense_gas(side, seed=5, phi=0.4)
    digests = []
    for target in (0, 1):
        h.check(h.lib.bl_set_batch_target(h.h, target))
        try:
            bld = build(billiards, cfg, fma)
            _, (t, i, j) = bld.evolve_events(events, log_capacity=events)
        finally:
            h.check(h.lib.bl_set_batch_target(h.h, 0))
        assert len(t) == events and np.all(np.diff(t) >= -1e-10)
        vel, pos = bld.balls_velocity.copy(), bld.balls_position.copy()
        digests.append((hashlib.sha256(t.tobytes() + i.tobytes() + j.tobytes()).hexdigest(), vel, pos,
                        int(bld.last_result.prof[5])))
        del bld
    assert digests[0][0] == digests[1][0]
    assert_bits_equal(digests[0][1], digests[1][1], "velocities")
    assert_bits_equal(digests[0][2], digests[1][2], "positions")
    assert 2 * digests[0][3] < events <= digests[1][3]  # batches really were batches; one per round really was one
    r = float(cfg.radius[0])
    assert abs((digests[0][1] ** 2).sum() - (cfg.vel ** 2).sum()) < 1e-9 * (cfg.vel ** 2).sum()
    assert digests[0][2].min() >= r * (1 - 1e-9) and digests[0][2].max() <= side - r * (1 - 1e-9)


def test_small_systems_through_the_grid_kernel(billiards, golden):
    """Systems of <= 32 balls normally run in the one-warp kernel; the grid-wide kernel must give the same bits."""
    from billiards import _lib
    g = golden("pool")
    fma = bool(g["dot_fma"])
    h = _lib.get_handle(None, fma)
    h.check(h.lib.bl_set_small_kernel(h.h, 0))
    try:
        bld = build(billiards, workloads.pool(), fma)
        counts, (t, i, j, p) = bld.evolve_events(10**9, end_time=100.0, log_capacity=512, payload=True)
        assert counts == (167, 248)
        assert_bits_equal(t, g["log_t"])
        assert i.tolist() == g["log_i"].tolist() and j.tolist() == g["log_j"].tolist()
        assert_bits_equal(p, g["log_payload"], "callback payloads")
        assert_bits_equal(bld.balls_position, g["end_pos"])
        assert_bits_equal(np.concatenate(bld.toi_table), g["end_toi_tri"], "toi_table")
        gg = golden("galperin")
        gal = build(billiards, workloads.galperin(), fma)
        _, (t, i, j) = gal.evolve_events(20000, log_capacity=20000)
        assert_bits_equal(t, gg["log_t"])
        assert i.tolist() == gg["log_i"].tolist() and j.tolist() == gg["log_j"].tolist()
        assert gal.evolve(16) == (157080 - int((gg["log_j"] >= 0).sum()), 157079 - int((gg["log_j"] < 0).sum()))
    finally:
        h.check(h.lib.bl_set_small_kernel(h.h, 1))


def test_tiny_systems_through_the_lane_group_kernel(billiards, golden):
    """Systems of <= 4 balls normally run one thread per system (bl_evolve_tiny.cu); the lane-group kernel with 2 / 4
    lanes per system must give the same bits: Galperin's log and total, Newton's cradle with 4 balls."""
    from billiards import _lib
    gg = golden("galperin")
    fma = bool(gg["dot_fma"])
    h = _lib.get_handle(None, fma)
    runs = {}
    for mode in (1, 2):
        h.check(h.lib.bl_set_small_kernel(h.h, mode))
        try:
            gal = build(billiards, workloads.galperin(), fma)
            _, (t, i, j, p) = gal.evolve_events(20000, log_capacity=20000, payload=True)
            assert_bits_equal(t, gg["log_t"])
            assert i.tolist() == gg["log_i"].tolist() and j.tolist() == gg["log_j"].tolist()
            assert gal.evolve(16) == (157080 - int((gg["log_j"] >= 0).sum()), 157079 - int((gg["log_j"] < 0).sum()))
            assert gal.next_ball_ball_collision == (INF, -1, 0) and gal.next_ball_obstacle_collision == (INF, 0, None)
            cr = build(billiards, workloads.newtons_cradle(4, True), fma)
            c = cr.evolve(40.0, time_callback=lambda t: None)
            runs[mode] = (p, gal.balls_velocity.copy(), gal.balls_position.copy(), c, cr.balls_position.copy(),
                          np.concatenate(cr.toi_table), cr._balls_toi.copy(), cr._obstacles_toi.copy(), cr.next_collision[0])
        finally:
            h.check(h.lib.bl_set_small_kernel(h.h, 1))
    a, b = runs[1], runs[2]
    assert a[3] == b[3] and a[8] == b[8]
    for q in (0, 1, 2, 4, 5, 6, 7):
        assert_bits_equal(a[q], b[q], f"item {q}")


def test_stop_resume_is_bit_stable(billiards, oracle):
    """tests/test_simulation.py:461-502: 300 partial evolves == one evolve, exactly; and == the oracle."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    cfg = workloads.Config("four", workloads.box_walls(-1, -1, 1, 1), [(0, 0.2), (0, 0.5), (0, -0.8), (0.7, -0.8)],
                           [(1, 2), (-1, 2), (0, -1), (0, 0)], 0.2, 1.0)
    once, step = build(billiards, cfg, fma), build(billiards, cfg, fma)
    once.evolve(10)
    for i in range(10):
        for j in range(1, 31):
            step.evolve(i + j / 30)
    assert np.linalg.norm(once.balls_position - step.balls_position, axis=1).max() == 0
    ref = cfg.build_oracle(oracle, dot_fma=fma)
    ref.evolve(10)
    assert_bits_equal(once.balls_position, ref.balls_position)


def test_recompute_scenario(billiards, golden):
    g = golden("four_balls")
    fma = bool(g["dot_fma"])
    cfg = workloads.Config("four", workloads.box_walls(-1, -1, 1, 1), [(0, 0.2), (0, 0.5), (0, -0.8), (0.7, -0.8)],
                           [(1, 2), (-1, 2), (0, -1), (0, 0)], 0.2, 1.0)

    def check(b, tag):
        assert_bits_equal(np.concatenate(b.toi_table), g[tag + "_toi_tri"], tag + " table")
        assert_bits_equal(b._balls_toi, g[tag + "_balls_toi"], tag)
        assert [int(x) for x in b._balls_idx] == g[tag + "_balls_idx"].tolist()
        assert_bits_equal(b._obstacles_toi, g[tag + "_obs_toi"], tag)
        assert_bits_equal(b.balls_position, g[tag + "_pos"], tag)
        assert_bits_equal(b.balls_initial_time[:, 0], g[tag + "_t0"], tag)

    b = build(billiards, cfg, fma, one_by_one=True)
    b.evolve(1)
    b.balls_position[0] = (0, 0)
    b.recompute_toi(0)
    assert b.time == 1
    check(b, "s1")
    b.evolve(2)
    b.balls_velocity[1] = (0, 0)
    b.balls_radius[2] = 0.1
    b.recompute_toi([1, 2])
    check(b, "s2")
    b.evolve(3)
    b.recompute_toi()
    check(b, "s3")
    assert b.evolve(5) == tuple(g["counts5"].tolist())
    check(b, "s5")


@pytest.mark.parametrize("n_side,grid_kernel_small", [(12, False), (48, False), (4, True)])
def test_edits_and_new_balls_between_launches(billiards, oracle, n_side, grid_kernel_small):
    """recompute_toi after in-place edits and add_ball on a system that has already evolved -- through the cluster
    (144 balls), the grid (2304 balls) and, with the one-warp kernel switched off, 16 balls in the grid-wide kernel:
    the table mirror refresh (bl_symmetrize), the candidate rebuild and the sequence numbers have to line up."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg = workloads.dense_gas(n_side, seed=5, phi=0.3)
    n1 = 900 if n_side > 4 else 120
    if grid_kernel_small:
        h.check(h.lib.bl_set_small_kernel(h.h, 0))
    try:
        bld = build(billiards, cfg, fma)
        ref = cfg.build_oracle(oracle, dot_fma=fma, flags=oracle.FAST_ROWMIN)

        def step(k):
            ref.clear_log()
            ref.enable_log(k)
            ref.evolve(INF, max_events=k)
            rt, ri, rj = ref.log()
            _, (t, i, j) = bld.evolve_events(k, log_capacity=k)
            assert i.tolist() == ri.tolist() and j.tolist() == rj.tolist()
            assert_bits_equal(t, rt, "event times")

        def same_tables():
            assert_bits_equal(np.concatenate(bld.toi_table), np.concatenate(ref.toi_table), "toi_table")
            bt, bi, ot, oo, _ = ref.minima()
            assert_bits_equal(bld._balls_toi, bt)
            assert [int(x) for x in bld._balls_idx] == bi.tolist()
            assert_bits_equal(bld._obstacles_toi, ot)

        step(n1)
        # in-place velocity edit of one ball + recompute_toi(index), simulation.py:226-309
        v = bld.balls_velocity[5].copy()
        bld.balls_velocity[5] = (-v[1], v[0])
        bld.recompute_toi(5)
        ref.set_ball(5, vel=(-v[1], v[0]))
        ref.recompute_toi(5)
        same_tables()
        step(n1)
        # a new ball in the middle of a run (no table sync in between: add_ball must do it)
        side = float(n_side)
        bld.add_ball((side / 2 + 0.013, side / 2 + 0.021), (0.3, -0.4), radius=0.0, mass=2.0)
        ref.add_ball((side / 2 + 0.013, side / 2 + 0.021), (0.3, -0.4), radius=0.0, mass=2.0)
        step(n1)
        same_tables()
        # recompute everything, then go on
        bld.recompute_toi()
        ref.recompute_toi()
        same_tables()
        step(n1)
        assert_bits_equal(bld.balls_velocity, ref.state()["v"])
        assert_bits_equal(bld.balls_initial_position, ref.state()["p0"])
    finally:
        h.check(h.lib.bl_set_small_kernel(h.h, 1))


def test_cradle_nudge(billiards, golden):
    g = golden("cradle")
    b = build(billiards, workloads.newtons_cradle(5, True), bool(g["dot_fma"]), one_by_one=True)
    assert b.next_ball_ball_collision == tuple(g["next_bb"].tolist())
    _, (t, i, j) = b.evolve_events(10**6, end_time=4.0, log_capacity=64)
    assert_bits_equal(t, g["log_t"])
    assert (i.tolist(), j.tolist()) == (g["log_i"].tolist(), g["log_j"].tolist())
    b.balls_position[2] += (0, 1e-10)
    b.recompute_toi(2)
    assert_bits_equal(np.concatenate(b.toi_table), g["nudge_toi_tri"])
    assert b.evolve(30) == tuple(g["counts30"].tolist())
    assert_bits_equal(b.balls_position, g["end_pos"])


def test_sinai_ensemble(billiards, golden, oracle):
    g = golden("sinai")
    fma = bool(g["dot_fma"])
    obstacles, pos, vel = workloads.sinai(64, seed=0)
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, dot_fma=fma,
                                     count_hits=True)
    ens.run(int(g["n_coll"]))
    assert_bits_equal(ens.time, g["t"])
    assert_bits_equal(ens.balls_initial_position, g["p0"])
    assert_bits_equal(ens.balls_velocity, g["v"])
    assert ens.hits.tolist() == g["hits"].tolist()
    assert ens.gather_stats().tolist() == [64 * int(g["n_coll"])] + g["hits"].sum(axis=0).tolist()
    # resumability: 2 x 1000 == 1 x 2000, and a bigger seeded batch against the oracle loop
    two = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, dot_fma=fma)
    two.run(1000)
    two.run(1000)
    assert_bits_equal(two.time, g["t"])
    obstacles, pos, vel = workloads.sinai(20000, seed=5)
    big = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, dot_fma=fma)
    big.run(300)
    want = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos, vel, 300, dot_fma=fma)
    assert_bits_equal(big.time, want["t"])
    assert_bits_equal(big.balls_velocity, want["v"])
    assert_bits_equal(big.balls_initial_position, want["p0"])
    # the same single system through Billiard
    one = billiards.Billiard(workloads.make_obstacles(billiards, obstacles), dot_fma=fma)
    one.add_ball(pos[7], vel[7], radius=0)
    assert one.evolve_events(300) == (0, 300)
    assert one.time == want["t"][7]


def test_ensemble_kernel_variants(billiards, oracle):
    """Obstacle lists that take the generic kernel (disk first, two disks, nine walls), a radius > 0, and more than
    65535 collisions per launch (the packed per-obstacle hit counters are flushed in between)."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    walls = workloads.box_walls(-1, -1, 1, 1)
    lists = {
        "disk_first": [("disk", (0.0, 0.0), 0.5)] + walls,
        "two_disks": walls + [("disk", (-0.4, 0.1), 0.2), ("disk", (0.45, -0.3), 0.25)],
        "walls_only": walls,
        "octagon": walls + [("wall", (1.3, 0.0), (0.0, 1.3), "left"), ("wall", (0.0, 1.3), (-1.3, 0.0), "left"),
                            ("wall", (-1.3, 0.0), (0.0, -1.3), "left"), ("wall", (0.0, -1.3), (1.3, 0.0), "left"),
                            ("disk", (0.0, 0.0), 0.3)],
    }
    rng = np.random.default_rng(11)
    for name, obstacles in lists.items():
        n = 3000
        pos = rng.uniform(-0.95, 0.95, (n, 2))
        ok = np.ones(n, dtype=bool)
        for ob in obstacles:
            if ob[0] == "disk":
                ok &= ((pos - np.asarray(ob[1])) ** 2).sum(axis=1) > (ob[2] + 0.03) ** 2
        pos = pos[ok]
        theta = rng.uniform(0, 2 * np.pi, len(pos))
        vel = np.stack([np.cos(theta), np.sin(theta)], axis=1)
        radius = 0.0 if name != "two_disks" else 0.02
        ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, radius=radius,
                                         dot_fma=fma, count_hits=True)
        ens.run(500)
        want = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos, vel, 500, radius=radius, dot_fma=fma,
                                   count_hits=True)
        assert_bits_equal(ens.time, want["t"], name)
        assert_bits_equal(ens.balls_velocity, want["v"], name)
        assert_bits_equal(ens.balls_initial_position, want["p0"], name)
        assert ens.hits.tolist() == want["hits"].tolist(), name
    obstacles, pos, vel = workloads.sinai(96, seed=2)
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, obstacles), pos, vel, dot_fma=fma, count_hits=True)
    ens.run(70000)
    want = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos, vel, 70000, dot_fma=fma, count_hits=True)
    assert_bits_equal(ens.time, want["t"])
    assert_bits_equal(ens.balls_velocity, want["v"])
    assert ens.hits.tolist() == want["hits"].tolist()
    assert int(ens.hits.sum()) == 96 * 70000


def test_full_size_ideal_gas_two_schedules(billiards):
    """Config I at its full size (BASELINE.json: N = 1024, 10^6 collisions), far beyond the oracle's reach: the default
    schedule (32 x 16 CTAs, ~22 keys per round) against another decomposition of the same work (64 x 8 CTAs, 7 keys
    per round) -- log digest, state and row minima bit for bit -- plus what the physics conserves."""
    from billiards import _lib
    import hashlib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg = workloads.ideal_gas()
    n_events = 10**6
    runs = []
    for limit, target in ((-1, 0), (0, 7)):
        h.check(h.lib.bl_set_cta_shape_limit(h.h, limit))
        h.check(h.lib.bl_set_batch_target(h.h, target))
        try:
            bld = build(billiards, cfg, fma)
            counts, (t, i, j) = bld.evolve_events(n_events, log_capacity=n_events)
        finally:
            h.check(h.lib.bl_set_cta_shape_limit(h.h, -1))
            h.check(h.lib.bl_set_batch_target(h.h, 0))
        assert sum(counts) == n_events and (np.diff(t) >= -1e-10).all()
        digest = hashlib.sha256(t.tobytes() + i.tobytes() + j.tobytes()).hexdigest()
        runs.append((counts, digest, bld.balls_velocity.copy(), bld.balls_position.copy(), bld._balls_toi.copy(),
                     bld.time))
    a, b = runs
    assert a[0] == b[0] and a[1] == b[1] and a[5] == b[5]
    assert_bits_equal(a[2], b[2], "velocities")
    assert_bits_equal(a[3], b[3], "positions")
    assert_bits_equal(a[4], b[4], "_balls_toi")
    v, p, r = a[2], a[3], float(cfg.radius[0])
    assert abs(float((v * v).sum()) / (cfg.n * 0.04) - 1.0) < 1e-9       # equal masses: sum |v|^2 is conserved
    assert p.min() >= -1.0 + r - 1e-9 and p.max() <= 1.0 - r + 1e-9
    d = p[:, None, :] - p[None, :, :]
    gap = np.sqrt((d * d).sum(axis=2)) + 4.0 * np.eye(cfg.n)
    assert gap.min() >= 2.0 * r - 1e-9, "overlapping balls"


def test_full_size_dense_gas_two_schedules(billiards):
    """Config D at its full size (BASELINE.json: N = 16384 at packing fraction 0.5, 10^7 collisions) -- what bench.py
    times.  Two decompositions of the same work must agree bit for bit: the default schedule (~30 collisions per
    grid-wide round, launches of 10^6 collisions) against 9 keys per round in launches of 777 777 -- digest of the whole
    event log, end state, reference-shaped row minima -- plus what the physics conserves on the state that is left."""
    from billiards import _lib
    import hashlib
    import torch
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)
    cfg = workloads.dense_gas(128)
    n_events = 10**7
    runs = []
    for target, chunk in ((0, 10**6), (9, 777777)):
        h.check(h.lib.bl_set_batch_target(h.h, target))
        try:
            bld = build(billiards, cfg, fma)
            sha = [hashlib.sha256() for _ in range(3)]  # one stream each for t, i, j: launch boundaries must not matter
            done, last_t, monotone = [0, 0], -INF, True
            while sum(done) < n_events:
                k = min(chunk, n_events - sum(done))
                counts, (t, i, j) = bld.evolve_events(k, log_capacity=k)
                assert sum(counts) == k == len(t)
                done[0] += counts[0]
                done[1] += counts[1]
                monotone = monotone and bool((np.diff(t) >= -1e-10).all()) and t[0] >= last_t - 1e-10
                last_t = float(t[-1])
                for q, x in enumerate((t, i, j)):
                    sha[q].update(x.tobytes())
        finally:
            h.check(h.lib.bl_set_batch_target(h.h, 0))
        assert monotone
        runs.append((tuple(done), tuple(x.hexdigest() for x in sha), bld.balls_velocity.copy(), bld.balls_position.copy(),
                     bld._balls_toi.copy(), bld.time, bld._obstacles_toi.copy()))
    a, b = runs
    assert a[0] == b[0] and a[1] == b[1] and a[5] == b[5]
    assert_bits_equal(a[2], b[2], "velocities")
    assert_bits_equal(a[3], b[3], "positions")
    assert_bits_equal(a[4], b[4], "_balls_toi")
    assert_bits_equal(a[6], b[6], "_obstacles_toi")
    v, p, r = a[2], a[3], float(cfg.radius[0])
    assert abs(float((v * v).sum()) / cfg.n - 1.0) < 1e-9       # unit speeds, equal masses: sum |v|^2 is conserved
    assert p.min() >= r - 1e-9 and p.max() <= 128.0 - r + 1e-9
    pd = torch.from_numpy(p).cuda()
    for lo in range(0, cfg.n, 2048):
        d = torch.cdist(pd[lo:lo + 2048], pd, compute_mode="donot_use_mm_for_euclid_dist")
        d[torch.arange(d.shape[0]), torch.arange(lo, lo + d.shape[0])] = INF
        assert float(d.min().item()) >= 2.0 * r - 1e-9, "overlapping balls"


def test_full_size_sinai_ensemble_is_partition_invariant(billiards):
    """Config S at its full size (2^20 systems x 10^4 collisions): one launch over all systems == two contiguous
    shards (what two ranks would run) == the same shard in two launches of 5000; speeds stay 1, every ball stays
    inside the box and outside the disk."""
    from billiards import _lib
    from billiards.ensemble import shard_bounds
    fma = _lib.probe_dot_fma()
    n, n_coll = 1 << 20, 10**4
    obstacles, pos, vel = workloads.sinai(n, seed=1000)
    obs = workloads.make_obstacles(billiards, obstacles)
    whole = billiards.BilliardEnsemble(obs, pos, vel, dot_fma=fma, count_hits=True)
    whole.run(n_coll)
    t, p0, v = whole.time, whole.balls_initial_position, whole.balls_velocity
    assert (whole.counts == n_coll).all()
    stats = whole.gather_stats()
    assert int(stats[0]) == n * n_coll == int(stats[1:].sum())
    for rank in (0, 1):
        lo, hi = shard_bounds(n, 2, rank)
        part = billiards.BilliardEnsemble(obs, pos[lo:hi], vel[lo:hi], dot_fma=fma)
        if rank == 0:
            part.run(n_coll // 2)
            part.run(n_coll - n_coll // 2)
        else:
            part.run(n_coll)
        assert_bits_equal(part.time, t[lo:hi], f"shard {rank}: times")
        assert_bits_equal(part.balls_initial_position, p0[lo:hi], f"shard {rank}: positions")
        assert_bits_equal(part.balls_velocity, v[lo:hi], f"shard {rank}: velocities")
    assert np.abs((v * v).sum(axis=1) - 1.0).max() < 1e-9
    assert np.abs(p0).max() <= 1.0 + 1e-9 and (p0 * p0).sum(axis=1).min() >= 0.25 - 1e-9
    assert (t > 0).all() and np.isfinite(t).all()


def test_line_segments(billiards, golden, oracle):
    """f3: LineSegment.  Callback payloads (incl. the line parameter u) of the wide maze against the live reference;
    3000 random detect / resolve problems; next_ball_obstacle_collision's args; a segment ensemble against the oracle."""
    g = golden("maze_wide")
    fma = bool(g["dot_fma"])
    bld = build(billiards, workloads.maze_wide(), fma)
    _, (t, i, j, p) = bld.evolve_events(4000, log_capacity=4000, payload=True)
    assert_bits_equal(t, g["log_t"])
    assert_bits_equal(p, g["log_payload"], "callback payloads")
    nb = bld.next_ball_obstacle_collision
    assert nb[0] == g["end_next_bo"][0] and nb[1] == int(g["end_next_bo"][1])
    assert bld.obstacles.index(nb[2][0]) == int(g["end_next_bo"][2])
    s = golden("segment_vectors")
    from billiards import _lib
    import torch
    h = _lib.get_handle(None, fma)
    dev = h.torch_device
    n = len(s["rad"])
    # group the problems by segment and run each group as one batch through the C ABI
    keys = [tuple(r) for r in s["seg"]]
    for key in sorted(set(keys)):
        idx = np.asarray([k for k in range(n) if keys[k] == key and s["vel"][k].any()])
        rec = _lib.BlObstacle(_lib.BL_OBS_SEGMENT, 0, *key)
        rows = np.ascontiguousarray(np.concatenate([s["pos"][idx], s["vel"][idx], s["rad"][idx, None]], axis=1))
        d_in = torch.from_numpy(rows).to(dev)
        d_r2 = torch.from_numpy(np.asarray([x ** 2 for x in s["rad"][idx].tolist()])).to(dev)
        d_t = torch.empty(len(idx), dtype=torch.float64, device=dev)
        d_u = torch.empty(len(idx), dtype=torch.float64, device=dev)
        d_c = torch.empty(len(idx), dtype=torch.int32, device=dev)
        h.check(h.lib.bl_obstacle_detect(h.h, rec, _lib.ptr(d_in), _lib.ptr(d_r2), _lib.ptr(d_t), _lib.ptr(d_u),
                                         len(idx), h.stream()))
        tt, uu = d_t.cpu().numpy(), d_u.cpu().numpy()
        assert_bits_equal(tt, s["detect_t"][idx], "LineSegment.detect_collision time")
        fin = np.isfinite(tt)
        assert_bits_equal(uu[fin], s["detect_u"][idx][fin], "LineSegment.detect_collision u")
        h.check(h.lib.bl_toi_and_param_ball_segment(h.h, rec, _lib.ptr(d_in), -1e-10, _lib.ptr(d_t), _lib.ptr(d_u),
                                                    _lib.ptr(d_c), len(idx), h.stream()))
        assert_bits_equal(d_t.cpu().numpy(), s["param_t"][idx], "toi_and_param_ball_segment time")
        assert d_c.cpu().numpy().tolist() == s["param_code"][idx].tolist()
        face = s["param_code"][idx] == 0
        assert_bits_equal(d_u.cpu().numpy()[face], s["param_u"][idx][face], "toi_and_param_ball_segment u")
        # toi_ball_point against the segment's end point
        rows6 = np.ascontiguousarray(np.concatenate([s["pos"][idx], s["vel"][idx],
                                                     np.tile(np.asarray(key[6:8]), (len(idx), 1))], axis=1))
        d_in6 = torch.from_numpy(rows6).to(dev)
        h.check(h.lib.bl_toi_ball_point(h.h, _lib.ptr(d_in6), _lib.ptr(d_r2), -1e-10, _lib.ptr(d_t), len(idx), h.stream()))
        assert_bits_equal(d_t.cpu().numpy(), s["point_t"][idx], "toi_ball_point")
        for u_arg, key_res, col in ((0.37, "resolve_face", 0), (0.0, "resolve_start", 1), (1.0, "resolve_end", 2)):
            d_arg = torch.full((len(idx),), u_arg, dtype=torch.float64, device=dev)
            d_w = torch.empty((len(idx), 2), dtype=torch.float64, device=dev)
            d_st = torch.empty(len(idx), dtype=torch.int32, device=dev)
            h.check(h.lib.bl_obstacle_resolve(h.h, rec, _lib.ptr(d_in), _lib.ptr(d_arg), _lib.ptr(d_w), _lib.ptr(d_st),
                                              len(idx), h.stream()))
            st = d_st.cpu().numpy()
            assert ((st != 0).astype(int)).tolist() == s["resolve_sep"][idx, col].tolist()
            ok = st == 0
            assert_bits_equal(d_w.cpu().numpy()[ok], s[key_res][idx][ok], key_res)
    # a one-ball-per-system ensemble in the maze against the oracle's loop (generic ensemble kernel)
    rng = np.random.default_rng(3)
    cfgm = workloads.maze_wide(num_balls=400, seed=9)
    pos = np.asarray(cfgm.pos)
    theta = rng.uniform(0, 2 * np.pi, len(pos))
    vel = np.stack([np.cos(theta), np.sin(theta)], axis=1)
    ens = billiards.BilliardEnsemble(workloads.make_obstacles(billiards, cfgm.obstacles), pos, vel, radius=0.02,
                                     dot_fma=fma, count_hits=True)
    ens.run(800)
    want = oracle.ensemble_run(workloads.oracle_obstacles(cfgm.obstacles), pos, vel, 800, radius=0.02, dot_fma=fma,
                               count_hits=True)
    assert_bits_equal(ens.time, want["t"])
    assert_bits_equal(ens.balls_velocity, want["v"])
    assert ens.hits.tolist() == want["hits"].tolist()


def test_unit_vectors(billiards, golden):
    from billiards import physics
    g = golden("unit_vectors")
    fma = bool(g["dot_fma"])
    toi = physics.toi_ball_ball_batch(g["p1"], g["v1"], g["r1"], g["p2"], g["v2"], g["r2"], dot_fma=fma)
    assert_bits_equal(toi, g["toi"], "toi_ball_ball")
    w1, w2, st = physics.elastic_collision_batch(g["q1"], g["u1"], g["m1"], g["q2"], g["u2"], g["m2"], dot_fma=fma)
    ok = g["sep"] == 0
    assert ((st == 0) == ok).all() and (st[~ok] == -3).all()
    assert_bits_equal(w1[ok], g["w1"][ok])
    assert_bits_equal(w2[ok], g["w2"][ok])


def test_pow_is_libm_not_square(billiards, oracle):
    """(r1+r2)**2 goes through libm pow like CPython's float.__pow__: a radius sum for
    which pow(s, 2) != s*s on glibc must give the oracle's (= the reference's) bits."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    s = 8.237813583927716
    r1 = 3.0
    r2 = s - r1
    assert r1 + r2 == s
    cfg = workloads.Config("pow", [], [(0.0, 0.0), (20.0, 0.3)], [(1.0, 0.0), (-1.0, 0.0)], [r1, r2], 1.0)
    b = build(billiards, cfg, fma)
    ref = cfg.build_oracle(oracle, dot_fma=fma)
    assert_bits_equal(np.concatenate(b.toi_table), np.concatenate(ref.toi_table))


def test_single_system_entry_points_through_the_one_cta_kernel(billiards):
    """Everything a single Billiard asks of its event kernel -- forced first kinds (bounce_ball_ball /
    bounce_ball_obstacle), next-collision queries, payload logs, callbacks replayed from the log, a log that fills up,
    count and time limits, stop and resume -- on a 49-ball and a 121-ball gas: the one-CTA kernel (default up to 80
    balls, forced above) against the grid-wide kernel, bit for bit."""
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    h = _lib.get_handle(None, fma)

    def session(cfg):
        out = []
        b = build(billiards, cfg, fma)
        def next_plain(bld):  # (the obstacle of an obstacle hit is an object of that build: keep its arguments only)
            nc = bld.next_collision
            return tuple(nc[:2]) + ((nc[2][1],) if isinstance(nc[2], tuple) else (nc[2],))
        out.append(next_plain(b))
        out.append(b.next_ball_ball_collision)
        out.append(b.next_ball_obstacle_collision[:2])
        for _ in range(2):
            b.bounce_ball_obstacle()
            out.append((b.time, b.balls_velocity.copy()))
            b.bounce_ball_ball()
            out.append((b.time, b.balls_velocity.copy()))
        c, (t, i, j, p) = b.evolve_events(300, log_capacity=300, payload=True)
        out += [c, t, i, j, p]
        c, (t, i, j) = b.evolve_events(10**6, end_time=b.time + 0.7, log_capacity=64)   # the log fills up first
        out += [c, t, i, j, b.time]
        seen = []
        cbs = {k: (lambda t_, p_, u_, v_, o_, k=k: seen.append((k, t_, p_.tolist(), u_.tolist(), v_.tolist(),
                                                                  int(o_) if isinstance(o_, (int, np.integer)) else type(o_).__name__)))
               for k in (0, 5, 17, 40)}
        out.append(b.evolve(b.time + 0.5, time_callback=lambda t_: seen.append(("t", t_)), ball_callbacks=cbs))
        out.append(seen)
        out.append(b.evolve(b.time + 0.25))
        out += [b.time, b.balls_position.copy(), b.balls_velocity.copy(), np.concatenate(b.toi_table),
                np.asarray(b._balls_toi), [int(x) for x in b._balls_idx], np.asarray(b._obstacles_toi), next_plain(b)]
        return out

    def same(x, y):
        if isinstance(x, np.ndarray):
            return x.shape == y.shape and np.array_equal(x.view(np.uint64) if x.dtype == np.float64 else x,
                                                         y.view(np.uint64) if y.dtype == np.float64 else y)
        if isinstance(x, (tuple, list)):
            return len(x) == len(y) and all(same(a, b) for a, b in zip(x, y))
        return x == y or (x != x and y != y)

    for cfg in (workloads.dense_gas(7, seed=3), workloads.dense_gas(11, seed=4)):  # 49 and 121 balls
        results = {}
        for mode in (3, 0):
            h.check(h.lib.bl_set_small_kernel(h.h, mode))
            try:
                results[mode] = session(cfg)
            finally:
                h.check(h.lib.bl_set_small_kernel(h.h, 1))
        for q, (x, y) in enumerate(zip(results[3], results[0])):
            assert same(x, y), f"{cfg.name}: item {q} differs between the one-CTA kernel and the grid-wide kernel"
