"""The CPU oracle (oracle/billiards_oracle.c) against the golden vectors recorded from the live
reference (tests/golden/make_golden.py) and the known answers of the reference's own tests/doctests.
CPU only; this is what pins the oracle (parity gate 3 of the task)."""

import numpy as np
import pytest

import workloads
from conftest import assert_bits_equal

INF = float("inf")


def run_logged(o, n_events=None, end_time=None):
    o.enable_log(1 << 20)
    if n_events is not None:
        o.evolve(INF, max_events=n_events)
    else:
        o.evolve(end_time)
    return o.log()


LOGGED = [
    ("pool", workloads.pool, dict(end_time=100.0)),
    ("hetero_gas", workloads.hetero_gas, dict(n_events=4000)),
    ("dense_gas_256", lambda: workloads.dense_gas(16), dict(n_events=2000)),
    ("dense_gas_1024", lambda: workloads.dense_gas(32), dict(n_events=1000)),
    ("ideal_gas_1024", lambda: workloads.ideal_gas(32), dict(n_events=1000)),
    ("maze", workloads.maze, dict(n_events=3000)),            # LineSegment obstacles, examples/maze.py
    ("maze_wide", workloads.maze_wide, dict(n_events=4000)),  # all four segments, both faces, end points
    ("collapse", workloads.collapse, dict(end_time=15.0)),    # no obstacles at all (examples/collapse.py)
]


@pytest.mark.parametrize("flags", [0, 1], ids=["faithful", "fast_rowmin"])
@pytest.mark.parametrize("name,make,kw", LOGGED, ids=[c[0] for c in LOGGED])
def test_event_log_bit_exact(oracle, golden, name, make, kw, flags):
    g = golden(name)
    o = make().build_oracle(oracle, dot_fma=bool(g["dot_fma"]), flags=flags)
    bt, bi, ot, oo, _ = o.minima()
    assert_bits_equal(bt, g["start_balls_toi"], "initial _balls_toi")
    assert bi.tolist() == g["start_balls_idx"].tolist()
    assert_bits_equal(ot, g["start_obs_toi"], "initial _obstacles_toi")
    assert oo.tolist() == g["start_obs_obs"].tolist()
    t, i, j = run_logged(o, **kw)
    assert len(t) == len(g["log_t"])
    assert i.tolist() == g["log_i"].tolist()
    assert j.tolist() == g["log_j"].tolist()
    assert_bits_equal(t, g["log_t"], "event times")
    if "end_time" in kw:
        assert o.time == float(g["end_time"])
        assert_bits_equal(o.balls_position, g["end_pos"], "end positions")
    st = o.state()
    assert_bits_equal(st["v"], g["end_vel"], "end velocities")
    assert_bits_equal(st["p0"], g["end_p0"], "end initial positions")
    assert_bits_equal(st["t0"], g["end_t0"], "end initial times")
    bt, bi, ot, oo, _ = o.minima()
    assert_bits_equal(bt, g["end_balls_toi"], "_balls_toi")
    assert bi.tolist() == g["end_balls_idx"].tolist()
    assert_bits_equal(ot, g["end_obs_toi"], "_obstacles_toi")
    assert oo.tolist() == g["end_obs_obs"].tolist()
    if "end_toi_tri" in g:
        assert_bits_equal(np.concatenate(o.toi_table), g["end_toi_tri"], "toi_table")


def test_headline_config_prefix(oracle, golden):
    """N = 16384 dense gas (BASELINE.json config D): first 100 collisions of the live reference (recorded with
    make_golden.py --big, 500 s of add_ball + 24 s of events there)."""
    g = golden("dense_gas_16384")
    o = workloads.dense_gas(128).build_oracle(oracle, dot_fma=bool(g["dot_fma"]), flags=oracle.FAST_ROWMIN)
    t, i, j = run_logged(o, n_events=100)
    assert_bits_equal(t, g["log_t"], "event times")
    assert (i.tolist(), j.tolist()) == (g["log_i"].tolist(), g["log_j"].tolist())
    assert_bits_equal(o.state()["v"], g["end_vel"])
    bt, bi, ot, oo, _ = o.minima()
    assert_bits_equal(bt, g["end_balls_toi"])
    assert bi.tolist() == g["end_balls_idx"].tolist()


def test_wrong_dot_mode_diverges(oracle, golden):
    """the un-fused dot does not reproduce the reference (times differ in the last bits,
    and in a chaotic gas the order follows) -- the dot mode is part of the contract."""
    g = golden("dense_gas_256")
    o = workloads.dense_gas(16).build_oracle(oracle, dot_fma=not bool(g["dot_fma"]))
    t, i, j = run_logged(o, n_events=2000)
    same_order = (i.tolist(), j.tolist()) == (g["log_i"].tolist(), g["log_j"].tolist())
    same_bits = t.view(np.uint64).tolist() == g["log_t"].view(np.uint64).tolist()
    assert not (same_order and same_bits)


def test_pool_counts_and_payload(oracle, golden):
    g = golden("pool")
    o = workloads.pool().build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    o.enable_log(1000, payload=True)
    assert o.evolve(100.0) == (167, 248)
    t, i, j, p = o.log(payload=True)
    assert t[0] == 1.4236607142857143 and (i[0], j[0]) == (0, 15)
    assert_bits_equal(p, g["log_payload"], "callback payloads")


def test_galperin(oracle, golden):
    """docs/quickstart.rst:50-131 and README.md:134-135 of the reference."""
    g = golden("galperin")
    cfg = workloads.galperin()
    o = cfg.build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    assert o.next_collision[:3] == (1.8000000000000005, 0, 1)
    total, cum = 0, []
    for t in (1, 2, 3, 4, 5):
        total += sum(o.evolve(t))
        cum.append(total)
    assert cum == [0, 1, 1, 4, 314152]
    total += sum(o.evolve(16))
    assert total == 314159
    st = o.state()
    assert st["v"][0, 0] == float.fromhex("0x1.78217eaf374a7p-1") and abs(st["v"][1, 0] - 1.0) < 1e-10
    assert_bits_equal(st["pos"], g["end_pos"], "positions")
    assert_bits_equal(st["v"], g["end_vel"], "velocities")
    ke = float((cfg.mass * (st["v"] ** 2).sum(axis=1)).sum() / 2)
    assert ke == 4999999999.989935 == float(g["kinetic_energy"])
    assert o.next_ball_ball_collision == (INF, -1, 0)
    assert o.next_ball_obstacle_collision[:3] == (INF, 0, None)
    # one evolve, and the first 20000 events of the log
    o2 = cfg.build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    assert o2.evolve(16) == (157080, 157079)
    o3 = cfg.build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    t, i, j = run_logged(o3, n_events=20000)
    assert_bits_equal(t, g["log_t"], "galperin event times")
    assert i.tolist() == g["log_i"].tolist() and j.tolist() == g["log_j"].tolist()


def test_cradle_and_recompute(oracle, golden):
    g = golden("cradle")
    o = workloads.newtons_cradle(5, True).build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    assert o.next_ball_ball_collision == tuple(g["next_bb"].tolist())
    t, i, j = run_logged(o, end_time=4.0)
    assert_bits_equal(t, g["log_t"], "cradle times")
    assert (i.tolist(), j.tolist()) == (g["log_i"].tolist(), g["log_j"].tolist())
    pos = o.balls_position
    o.set_ball(2, pos=pos[2] + np.asarray([0, 1e-10]))
    o.recompute_toi(2)
    assert_bits_equal(np.concatenate(o.toi_table), g["nudge_toi_tri"], "table after nudge")
    assert o.evolve(30) == tuple(g["counts30"].tolist())
    assert_bits_equal(o.balls_position, g["end_pos"], "end positions")


def test_recompute_scenario(oracle, golden):
    """tests/test_simulation.py:550-583 of the reference, with bit-exact expectations."""
    g = golden("four_balls")
    cfg = workloads.Config("four", workloads.box_walls(-1, -1, 1, 1), [(0, 0.2), (0, 0.5), (0, -0.8), (0.7, -0.8)],
                           [(1, 2), (-1, 2), (0, -1), (0, 0)], 0.2, 1.0)
    o = cfg.build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    t, i, j = run_logged(o, end_time=10.0)
    assert_bits_equal(t, g["log_t"], "times")
    assert_bits_equal(o.balls_position, g["once_pos"], "pos at t=10")

    def check(o, tag):
        assert_bits_equal(np.concatenate(o.toi_table), g[tag + "_toi_tri"], tag + " table")
        bt, bi, ot, oo, _ = o.minima()
        assert_bits_equal(bt, g[tag + "_balls_toi"], tag)
        assert bi.tolist() == g[tag + "_balls_idx"].tolist()
        assert_bits_equal(ot, g[tag + "_obs_toi"], tag)
        assert oo.tolist() == g[tag + "_obs_obs"].tolist()
        assert_bits_equal(o.balls_position, g[tag + "_pos"], tag)

    o = cfg.build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    o.evolve(1)
    o.set_ball(0, pos=(0, 0))
    o.recompute_toi(0)
    check(o, "s1")
    o.evolve(2)
    o.set_ball(1, vel=(0, 0))
    o.set_ball(2, radius=0.1)
    o.recompute_toi([1, 2])
    check(o, "s2")
    o.evolve(3)
    o.recompute_toi()
    check(o, "s3")
    assert o.evolve(5) == tuple(g["counts5"].tolist())
    check(o, "s5")


def test_step_consistency(oracle):
    """tests/test_simulation.py:461-502: 300 partial evolves == one evolve, exactly."""
    cfg = workloads.Config("four", workloads.box_walls(-1, -1, 1, 1), [(0, 0.2), (0, 0.5), (0, -0.8), (0.7, -0.8)],
                           [(1, 2), (-1, 2), (0, -1), (0, 0)], 0.2, 1.0)
    once, step = cfg.build_oracle(oracle), cfg.build_oracle(oracle)
    once.evolve(10)
    for i in range(10):
        for j in range(1, 31):
            step.evolve(i + j / 30)
    assert_bits_equal(once.balls_position, step.balls_position)


def test_sinai(oracle, golden):
    g = golden("sinai")
    obstacles, pos, vel = workloads.sinai(64, seed=0)
    out = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos, vel, int(g["n_coll"]),
                              dot_fma=bool(g["dot_fma"]), count_hits=True)
    assert_bits_equal(out["t"], g["t"], "times")
    assert_bits_equal(out["p0"], g["p0"], "positions")
    assert_bits_equal(out["v"], g["v"], "velocities")
    assert out["hits"].tolist() == g["hits"].tolist()
    one = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos[:1], vel[:1], 10000, dot_fma=bool(g["dot_fma"]))
    assert one["t"][0] == float(g["long_t"])
    assert_bits_equal(one["p0"][0], g["long_pos"])
    assert_bits_equal(one["v"][0], g["long_vel"])
    # the single-system container gives the same answer as the ensemble loop
    ob = oracle.OracleBilliard(workloads.oracle_obstacles(obstacles), dot_fma=bool(g["dot_fma"]))
    ob.add_ball(pos[3], vel[3], 0.0, 1.0)
    assert ob.evolve(INF, max_events=2000) == (0, 2000)
    assert ob.time == g["t"][3]


def test_unit_vectors(oracle, golden):
    g = golden("unit_vectors")
    fma = bool(g["dot_fma"])
    n = len(g["toi"])
    toi = np.asarray([oracle.toi_ball_ball(g["p1"][k], g["v1"][k], g["r1"][k], g["p2"][k], g["v2"][k], g["r2"][k], fma)
                      for k in range(n)])
    assert_bits_equal(toi, g["toi"], "toi_ball_ball")
    for k in range(n):
        if g["sep"][k]:
            with pytest.raises(oracle.OracleError) as e:
                oracle.elastic_collision(g["q1"][k], g["u1"][k], g["m1"][k], g["q2"][k], g["u2"][k], g["m2"][k], fma)
            assert e.value.code == oracle.ORC_ESEPARATING
        else:
            w1, w2 = oracle.elastic_collision(g["q1"][k], g["u1"][k], g["m1"][k], g["q2"][k], g["u2"][k], g["m2"][k], fma)
            assert_bits_equal(w1, g["w1"][k])
            assert_bits_equal(w2, g["w2"][k])
    wt = np.empty(n)
    wh = np.empty(n)
    dt = np.empty(n)
    for k in range(n):
        wt[k], wh[k] = oracle.wall_detect(g["wall_start"], g["wall_normal"], g["wp"][k], g["wv"][k], g["wr"][k], fma)
        dt[k] = oracle.disk_detect(g["disk"][:2], g["disk"][2], g["wp"][k], g["wv"][k], g["wr"][k], fma)
    assert_bits_equal(wt, g["wall_t"], "wall toi")
    assert_bits_equal(wh, g["wall_headway"], "wall headway")
    assert_bits_equal(dt, g["disk_t"], "disk toi")


def test_segment_vectors(oracle, golden):
    """physics.toi_ball_point, physics.toi_and_param_ball_segment, LineSegment.detect/resolve_collision of the live
    reference on 3000 random inputs (a third aimed at the segment, a third at its end points)."""
    g = golden("segment_vectors")
    fma = bool(g["dot_fma"])
    for k in range(len(g["rad"])):
        row = g["seg"][k]
        seg = [tuple(row[0:2]), tuple(row[2:4]), tuple(row[4:6]), tuple(row[6:8])]
        pos, vel, rad = g["pos"][k], g["vel"][k], float(g["rad"][k])
        moving = bool(vel.any())
        if moving:
            assert oracle.toi_ball_point(pos, vel, rad, seg[3], dot_fma=fma) == g["point_t"][k], k
        t, u = oracle.segment_param(seg, pos, vel, rad, dot_fma=fma)
        code = 3 if u is None else (0 if isinstance(u, float) else 1 + u)
        assert (t, code) == (g["param_t"][k], int(g["param_code"][k])), k
        if code == 0:
            assert u == g["param_u"][k], k
        if moving:
            t, u = oracle.segment_detect(seg, pos, vel, rad, dot_fma=fma)
            assert t == g["detect_t"][k], k
            if np.isfinite(t):
                assert u == g["detect_u"][k], k
        for col, uu, key in ((0, 0.37, "resolve_face"), (1, 0, "resolve_start"), (2, 1, "resolve_end")):
            try:
                w = oracle.segment_resolve(seg, pos, vel, uu, dot_fma=fma)
                assert not g["resolve_sep"][k, col], k
                assert_bits_equal(w, g[key][k], f"{key}[{k}]")
            except oracle.OracleError:
                assert g["resolve_sep"][k, col], k


def test_maze_payload(oracle, golden):
    g = golden("maze_wide")
    o = workloads.maze_wide().build_oracle(oracle, dot_fma=bool(g["dot_fma"]))
    o.enable_log(4000, payload=True)
    o.evolve(float("inf"), max_events=4000)
    t, i, j, p = o.log(payload=True)
    assert_bits_equal(p, g["log_payload"], "callback payloads (position, velocities, line parameter u)")
    seg = j <= -5
    assert int(((p[seg, 6] == 0) | (p[seg, 6] == 1)).sum()) >= 30   # end-point hits are in the fixture


def test_known_answers(oracle):
    """Known answers of the reference's unit tests (tests/test_physics.py:15-70, tests/test_obstacles.py:16-77)."""
    toi = oracle.toi_ball_ball
    assert toi((0, 0), (0, 0), 1, (5, 0), (-1, 0), 1) == 3.0            # head-on
    assert toi((0, 0), (2, 1), 1, (5, 0), (1, 1), 1) == 3.0             # Galilean shift
    assert toi((0, 0), (0, 0), 1, (5, 0), (1, 0), 1) == INF             # moving away
    assert toi((0, 0), (0, 0), 1, (5, 2), (-1, 0), 1) == INF            # slide past (discriminant 0)
    assert toi((0, 0), (0, 0), 1, (1, 0), (-1, 0), 1) == INF            # overlapping
    assert toi((0, 0), (0, 0), 0, (5, 0), (-1, 0), 0) == INF            # two point particles never collide
    assert oracle.disk_detect((0, 0), 1, (-10, 0), (1, 0), 1) == 8.0
    n = workloads.wall_normal((-1, 0), (1, 0), "left")                   # ground floor, tests/test_obstacles.py:35-64
    assert tuple(n) == (0.0, 1.0)
    wd = lambda p, v, r: oracle.wall_detect((-1, 0), n, p, v, r)         # noqa: E731
    assert wd((0, 10), (0, -1), 1) == (9.0, 1.0)
    assert wd((-100, 10), (0, -1), 1) == (9.0, 1.0)
    assert wd((0, 10), (100, -1), 1) == (9.0, 1.0)
    assert wd((0, -10), (0, 1), 1)[0] == INF                             # coming from the outside
    assert wd((0, 10), (0, -1), 10) == (0.0, 1.0)                        # touching and colliding
    assert wd((0, 10), (0, 1), 10)[0] == INF
    assert wd((0, 10), (0, -1), 11)[0] == INF                            # overlap
    assert wd((0, 1), (0, -1), 0) == (1.0, 1.0) and wd((0, 0), (0, -1), 0) == (0.0, 1.0)
    nr = workloads.wall_normal((-1, 0), (1, 0), "right")                 # ceiling, :74-77
    assert oracle.wall_detect((-1, 0), nr, (0, -10), (10, 1), 1) == (9.0, 1.0)
    w1, w2 = oracle.elastic_collision((0, 0), (1, 0), 1, (2, 0), (-1, 0), 1)
    assert w1.tolist() == [-1, 0] and w2.tolist() == [1, 0]
    with pytest.raises(oracle.OracleError):
        oracle.elastic_collision((0, 0), (-1, 0), 1, (2, 0), (1, 0), 1)


def test_toi_table_contents(oracle):
    """tests/test_simulation.py:113-140,224-252 of the reference."""
    o = oracle.OracleBilliard()
    o.add_ball((0, 0), (0, 0), 1)
    assert o.next_ball_ball_collision == (INF, -1, 0)
    o.add_ball((4, 0), (-1, 0), 1)
    assert o.toi_table[1].tolist() == [2.0]
    o.add_ball((0, 4), (0, -2), 1)
    bt, bi, *_ = o.minima()
    assert bt.tolist() == [INF, 2.0, 1.0] and bi.tolist() == [-1, 0, 0]
    assert o.next_ball_ball_collision == (1.0, 0, 2)

    o = oracle.OracleBilliard()
    o.add_ball((-3, 0), (1, 0), 1, mass=0)
    o.add_ball((0, 0), (0, 0), 1, mass=42)
    o.add_ball((4, 0), (-1, 0), 1, mass=INF)
    assert [r.tolist() for r in o.toi_table] == [[], [1.0], [2.5, 2.0]]
    o.evolve(1.0)
    assert [r.tolist() for r in o.toi_table] == [[], [INF], [INF, 2.0]]
    o.evolve(2.0)
    assert o.balls_velocity[1].tolist() == [-2, 0]
    assert [r.tolist() for r in o.toi_table] == [[], [3.0], [INF, INF]]
    o.evolve(3.0)
    assert o.balls_position[0].tolist() == [-4, 0] and o.balls_velocity[0].tolist() == [-3, 0]
    o.add_ball((-8, 0), (0, 0), 1, mass=0)      # 0-mass meets 0-mass: division by zero
    with pytest.raises(oracle.OracleError) as e:
        o.evolve(10.0)
    assert e.value.code == oracle.ORC_EDIVZERO
