"""Drop-in behaviour of the Python API on the GPU: the known answers and error behaviour the reference's own test
suite pins (tests/test_simulation.py, tests/test_obstacles.py, tests/test_physics.py of the reference), re-stated
against ``billiards`` of this repo.  Everything here runs through libbilliards_b200.so."""

from math import cos, sin, sqrt

import numpy as np
import pytest
from pytest import approx

import workloads

pytestmark = pytest.mark.gpu
INF = float("inf")


@pytest.fixture(scope="module")
def bl():
    import billiards
    return billiards


def rows(table):
    return [r.tolist() for r in table]


def cradle(bl, num_balls=5, with_walls=True):
    return workloads.newtons_cradle(num_balls, with_walls).build(bl)


# ---- physics.py ---------------------------------------------------------------------------------------
def test_toi_ball_ball_known_answers(bl):
    from billiards.physics import toi_ball_ball
    assert toi_ball_ball((0, 0), (0, 0), 1, (5, 0), (-1, 0), 1) == 3
    assert toi_ball_ball((0, 42), (42, 0), 1, (5, 42), (41, 0), 1) == 3      # only relative motion matters

    def toi(p2, v2, r2, t_eps=-0.0):
        return toi_ball_ball((0, 0), (0, 0), 1, p2, v2, r2, t_eps)

    assert toi((2, 0), (1, 0), 1) == INF and toi((2, 0), (0, 1), 1) == INF   # misses
    assert toi((3, 0), (-1, 0), 1) == 1.0
    assert toi((0, 101), (0, -33), 1) == approx(3.0)
    assert toi((2, 10), (0, -1), 1) == INF                                   # slides past
    assert toi((sqrt(2), sqrt(2)), (1 - 1e-7, -1), 1) == approx(0.0, abs=1e-8)
    assert toi((1, 2), (0, -1), sqrt(2) - 1) == approx(1.0)
    assert toi((2, 0), (-1, 0), 1) == 0                                      # touching
    assert toi((1, 0), (-42, 0), 1) == INF and toi((1, 0), (-42, 0), 10) == INF   # one inside the other
    assert toi((2, 0), (-1, 0), 0) == 1 and toi((1, 0), (0, 1), 0) == INF    # point particle
    assert toi((0.5, 1), (0, -1), 0) == approx(1 - sqrt(3 / 4))
    diag = (sqrt(2), sqrt(2))
    assert toi(diag, (-1, 0), r2=1 + 1e-5) == INF
    assert toi(diag, (-1, 0), r2=1 + 1e-5, t_eps=-1e-4) == approx(0.0, abs=2e-5)
    x, y = 2 * cos(1 / 4), 2 * sin(1 / 4)
    assert toi((x, y), (-1, 0), 1) == INF                                    # rounding hides the contact ...
    assert toi((x, y), (-1, 0), 1, t_eps=-1e-10) == approx(0.0)              # ... t_eps recovers it


def test_elastic_collision_known_answers(bl):
    from billiards.physics import elastic_collision

    def ec(v1, v2, mass2=1):
        a, b = elastic_collision((0, 0), v1, 1, (2, 0), v2, mass2)
        return tuple(a), tuple(b)

    assert ec((0, 0), (-1, 0)) == ((-1, 0), (0, 0))
    assert ec((1, 0), (-1, 0)) == ((-1, 0), (1, 0))
    assert ec((1, 0), (0, 0)) == ((0, 0), (1, 0))
    assert ec((0, 0), (-1, 1)) == ((-1, 0), (0, 1))
    assert ec((0, 0), (-42, 1 / 42)) == ((-42, 0), (0, 1 / 42))
    assert ec((-1, 0), (-20, 0), mass2=0) == ((-1, 0), (18, 0))
    assert ec((0, 0), (0, 1)) == ((0, 0), (0, 1))
    assert ec((0, 0), (5e-16, 1)) == ((5e-16, 0), (0, 1))                    # below the 1e-15 threshold
    with pytest.raises(ValueError):
        ec((0, 0), (6e-16, 1))
    with pytest.raises(FloatingPointError):
        elastic_collision((0, 0), (1, 0), 0, (2, 0), (0, 0), 0)


# ---- obstacles.py -----------------------------------------------------------------------------------------
def test_disk(bl):
    d = bl.Disk((0, 0), 1)
    assert tuple(d.center) == (0, 0) and d.radius == 1
    assert d.detect_collision((-10, 0), (1, 0), 1) == (8.0, ())
    assert tuple(d.resolve_collision((-2, 0), (1, 0), 1)) == (-1, 0)


def test_infinite_wall(bl):
    with pytest.raises(ValueError):
        bl.InfiniteWall((-1, 0), (-1, 0))
    with pytest.raises(ValueError):
        bl.InfiniteWall((-1, 0), (1, 0), "Left")
    w = bl.InfiniteWall((-1, 0), (1, 0))
    assert tuple(w.start_point) == (-1.0, 0.0) and tuple(w.end_point) == (1.0, 0.0) and tuple(w._normal) == (0.0, 1.0)
    for p, v in (((0, 10), (0, -1)), ((-100, 10), (0, -1)), ((0, 10), (100, -1))):
        assert w.detect_collision(p, v, 1) == (9, (1.0,))
    assert w.detect_collision((0, -10), (0, 1), 1)[0] == INF
    assert w.detect_collision((0, 10), (0, -1), 10) == (0, (1.0,))
    assert w.detect_collision((0, 10), (0, 1), 10)[0] == INF
    assert w.detect_collision((0, 10), (0, -1), 11)[0] == INF
    assert w.detect_collision((0, 1), (0, -1), 0) == (1, (1.0,))
    assert w.detect_collision((0, 0), (0, -1), 0) == (0, (1.0,))
    assert w.detect_collision((0, 0), (0, 1), 0)[0] == INF
    assert tuple(w.resolve_collision((0, 10), (0, -1), 1, 1.0)) == (0, 1)
    assert tuple(w.resolve_collision((0, 10), (10, -1), 1, 1.0)) == (10, 1)
    with pytest.raises(AssertionError):
        w.resolve_collision((0, -10), (10, 1), 1, -1.0)
    c = bl.InfiniteWall((-1, 0), (1, 0), exterior="right")
    assert c.detect_collision((0, -10), (10, 1), 1) == (9, (1.0,))
    assert tuple(c.resolve_collision((0, -10), (10, 1), 1, 1.0)) == (10, -1)


def test_line_segment_known_answers(bl):
    """tests/test_obstacles.py:80-147 and tests/test_physics.py:73-108 of the reference."""
    from math import cos, pi, sin, sqrt
    from billiards.physics import toi_and_param_ball_segment, toi_ball_point
    LineSegment = bl.LineSegment
    with pytest.raises(ValueError):
        LineSegment((-1, 0), (-1, 0))
    line = LineSegment((-1, 0), (1, 0))
    np.testing.assert_allclose(line._covector, (1 / 2, 0))
    np.testing.assert_allclose(line._normal, (0, 1))
    for a in [0, 1 / 6, 1 / 4, -1 / 2, pi / 2 - 1e-6]:      # left end point
        pos = np.asarray([-cos(a) - 1, sin(a)])
        vel = np.asarray([cos(a), -sin(a)])
        assert line.detect_collision(pos, vel, 1 / 2) == (approx(0.5), (0,)), a
        np.testing.assert_allclose(line.resolve_collision(pos + 0.5 * vel, vel, 1 / 2, 0), (-vel[0], -vel[1]), atol=1e-14)
    pos, vel = np.asarray([-sqrt(1 / 2) - 2, sqrt(1 / 2) - 1]), np.asarray([1, 1])
    assert line.detect_collision(pos, vel, 1 + 1e-14) == (approx(1.0), (0,))
    pos, vel = np.asarray([-2, 1 - 1e-14]), np.asarray([1, 0])
    assert line.detect_collision(pos, vel, 1) == (approx(1.0), (0,))
    np.testing.assert_allclose(line.resolve_collision(pos + 1.0 * vel, vel, 1, 0), vel, atol=1e-14)
    for a in [pi / 2, pi / 2 + 1e-10, pi / 2 + 1e-6, 5 / 6 * pi - 1e-15]:   # along the line
        pos = np.asarray([-cos(a) - 1, sin(a)])
        vel = np.asarray([cos(a), -sin(a)])
        t, (u,) = line.detect_collision(pos, vel, 1 / 2)
        assert t == approx((pos[1] - 1 / 2) / (-vel[1])), a
        assert u is not None and u > 0, a
        np.testing.assert_allclose(line.resolve_collision(pos + t * vel, vel, 1 / 2, u), (vel[0], -vel[1]), atol=1e-14)
    for a in [0, 1 / 6, 1 / 4, -1 / 2, pi / 2 - 1e-6]:      # right end point
        pos = np.asarray([-cos(a + pi) + 1, sin(a + pi)])
        vel = np.asarray([cos(a + pi), -sin(a + pi)])
        assert line.detect_collision(pos, vel, 1 / 2) == (approx(0.5), (1,)), a
        np.testing.assert_allclose(line.resolve_collision(pos + 0.5 * vel, vel, 1 / 2, 1), (-vel[0], -vel[1]), atol=1e-14)
    pos, vel = np.asarray([2, 1 - 1e-14]), np.asarray([-1, 0])
    assert line.detect_collision(pos, vel, 1) == (approx(1.0), (1,))
    assert line.detect_collision((0, 5), (0, 1), 0.5) == (INF, (None,))
    # physics.toi_ball_point
    assert toi_ball_point((0, 0), (1, 0), 1, (5, 0)) == 4
    assert toi_ball_point((0, 42), (1, 0), 1, (5, 42)) == 4
    assert toi_ball_point((2, 0), (1, 0), 1, (0, 0), -0.0) == INF
    assert toi_ball_point((2, 0), (-1, 0), 1, (0, 0), -0.0) == 1.0
    assert toi_ball_point((1, 2), (0, -1), sqrt(2), (0, 0), -0.0) == approx(1.0)
    assert toi_ball_point((1 - 1e-12, 0), (-42, 0), 1, (0, 0), -0.0) == INF
    # physics.toi_and_param_ball_segment on the segment (0,0)-(1,0)
    seg = (np.asarray([0, 0]), np.asarray([1.0, 0.0]), np.asarray([0.0, 1.0]))
    t, u = toi_and_param_ball_segment((0.5, 2), (0, -1), 1, *seg)
    assert (t, u) == (approx(1.0), approx(0.5))
    assert toi_and_param_ball_segment((0.5, 2), (0, 1), 1, *seg) == (INF, None)
    assert toi_and_param_ball_segment((-3, 0.5), (1, 0), 1, *seg) == (INF, 0)
    assert toi_and_param_ball_segment((4, 0.5), (-1, 0), 1, *seg) == (INF, 1)


def test_unsupported_obstacles_fail_loudly(bl):
    class Custom(bl.Obstacle):
        def detect_collision(self, pos, vel, radius):
            return INF, ()

    with pytest.raises(NotImplementedError):
        bl.Billiard([Custom()])
    with pytest.raises(TypeError):
        bl.Billiard(obstacles=[42])


# ---- simulation.py ------------------------------------------------------------------------------------------
def test_time_and_empty_table(bl):
    b = bl.Billiard()
    assert b.time == 0.0 and b.count == 0
    assert b.evolve(1.0) == (0, 0) and b.time == 1.0
    assert b.evolve(42.0) == (0, 0) and b.time == 42.0
    assert b.next_ball_ball_collision == (INF, -1, 0)


def test_index_arguments_and_free_flight(bl):
    b = bl.Billiard()
    for i in range(10):
        assert b.add_ball((i, 0), (0, i)) == i
    assert b.count == 10
    for i in range(10):
        assert tuple(b.balls_position[i]) == (i, 0) and tuple(b.balls_velocity[i]) == (0, i)
    for bad in (((0, 0), None), ((0, 0), (0, 0, 0))):
        with pytest.raises(ValueError):
            bl.Billiard().add_ball(*bad)
    with pytest.raises(TypeError):
        bl.Billiard().add_ball((0, 0), (0, 0), None)
    with pytest.raises(TypeError):
        bl.Billiard().add_ball((0, 0), (0, 0), 1, None)
    assert bl.Billiard().add_ball((0, 0), (0, 0), 0, 0) == 0
    f = bl.Billiard()
    for i in range(10):
        f.add_ball((i, 0), (0, 1))
    assert f.evolve(42.0) == (0, 0)
    for i in range(10):
        assert tuple(f.balls_position[i]) == (i, 42.0) and tuple(f.balls_velocity[i]) == (0, 1)


def test_toi_structure_and_contents(bl):
    b = bl.Billiard()
    for i in range(1, 10):
        b.add_ball((i, 0), (0, 1))
        assert b.count == i and len(b.toi_table) == i and b.toi_table[i - 1].shape == (i - 1,)
        assert b._balls_toi.shape == (i,) and len(b._balls_idx) == i
        assert b._obstacles_toi.shape == (i,) and len(b._obstacles_obs) == i
    b = bl.Billiard()
    b.add_ball((0, 0), (0, 0), 1)
    assert b._balls_toi.tolist() == [INF] and b._balls_idx == [-1]
    assert b.next_ball_ball_collision == (INF, -1, 0)
    b.add_ball((4, 0), (-1, 0), 1)
    assert b.toi_table[1].tolist() == [2.0]
    assert b._balls_toi.tolist() == [INF, 2.0] and b._balls_idx == [-1, 0]
    assert b.next_ball_ball_collision == (2.0, 0, 1)
    b.add_ball((0, 4), (0, -2), 1)
    assert b.toi_table[2].tolist() == [1.0, approx(2.0)]
    assert b._balls_toi.tolist() == [INF, 2.0, 1.0] and b._balls_idx == [-1, 0, 0]
    assert b.next_ball_ball_collision == (1.0, 0, 2)
    assert b._detect_ball_collision(0, 1) == 2.0 and b._detect_ball_collision(0, 2) == 1.0
    assert b._detect_ball_collision(1, 2) == approx(2.0)


def test_simple_collision_with_callbacks(bl):
    b = bl.Billiard()
    b.add_ball((2, 0), (4, 0), radius=1)
    b.evolve(10.0)
    assert tuple(b.balls_position[0]) == (42.0, 0.0) and tuple(b.balls_velocity[0]) == (4.0, 0.0)
    b.add_ball((50, 18), (0, -9), radius=1, mass=2)       # added mid-simulation: its row is solved at t = 10
    toi = approx(11.79693)
    assert b.next_ball_ball_collision == (toi, 0, 1)
    seen = []

    def record(t, p, u, v, j):
        seen.append((t, p.tolist(), u.tolist(), v.tolist(), j))

    assert b.evolve(14.0, ball_callbacks={0: record, 1: record}) == (1, 0)
    assert b.time == 14 and len(seen) == 2
    assert seen[0] == (toi, [approx(49.18772), 0], [4, 0], [approx(-4 / 3), -12], 1)
    assert seen[1] == (toi, [50, approx(1.827623)], [0, -9], [approx(8 / 3), -3], 0)
    assert tuple(b.balls_position[0]) == (approx(46.2503), approx(-26.43683))
    assert tuple(b.balls_position[1]) == (approx(55.8748), approx(-4.78158))
    assert tuple(b.balls_velocity[0]) == (approx(-4 / 3), -12) and tuple(b.balls_velocity[1]) == (approx(8 / 3), -3)


def test_newton_cradle_touching_balls(bl):
    b = bl.Billiard()
    b.add_ball((-3, 0), (1, 0), 1)
    b.add_ball((0, 0), (0, 0), 1)
    b.add_ball((5, 0), (0, 0), 1)
    b.add_ball((3, 0), (0, 0), 1)      # touches the third ball
    assert b.next_ball_ball_collision == (1.0, 0, 1)
    assert b.evolve(1.0) == (1, 0)
    assert tuple(b.balls_position[0]) == (-2, 0) and tuple(b.balls_velocity[0]) == (0, 0)
    assert tuple(b.balls_position[1]) == (0, 0) and tuple(b.balls_velocity[1]) == (1, 0)
    assert b.next_ball_ball_collision == (2.0, 1, 3)
    assert b.evolve(11.0) == (2, 0)
    assert tuple(b.balls_position[1]) == (1, 0) and tuple(b.balls_velocity[1]) == (0, 0)
    assert tuple(b.balls_position[2]) == (5 + (11 - 2) * 1, 0) and tuple(b.balls_velocity[2]) == (1, 0)
    assert tuple(b.balls_position[3]) == (3, 0) and tuple(b.balls_velocity[3]) == (0, 0)
    assert rows(b.toi_table) == [[], [INF], [INF, INF], [INF, INF, INF]]
    assert b._balls_toi.tolist() == [INF] * 4 and b._balls_idx == [-1, 0, 0, 0]
    assert b.next_ball_ball_collision == (INF, -1, 0)


def test_masses_zero_finite_infinite(bl):
    b = bl.Billiard()
    b.add_ball((-3, 0), (1, 0), 1, mass=0)
    b.add_ball((0, 0), (0, 0), 1, mass=42)
    b.add_ball((4, 0), (-1, 0), 1, mass=INF)
    assert rows(b.toi_table) == [[], [1.0], [2.5, 2.0]]
    b.evolve(1.0)
    assert tuple(b.balls_velocity[0]) == (-1, 0) and tuple(b.balls_velocity[1]) == (0, 0)
    assert rows(b.toi_table) == [[], [INF], [INF, 2.0]]
    b.evolve(2.0)
    assert tuple(b.balls_velocity[1]) == (-2, 0) and tuple(b.balls_position[2]) == (2, 0)
    assert rows(b.toi_table) == [[], [3.0], [INF, INF]]
    b.evolve(3.0)
    assert tuple(b.balls_position[0]) == (-4, 0) and tuple(b.balls_velocity[0]) == (-3, 0)
    assert rows(b.toi_table) == [[], [INF], [INF, INF]]


def test_exceptional_balls(bl):
    b = bl.Billiard()
    b.add_ball((-3, 0), (1, 0), 0, mass=0)
    b.add_ball((-2, 0), (0, 0), 0, mass=42)
    assert rows(b.toi_table) == [[], [INF]]                 # two point particles never meet
    b.add_ball((0, 0), (0, 0), 1, mass=INF)
    b.add_ball((100, 0), (-20, 0), 1, mass=INF)
    assert b.toi_table[2].tolist() == [2.0, INF]
    assert b.toi_table[3].tolist() == [approx(102 / 21), approx(101 / 20), approx(98 / 20)]
    b.evolve(2.0)
    assert tuple(b.balls_velocity[0]) == (-1, 0) and tuple(b.balls_velocity[2]) == (0, 0)
    assert b.next_ball_ball_collision == (approx(98 / 20), 2, 3)
    b.evolve(5.0)                                           # inf-mass meets inf-mass: the first one wins
    assert tuple(b.balls_velocity[2]) == (0, 0) and tuple(b.balls_velocity[3]) == (20, 0)
    assert b.toi_table[3].tolist() == [INF, INF, INF]
    b.add_ball((-6, 0), (0, 0), 1, mass=0)
    assert b.next_ball_ball_collision == (approx(6.0), 0, 4)
    with pytest.raises(FloatingPointError):                 # two massless balls
        b.evolve(7.0)


def test_obstacle_cache_and_bounce(bl):
    disk = bl.Disk((0, 0), radius=1)
    b = bl.Billiard(obstacles=[disk])
    assert b.obstacles == [disk]
    b.add_ball((-10, 0), (1, 0), radius=1)
    assert b._obstacles_toi.tolist() == [8.0] and b._obstacles_obs == [(disk, ())]
    assert b.next_ball_obstacle_collision == (8.0, 0, (disk, ()))
    seen = []
    b.bounce_ball_obstacle(ball_callbacks={0: lambda t, p, u, v, o: seen.append((t, p.tolist(), u.tolist(), v.tolist(), o))})
    assert seen == [(8.0, [-2.0, 0.0], [1.0, 0.0], [-1.0, 0.0], disk)]
    assert b._obstacles_toi.tolist() == [INF] and b._obstacles_obs == [None]
    assert b.next_ball_obstacle_collision == (INF, 0, None)
    assert tuple(b.balls_velocity[0]) == (-1.0, 0.0)


def test_cradle_with_walls_and_callback_order(bl):
    b = cradle(bl, 2)
    left, right = b.obstacles
    assert b.next_ball_ball_collision == (3.0, 0, 1)
    assert b._obstacles_toi.tolist() == [9.0, INF] and b._obstacles_obs == [(right, (1.0,)), None]
    assert b.next_ball_obstacle_collision == (9.0, 0, (right, (1.0,)))
    assert b.next_collision == b.next_ball_ball_collision
    assert b.evolve(3.0) == (1, 0)
    assert b.next_ball_ball_collision == (INF, -1, 0)
    assert b._obstacles_toi.tolist() == [INF, 7.0] and b._obstacles_obs == [None, (right, (1.0,))]
    assert b.next_collision == b.next_ball_obstacle_collision == (7.0, 1, (right, (1.0,)))
    assert b.evolve(7.0) == (0, 1)
    assert b.next_ball_ball_collision == (11.0, 0, 1)
    assert b.next_ball_obstacle_collision == (16.0, 1, (left, (1.0,)))
    assert b.evolve(14.0) == (1, 1)

    b = cradle(bl, 2)
    left, right = b.obstacles
    for bad in (lambda x: x, [lambda x: x]):
        with pytest.raises(TypeError):
            b.evolve(15.0, ball_callbacks=bad)
        with pytest.raises(TypeError):
            b.evolve(15.0, obstacle_callbacks=bad)
    with pytest.raises(ValueError):
        b.evolve(15.0, ball_callbacks={"1": lambda x: x})
    with pytest.raises(ValueError):
        b.evolve(15.0, obstacle_callbacks={42: lambda x: x})
    times, hits, obs_hits = [], [], []
    ball_cb = {i: (lambda i: lambda t, p, u, v, other: hits.append((t, i, other)))(i) for i in range(b.count)}
    obs_cb = {o: (lambda o: lambda t, p, u, v, args, i: obs_hits.append((t, o, i)))(o) for o in b.obstacles}
    assert b.evolve(15.0, time_callback=times.append, ball_callbacks=ball_cb, obstacle_callbacks=obs_cb) == (2, 2)
    assert times == [3.0, 7.0, 11.0, 14.0]
    assert hits == [(3.0, 0, 1), (3.0, 1, 0), (7.0, 1, right), (11.0, 0, 1), (11.0, 1, 0), (14.0, 0, left)]
    assert obs_hits == [(7.0, right, 1), (14.0, left, 0)]
    assert b.time == 15.0


def test_bounce_ball_ball_single_steps(bl):
    b = cradle(bl, 3, with_walls=False)
    t1 = b.next_ball_ball_collision
    b.bounce_ball_ball()
    assert b.time == t1[0]
    assert b.next_ball_ball_collision[1:] == (1, 2)
    b.bounce_ball_ball()
    assert b.next_ball_ball_collision == (INF, -1, 0)
    with pytest.raises(RuntimeError):
        b.bounce_ball_ball()


def test_sinai_example_recompute_all(bl, oracle):
    """examples/sinai_billiard.py:33-41: many point particles in ONE Billiard, slowed down in place, recompute_toi()."""
    from billiards import _lib
    obstacles, pos, vel = workloads.sinai(50, seed=9)
    b = bl.Billiard(workloads.make_obstacles(bl, obstacles))
    for p, v in zip(pos, vel):
        b.add_ball(p, v, radius=0)
    b.balls_velocity /= 5
    b.recompute_toi()
    counts = b.evolve(10.0)
    assert counts[0] == 0 and counts[1] > 50
    ref = oracle.OracleBilliard(workloads.oracle_obstacles(obstacles), dot_fma=_lib.probe_dot_fma())
    ref.add_balls(pos, vel / 5, 0.0, 1.0)
    assert ref.evolve(10.0) == counts
    assert np.array_equal(b.balls_position.view(np.uint64), ref.balls_position.view(np.uint64))


def test_mass_edit_is_seen_without_recompute(bl):
    """balls_mass is read at resolve time by the reference (simulation.py:572-576)."""
    b = bl.Billiard()
    b.add_ball((0, 0), (1, 0), 1, mass=1)
    b.add_ball((5, 0), (0, 0), 1, mass=1)
    assert b.next_ball_ball_collision[0] == 3.0
    b.balls_mass[1] = INF
    b.evolve(4.0)
    assert tuple(b.balls_velocity[0]) == (-1, 0) and tuple(b.balls_velocity[1]) == (0, 0)


def test_evolve_event_limit_stops_an_accumulation_of_collisions(bl):
    """A zero-mass ball between two approaching heavy balls collides infinitely often in finite time: evolve(end_time)
    never returns in the reference (nor would the kernel).  With Billiard.evolve_event_limit the call raises after that
    many collisions and leaves a resumable state; BilliardEnsemble.evolve(max_collisions=) bounds each system."""
    bld = bl.Billiard()
    bld.add_ball((0.0, 0.0), (1.0, 0.0), radius=1.0, mass=10.0)
    bld.add_ball((5.0, 0.0), (0.0, 0.0), radius=0.1, mass=0.0)
    bld.add_ball((10.0, 0.0), (-1.0, 0.0), radius=1.0, mass=10.0)
    bld.evolve_event_limit = 3000
    with pytest.raises(RuntimeError):
        bld.evolve(100.0)
    assert bld.time < 100.0 and np.isfinite(bld.balls_position).all()
    t = bld.time
    with pytest.raises(RuntimeError):
        bld.evolve(100.0)            # resumable: the next 3000 collisions
    assert bld.time >= t
    pos = np.repeat(np.asarray([[(0.0, 0.0), (5.0, 0.0), (10.0, 0.0)]]), 4, axis=0)
    vel = np.repeat(np.asarray([[(1.0, 0.0), (0.0, 0.0), (-1.0, 0.0)]]), 4, axis=0)
    ens = bl.BilliardEnsemble([], pos, vel, radius=[1.0, 0.1, 1.0], mass=[10.0, 0.0, 10.0])
    bb, bo = ens.evolve(100.0, max_collisions=2500)
    assert (bb == 2500).all() and (bo == 0).all() and (ens.time < 100.0).all()


# ---- the reference's example scripts, headless ------------------------------------------------------------
EXAMPLE_COUNTS = {
    # collisions the live reference counts for the same set-up (tests/golden: collapse 2403 events to t = 15, Galperin's
    # 314159); the others are pinned by the frames-equal-one-call property below
    "collapse": 2403, "pi_with_pool": 314159,
}


@pytest.mark.parametrize("name", ["collapse", "ideal_gas", "maze", "newtons_cradle", "pi_with_pool", "pool_first_shot",
                                  "sinai_billiard"])
def test_reference_examples_run_headless(bl, name):
    """examples/*.py of the reference through examples/headless.py: advanced frame by frame the way visualize.animate
    does (evolve(t + dt), read positions and velocities) and in one evolve(end_time) call -- same counts, same bits."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "headless.py")
    spec = importlib.util.spec_from_file_location("examples_headless", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    frames = mod.run(name)
    once = mod.run(name, 0.0)
    assert frames["frames"] > 100 and once["frames"] == 1
    for key in ("balls", "end_time", "ball_ball", "ball_obstacle", "mean_speed", "kinetic_energy", "position_checksum"):
        assert frames[key] == once[key], key
    if name in EXAMPLE_COUNTS:
        assert frames["ball_ball"] + frames["ball_obstacle"] == EXAMPLE_COUNTS[name]
    if name == "pool_first_shot":
        assert frames["ball_ball"] + frames["ball_obstacle"] > 100


def test_evolve_snapshot_entry_point_pinned_and_pageable_host_memory(bl):
    """bl_evolve_snapshot (include/billiards_b200.h): evolve + host mirrors in one wait.  Page-locked host memory is
    written by the snapshot kernel directly, pageable memory through a staged copy -- same bytes, equal to what
    bl_evolve followed by bl_snapshot gives."""
    import ctypes
    import torch
    from billiards import _lib
    cfg = workloads.hetero_gas()
    out = []
    for kind in ("pinned", "pageable", "separate"):
        b = cfg.build(bl)
        b._flush()
        h, n = b._h(), b._n
        s = b._system()
        res = _lib.BlResult()
        d_snap = torch.empty(8 * n, dtype=torch.float64, device=h.torch_device)
        if kind == "pinned":
            host = torch.empty(8 * n, dtype=torch.float64, pin_memory=True).numpy()
        else:
            host = np.empty(8 * n, dtype=np.float64)
        host[:] = -1.0
        hp = ctypes.c_void_p(host.ctypes.data)
        if kind == "separate":
            h.check(h.lib.bl_evolve(h.h, ctypes.byref(s), 0.0, 2.5, -1, _lib.BL_ANY, None, ctypes.byref(res), h.stream()))
            h.check(h.lib.bl_snapshot(h.h, ctypes.byref(s), res.time, _lib.ptr(d_snap), hp, h.stream()))
        else:
            h.check(h.lib.bl_evolve_snapshot(h.h, ctypes.byref(s), 0.0, 2.5, -1, _lib.BL_ANY, None, ctypes.byref(res),
                                             _lib.ptr(d_snap), hp, h.stream()))
        assert res.time == 2.5 and res.n_ball_ball + res.n_ball_obstacle > 50
        assert np.array_equal(d_snap.cpu().numpy().view(np.uint64), host.view(np.uint64))
        out.append((host.copy(), res.n_ball_ball, res.n_ball_obstacle))
    for other in out[1:]:
        assert np.array_equal(out[0][0].view(np.uint64), other[0].view(np.uint64))
        assert out[0][1:] == other[1:]


def test_two_host_threads_share_one_handle(bl):
    """One bl_handle per (device, dot mode) is shared by every Billiard of the process and ctypes drops the GIL during
    a call: the entry points serialise on the handle's mutex (csrc/bl_internal.h: bl_entry_guard), so two threads
    evolving their own billiards -- a gas on the grid-wide kernel (barrier counters, exchange buffers and result block
    live in the handle) and a pool table on the lane kernel -- get exactly what they get alone."""
    import threading

    def gas_job():
        b = workloads.dense_gas(12).build(bl)
        out = []
        for f in range(1, 41):
            out.append(b.evolve(0.05 * f))
            out.append(b.balls_position.copy())
        return out

    def pool_job():
        b = workloads.pool().build(bl)
        out = []
        for f in range(1, 201):
            out.append(b.evolve(0.5 * f))
            out.append(b.balls_velocity.copy())
        return out

    alone = [gas_job(), pool_job()]
    got = [None, None]
    errors = []

    def worker(slot, job):
        try:
            got[slot] = job()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(0, gas_job)), threading.Thread(target=worker, args=(1, pool_job))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    for a, g in zip(alone, got):
        assert len(a) == len(g)
        for x, y in zip(a, g):
            if isinstance(x, tuple):
                assert x == y
            else:
                assert np.array_equal(x.view(np.uint64), y.view(np.uint64))
