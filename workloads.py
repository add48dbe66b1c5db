"""Synthetic initial conditions for the BASELINE.json configs.

Shared by tests/, tests/golden/make_golden.py and bench.py so that the reference, the CPU
oracle and the CUDA path are always fed the same bits.  Obstacles are described in user terms
(``("wall", start, end, exterior)`` / ``("disk", center, radius)``), balls as arrays.

Config letters follow SURVEY.md: G Galperin, P pool first shot, I ideal gas, D dense gas,
S Sinai ensemble, H heterogeneous gas (parity stress), C Newton's cradle.
"""

from math import cos, pi, sin, sqrt

import numpy as np

INF = float("inf")


def wall_normal(start, end, exterior="left"):
    """Unit normal of an InfiniteWall exactly as the reference constructor computes it
    (/root/reference/src/billiards/obstacles.py:98-110): left normal (-dy, dx) / |.|, negated for
    exterior == "right"."""
    start = np.asarray(start)
    end = np.asarray(end)
    dx, dy = end - start
    if dx == 0.0 and dy == 0.0:
        raise ValueError("this is not a line")
    n = np.asarray([-dy, dx])
    n = n / np.linalg.norm(n)
    if exterior == "right":
        n *= -1
    elif exterior != "left":
        raise ValueError(f'exterior must be "left" or "right", not {exterior}')
    return n


def box_walls(x0, y0, x1, y1):
    """bottom, right, top, left -- the order every reference example uses
    (examples/ideal_gas.py:23-28, examples/pool_first_shot.py:18-23)."""
    return [
        ("wall", (x0, y0), (x1, y0), "left"),
        ("wall", (x1, y0), (x1, y1), "left"),
        ("wall", (x1, y1), (x0, y1), "left"),
        ("wall", (x0, y1), (x0, y0), "left"),
    ]


def oracle_obstacles(obstacles):
    """user-level obstacle list -> the numeric form oracle.OracleBilliard takes"""
    out = []
    for ob in obstacles:
        if ob[0] == "wall":
            n = wall_normal(ob[1], ob[2], ob[3])
            out.append(("wall", (float(ob[1][0]), float(ob[1][1])), (float(n[0]), float(n[1]))))
        elif ob[0] == "segment":
            out.append(segment_record(ob[1], ob[2]))
        else:
            out.append(("disk", (float(ob[1][0]), float(ob[1][1])), ob[2]))
    return out


def segment_record(start, end):
    """("segment", start, covector, normal, end) with the numbers LineSegment.__init__ derives
    (obstacles.py:159-170: direction / |direction|^2 and the left-hand unit normal), computed with numpy like there"""
    start, end = np.asarray(start), np.asarray(end)
    direction = end - start
    length_sqrd = direction.dot(direction)
    covector = direction / length_sqrd
    normal = np.array([-direction[1], direction[0]]) / sqrt(length_sqrd)
    return ("segment", (float(start[0]), float(start[1])), (float(covector[0]), float(covector[1])),
            (float(normal[0]), float(normal[1])), (float(end[0]), float(end[1])))


def make_obstacles(mod, obstacles):
    """user-level obstacle list -> obstacle objects of package ``mod`` (the reference or the drop-in)"""
    out = []
    for ob in obstacles:
        if ob[0] == "wall":
            out.append(mod.InfiniteWall(ob[1], ob[2], ob[3]))
        elif ob[0] == "segment":
            out.append(mod.LineSegment(ob[1], ob[2]))
        else:
            out.append(mod.Disk(ob[1], ob[2]))
    return out


class Config:
    def __init__(self, name, obstacles, pos, vel, radius, mass, end_time=None, n_events=None):
        self.name = name
        self.obstacles = obstacles
        self.pos = np.ascontiguousarray(pos, dtype=np.float64)
        self.vel = np.ascontiguousarray(vel, dtype=np.float64)
        n = self.pos.shape[0]
        self.radius = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (n,)))
        self.mass = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, dtype=np.float64), (n,)))
        self.end_time = end_time
        self.n_events = n_events

    @property
    def n(self):
        return self.pos.shape[0]

    def build(self, mod, **kw):
        """Instantiate with package ``mod`` through the reference's public API (add_ball one by one)."""
        bld = mod.Billiard(make_obstacles(mod, self.obstacles), **kw)
        for k in range(self.n):
            bld.add_ball(tuple(self.pos[k].tolist()), tuple(self.vel[k].tolist()), float(self.radius[k]),
                         float(self.mass[k]))
        return bld

    def build_oracle(self, oracle_mod, dot_fma=True, flags=0):
        o = oracle_mod.OracleBilliard(oracle_obstacles(self.obstacles), dot_fma=dot_fma, flags=flags)
        o.add_balls(self.pos, self.vel, self.radius, self.mass)
        return o


def galperin(digits=5):
    """Config G (README.md:58-63,114 of the reference): pi from collisions, mass ratio 100**digits."""
    obstacles = [("wall", (0, -1), (0, 1), "right")]
    pos = [(3.0, 0.0), (6.0, 0.0)]
    vel = [(0.0, 0.0), (-1.0, 0.0)]
    return Config("galperin", obstacles, pos, vel, [0.2, 1.0], [1.0, float(100 ** digits)], end_time=16.0)


def pool():
    """Config P (examples/pool_first_shot.py:17-35)."""
    width, length = 112, 224
    radius = 2.85
    pos, vel = [], []
    for i in range(5):
        for j in range(i + 1):
            x = 0.75 * length + radius * sqrt(3) * i
            y = width / 2 + radius * (2 * j - i)
            pos.append((x, y))
            vel.append((0.0, 0.0))
    pos.append((0.25 * length, width / 2))
    vel.append((length / 3, 0.0))
    return Config("pool", box_walls(0, 0, length, width), pos, vel, radius, 1.0, end_time=100.0)


def lattice_gas(side, box, radius, speed=1.0, seed=0, name="gas"):
    """side x side balls on a square lattice filling box = (x0, x1) in both directions, ball order
    ix*side+iy, directions theta = default_rng(seed).uniform(0, 2 pi, N)."""
    x0, x1 = box
    n = side * side
    k = (np.arange(side) + 0.5) / side
    xs = x0 + (x1 - x0) * k
    pos = np.stack([np.repeat(xs, side), np.tile(xs, side)], axis=1)
    theta = np.random.default_rng(seed).uniform(0.0, 2.0 * pi, n)
    vel = speed * np.stack([np.cos(theta), np.sin(theta)], axis=1)
    return Config(name, box_walls(x0, x0, x1, x1), pos, vel, radius, 1.0)


def ideal_gas(side=32, seed=0):
    """Config I: N = side**2 (1024) small disks in [-1,1]^2, r = 0.01, speed 1/5 (examples/ideal_gas.py)."""
    x0, x1 = -0.99, 0.99
    n = side * side
    k = (np.arange(side) + 0.5) / side
    xs = x0 + (x1 - x0) * k
    pos = np.stack([np.repeat(xs, side), np.tile(xs, side)], axis=1)
    theta = np.random.default_rng(seed).uniform(0.0, 2.0 * pi, n)
    vel = 0.2 * np.stack([np.cos(theta), np.sin(theta)], axis=1)
    return Config(f"ideal_gas_{n}", box_walls(-1, -1, 1, 1), pos, vel, 0.01, 1.0)


def dense_gas(side=128, seed=0, phi=0.5):
    """Config D: N = side**2 (16384) unit-lattice disks at packing fraction phi in [0, side]^2."""
    r = sqrt(phi / pi)
    cfg = lattice_gas(side, (0.0, float(side)), r, 1.0, seed, name=f"dense_gas_{side * side}")
    return cfg


def hetero_gas(side=8, seed=3):
    """Config H: random radii 0.1..0.35 and masses 0.5..5, one infinite-mass and one zero-mass ball,
    4 walls + a Disk obstacle in the middle of the box -- the parity stress case of SURVEY.md section 0."""
    rng = np.random.default_rng(seed)
    n = side * side
    box = float(side)
    k = (np.arange(side) + 0.5)
    pos = np.stack([np.repeat(k, side), np.tile(k, side)], axis=1)
    theta = rng.uniform(0.0, 2.0 * pi, n)
    vel = np.stack([np.cos(theta), np.sin(theta)], axis=1) * rng.uniform(0.5, 1.5, n)[:, None]
    # multiples of 2**-10: every (r_a + r_b)**2 is exactly representable, so the fixture does not depend on
    # libm's pow() rounding; the pow path has its own test
    radius = np.round(rng.uniform(0.1, 0.35, n) * 1024.0) / 1024.0
    mass = rng.uniform(0.5, 5.0, n)
    mass[5] = INF
    mass[11] = 0.0
    # carve out the balls overlapping the central disk
    c = box / 2
    disk_r = 0.8
    keep = np.hypot(pos[:, 0] - c, pos[:, 1] - c) > disk_r + radius + 0.05
    obstacles = box_walls(0.0, 0.0, box, box) + [("disk", (c, c), disk_r)]
    return Config(f"hetero_gas_{int(keep.sum())}", obstacles, pos[keep], vel[keep], radius[keep], mass[keep])


SINAI_OBSTACLES = box_walls(-1, -1, 1, 1) + [("disk", (0, 0), 0.5)]


def sinai(n_systems, seed=0, speed=1.0):
    """Config S: n_systems independent point particles in the Sinai billiard (examples/sinai_billiard.py:23-38).
    Positions uniform in the box, rejected while p.p <= 0.25 + 1e-6 (drawn per system first), then all
    angles; returns (obstacles, pos[n,2], vel[n,2])."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1.0, 1.0, (n_systems, 2))
    bad = (pos * pos).sum(axis=1) <= 0.25 + 1e-6
    while bad.any():
        pos[bad] = rng.uniform(-1.0, 1.0, (int(bad.sum()), 2))
        bad = (pos * pos).sum(axis=1) <= 0.25 + 1e-6
    theta = rng.uniform(0.0, 2.0 * pi, n_systems)
    vel = speed * np.stack([np.cos(theta), np.sin(theta)], axis=1)
    return SINAI_OBSTACLES, pos, vel


MAZE_OBSTACLES = box_walls(-1, -1, 1, 1) + [
    ("segment", (-0.3, -1), (-0.3, 0.5)), ("segment", (0.3, 1), (0.3, -0.5)),
    ("segment", (-0.2, -0.4), (0.2, 0.4)), ("segment", (-0.2, 0.4), (0.2, -0.4)),
]


def maze(num_balls=200, radius=0.02):
    """Config M: examples/maze.py:17-43 -- an ideal gas released in one corner of a box with four LineSegments."""
    rs = np.random.RandomState(1)  # the example calls np.random.seed(1)
    pos, vel = [], []
    for _ in range(num_balls):
        pos.append(rs.uniform((-0.98, -0.98), (-0.32, -0.32)))
        angle = rs.uniform(0, 2 * pi)
        vel.append(np.asarray([cos(angle), sin(angle)]) / 2)
    return Config("maze", MAZE_OBSTACLES, pos, vel, radius, 1.0)


def maze_wide(num_balls=150, radius=0.02, seed=7):
    """Config M2: the maze of examples/maze.py with the gas spread over the whole box, so that all four segments,
    both of their faces and their end points get hit within a few thousand collisions."""
    rng = np.random.default_rng(seed)
    segs = [(np.asarray(ob[1], dtype=np.float64), np.asarray(ob[2], dtype=np.float64)) for ob in MAZE_OBSTACLES
            if ob[0] == "segment"]
    pos = []
    while len(pos) < num_balls:
        p = rng.uniform(-0.95, 0.95, 2)
        ok = all(np.hypot(*(p - q)) > 2.5 * radius for q in pos)
        for a, b in segs:
            u = min(1.0, max(0.0, float((p - a).dot(b - a) / (b - a).dot(b - a))))
            ok = ok and np.hypot(*(p - (a + u * (b - a)))) > 2.5 * radius
        if ok:
            pos.append(p)
    theta = rng.uniform(0, 2 * pi, num_balls)
    vel = np.stack([np.cos(theta), np.sin(theta)], axis=1) / 2
    return Config("maze_wide", MAZE_OBSTACLES, pos, vel, radius, 1.0)


def collapse(num_balls=200):
    """Config X: examples/collapse.py:11-26 -- a cloud of unit balls falling towards the origin in free space (no
    obstacle at all); the example calls np.random.seed(0) and draws position then velocity per ball."""
    rs = np.random.RandomState(0)
    pos, vel = [], []
    for _ in range(num_balls):
        p = rs.normal(0, scale=1, size=2)
        v = rs.normal(0, scale=5, size=2)
        p -= v * 10
        pos.append(p)
        vel.append(v)
    return Config("collapse", [], pos, vel, 1.0, 1.0, end_time=15.0)


def newtons_cradle(num_balls=5, with_walls=True):
    """Config C: the fixture of the reference's tests (tests/conftest.py:7-28)."""
    obstacles = []
    if with_walls:
        left, right = -4, 2 * num_balls + 3
        obstacles = [("wall", (left, -2), (left, 2), "right"), ("wall", (right, -2), (right, 2), "left")]
    pos = [(-3.0, 0.0)] + [(2.0 * i, 0.0) for i in range(1, num_balls)]
    vel = [(1.0, 0.0)] + [(0.0, 0.0)] * (num_balls - 1)
    return Config("cradle", obstacles, pos, vel, 1.0, 1.0)


def event_log_hash(t, i, j):
    """sha256 prefix over (t:f64, i:i64, j:i64) records, j = -1-obstacle_index for obstacle hits
    (the hash BASELINE.md section 3 quotes)."""
    import hashlib
    rec = np.empty(len(t), dtype=[("t", "<f8"), ("i", "<i8"), ("j", "<i8")])
    rec["t"], rec["i"], rec["j"] = t, i, j
    return hashlib.sha256(rec.tobytes()).hexdigest()[:16]
