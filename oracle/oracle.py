"""ctypes front end of the CPU oracle (``oracle/billiards_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package under
``python-billiards_b200/`` never imports this module.

The class mirrors the part of ``billiards.simulation.Billiard`` that lies on the hot path
(reference ``src/billiards/simulation.py:59-636``); obstacles are passed as plain numbers
(wall = start point + unit normal as computed by the reference constructor,
``obstacles.py:98-110``; disk = centre + radius).
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAST_ROWMIN = 1

ORC_ESEPARATING = -3
ORC_EDIVZERO = -4


class _Obstacle(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("a", ctypes.c_double), ("b", ctypes.c_double),
                ("c", ctypes.c_double), ("d", ctypes.c_double), ("e", ctypes.c_double), ("f", ctypes.c_double),
                ("g", ctypes.c_double), ("h", ctypes.c_double)]


def build(force=False):
    """Compile liboracle.so next to this file (gcc, see Makefile)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "billiards_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        vp = ctypes.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [ctypes.c_int, ctypes.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_add_wall.argtypes = [vp] + [ctypes.c_double] * 4
        L.orc_add_disk.argtypes = [vp] + [ctypes.c_double] * 3
        L.orc_add_segment.argtypes = [vp, dp]
        L.orc_toi_ball_point.restype = ctypes.c_double
        L.orc_toi_ball_point.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, dp, ctypes.c_double]
        L.orc_segment_param.restype = ctypes.c_double
        L.orc_segment_param.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp,
                                        ctypes.POINTER(ctypes.c_int)]
        L.orc_segment_detect.restype = ctypes.c_double
        L.orc_segment_detect.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, dp]
        L.orc_segment_resolve.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, dp]
        L.orc_add_ball.argtypes = [vp] + [ctypes.c_double] * 6
        L.orc_enable_log.argtypes = [vp, ctypes.c_longlong, ctypes.c_int]
        L.orc_bounce_ball_ball.argtypes = [vp]
        L.orc_bounce_ball_obstacle.argtypes = [vp]
        L.orc_evolve.argtypes = [vp, ctypes.c_double, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong)]
        L.orc_set_ball.argtypes = [vp, ctypes.c_int, dp, dp, dp, dp]
        L.orc_recompute_toi.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.orc_set_time.argtypes = [vp, ctypes.c_double]
        L.orc_count.argtypes = [vp]
        L.orc_time.argtypes = [vp]
        L.orc_time.restype = ctypes.c_double
        L.orc_solves.argtypes = [vp]
        L.orc_solves.restype = ctypes.c_longlong
        L.orc_log_len.argtypes = [vp]
        L.orc_log_len.restype = ctypes.c_longlong
        L.orc_get_log.argtypes = [vp, dp, ip, ip, dp]
        L.orc_clear_log.argtypes = [vp]
        L.orc_get_state.argtypes = [vp, dp, dp, dp, dp, dp, dp]
        L.orc_get_table.argtypes = [vp, dp]
        L.orc_get_minima.argtypes = [vp, dp, ip, dp, ctypes.POINTER(ctypes.c_int), dp]
        L.orc_get_next.argtypes = [vp, dp, ip, dp]
        L.orc_toi_ball_ball.restype = ctypes.c_double
        L.orc_toi_ball_ball.argtypes = [ctypes.c_int, dp]
        L.orc_elastic_collision.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, dp, dp, ctypes.c_double, dp, dp]
        L.orc_wall_detect.restype = ctypes.c_double
        L.orc_wall_detect.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, dp]
        L.orc_disk_detect.restype = ctypes.c_double
        L.orc_disk_detect.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double]
        L.orc_ensemble_run.argtypes = [ctypes.c_int, ctypes.POINTER(_Obstacle), ctypes.c_int, ctypes.c_longlong,
                                       ctypes.c_double, dp, dp, ctypes.c_longlong,
                                       ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int32)]
        L.orc_ensemble_run2.argtypes = L.orc_ensemble_run.argtypes + [ctypes.POINTER(ctypes.c_int32)]
        _LIB = L
    return _LIB


def probe_dot_fma():
    """Which formula does THIS host's numpy use for a 2-vector dot?  True: fma(a1, b1, fl(a0*b0)) (OpenBLAS on FMA
    hosts), False: fl(fl(a0*b0) + fl(a1*b1)).  The oracle restates whichever the reference
    would compute on this machine; bench.py's CPU legs call this instead of importing the product package."""
    L = lib()
    L.orc_fma.restype = ctypes.c_double
    L.orc_fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(20240229)
    a, b = rng.standard_normal((2048, 2)), rng.standard_normal((2048, 2))
    fused = plain = True
    for x, y in zip(a, b):
        d = float(x.dot(y))
        x0, x1, y0, y1 = float(x[0]), float(x[1]), float(y[0]), float(y[1])
        fused = fused and d == L.orc_fma(x1, y1, x0 * y0)
        plain = plain and d == x0 * y0 + x1 * y1
    return fused if fused != plain else True


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _f64(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if n is not None and a.shape != (n,):
        raise ValueError(f"expected {n} numbers, got shape {a.shape}")
    return a


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle status {code}")
        self.code = code


class OracleBilliard:
    """Hot-path mirror of ``billiards.Billiard`` running the C restatement.

    obstacles: list of ``("wall", (sx, sy), (nx, ny))`` / ``("disk", (cx, cy), R)`` in list order.
    """

    def __init__(self, obstacles=(), dot_fma=True, flags=0):
        self._L = lib()
        self._h = self._L.orc_create(int(bool(dot_fma)), int(flags))
        self.obstacles = list(obstacles)
        for ob in self.obstacles:
            if ob[0] == "wall":
                self._L.orc_add_wall(self._h, float(ob[1][0]), float(ob[1][1]), float(ob[2][0]), float(ob[2][1]))
            elif ob[0] == "disk":
                self._L.orc_add_disk(self._h, float(ob[1][0]), float(ob[1][1]), float(ob[2]))
            elif ob[0] == "segment":  # ("segment", start, covector, normal, end), all pairs of floats
                sg = _f64([x for pair in ob[1:5] for x in pair], 8)
                self._L.orc_add_segment(self._h, _dp(sg))
            else:
                raise ValueError(ob)

    def __del__(self):
        try:
            self._L.orc_destroy(self._h)
        except Exception:
            pass

    # --- construction -------------------------------------------------------------
    def add_ball(self, pos, vel=(0.0, 0.0), radius=0.0, mass=1.0):
        return self._L.orc_add_ball(self._h, float(pos[0]), float(pos[1]), float(vel[0]), float(vel[1]),
                                    float(radius), float(mass))

    def add_balls(self, pos, vel, radius, mass):
        pos = np.asarray(pos, dtype=np.float64)
        vel = np.asarray(vel, dtype=np.float64)
        n = pos.shape[0]
        radius = np.broadcast_to(np.asarray(radius, dtype=np.float64), (n,))
        mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (n,))
        for k in range(n):
            self._L.orc_add_ball(self._h, pos[k, 0], pos[k, 1], vel[k, 0], vel[k, 1], radius[k], mass[k])

    def enable_log(self, cap, payload=False):
        self._L.orc_enable_log(self._h, int(cap), int(bool(payload)))

    # --- stepping -----------------------------------------------------------------
    def evolve(self, end_time, max_events=-1):
        c = (ctypes.c_longlong * 2)()
        st = self._L.orc_evolve(self._h, float(end_time), int(max_events), c)
        if st:
            raise OracleError(st)
        return int(c[0]), int(c[1])

    def bounce_ball_ball(self):
        st = self._L.orc_bounce_ball_ball(self._h)
        if st:
            raise OracleError(st)

    def bounce_ball_obstacle(self):
        st = self._L.orc_bounce_ball_obstacle(self._h)
        if st:
            raise OracleError(st)

    def set_ball(self, idx, pos=None, vel=None, radius=None, mass=None):
        def opt(x, n):
            return None if x is None else _dp(_f64(x, n))
        keep = [None if x is None else _f64(x, n) for x, n in ((pos, 2), (vel, 2))]
        r = None if radius is None else ctypes.c_double(float(radius))
        m = None if mass is None else ctypes.c_double(float(mass))
        st = self._L.orc_set_ball(self._h, int(idx),
                                  None if keep[0] is None else _dp(keep[0]),
                                  None if keep[1] is None else _dp(keep[1]),
                                  None if r is None else ctypes.byref(r),
                                  None if m is None else ctypes.byref(m))
        if st:
            raise OracleError(st)

    def recompute_toi(self, indices=None):
        if indices is None:
            indices = range(self.count)
        elif isinstance(indices, int):
            indices = [indices]
        idx = np.ascontiguousarray(np.asarray(list(indices), dtype=np.int32))
        st = self._L.orc_recompute_toi(self._h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), idx.size)
        if st:
            raise OracleError(st)

    # --- read-out -----------------------------------------------------------------
    @property
    def count(self):
        return self._L.orc_count(self._h)

    @property
    def time(self):
        return self._L.orc_time(self._h)

    @time.setter
    def time(self, t):
        self._L.orc_set_time(self._h, float(t))

    @property
    def solves(self):
        return self._L.orc_solves(self._h)

    def state(self):
        n = self.count
        out = {k: np.empty((n, 2)) for k in ("p0", "v", "pos")}
        out.update({k: np.empty(n) for k in ("t0", "r", "m")})
        self._L.orc_get_state(self._h, _dp(out["t0"]), _dp(out["p0"]), _dp(out["v"]), _dp(out["pos"]),
                              _dp(out["r"]), _dp(out["m"]))
        return out

    @property
    def balls_position(self):
        return self.state()["pos"]

    @property
    def balls_velocity(self):
        return self.state()["v"]

    @property
    def toi_table(self):
        n = self.count
        tri = np.empty(n * (n - 1) // 2 if n else 0)
        if tri.size:
            self._L.orc_get_table(self._h, _dp(tri))
        rows, off = [], 0
        for i in range(n):
            rows.append(tri[off:off + i].copy())
            off += i
        return rows

    def minima(self):
        n = self.count
        bt, bi = np.empty(n), np.empty(n, dtype=np.int64)
        ot, oo, oa = np.empty(n), np.empty(n, dtype=np.int32), np.empty(n)
        if n:
            self._L.orc_get_minima(self._h, _dp(bt), bi.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dp(ot),
                                   oo.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(oa))
        return bt, bi, ot, oo, oa

    def _next(self):
        tt = np.empty(2)
        idx = np.empty(4, dtype=np.int64)
        arg = ctypes.c_double()
        self._L.orc_get_next(self._h, _dp(tt), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(arg))
        return tt, idx, arg.value

    @property
    def next_ball_ball_collision(self):
        tt, idx, _ = self._next()
        return (float(tt[0]), int(idx[0]), int(idx[1]))

    @property
    def next_ball_obstacle_collision(self):
        """(time, ball, obstacle index or None, headway)"""
        tt, idx, arg = self._next()
        return (float(tt[1]), int(idx[2]), None if idx[3] < 0 else int(idx[3]), arg)

    @property
    def next_collision(self):
        bb, bo = self.next_ball_ball_collision, self.next_ball_obstacle_collision
        return bb if bb[0] <= bo[0] else bo

    def log(self, payload=False):
        n = min(self._L.orc_log_len(self._h), 1 << 62)
        t = np.empty(n)
        i = np.empty(n, dtype=np.int64)
        j = np.empty(n, dtype=np.int64)
        p = np.empty((n, 12)) if payload else None
        ip = ctypes.POINTER(ctypes.c_int64)
        if n:
            self._L.orc_get_log(self._h, _dp(t), i.ctypes.data_as(ip), j.ctypes.data_as(ip),
                                None if p is None else _dp(p))
        return (t, i, j, p) if payload else (t, i, j)

    def clear_log(self):
        self._L.orc_clear_log(self._h)


# --- scalar kernels ------------------------------------------------------------------

def toi_ball_ball(p1, v1, r1, p2, v2, r2, dot_fma=True):
    a = _f64([p1[0], p1[1], v1[0], v1[1], r1, p2[0], p2[1], v2[0], v2[1], r2])
    return lib().orc_toi_ball_ball(int(bool(dot_fma)), _dp(a))


def elastic_collision(p1, v1, m1, p2, v2, m2, dot_fma=True):
    w1, w2 = np.empty(2), np.empty(2)
    a = [_f64(x, 2) for x in (p1, v1, p2, v2)]
    st = lib().orc_elastic_collision(int(bool(dot_fma)), _dp(a[0]), _dp(a[1]), float(m1), _dp(a[2]), _dp(a[3]),
                                     float(m2), _dp(w1), _dp(w2))
    if st:
        raise OracleError(st)
    return w1, w2


def wall_detect(start, normal, pos, vel, radius, dot_fma=True):
    w = _f64([start[0], start[1], normal[0], normal[1]])
    hw = ctypes.c_double()
    t = lib().orc_wall_detect(int(bool(dot_fma)), _dp(w), _dp(_f64(pos, 2)), _dp(_f64(vel, 2)), float(radius),
                              ctypes.byref(hw))
    return t, hw.value


def disk_detect(center, R, pos, vel, radius, dot_fma=True):
    d = _f64([center[0], center[1], R])
    return lib().orc_disk_detect(int(bool(dot_fma)), _dp(d), _dp(_f64(pos, 2)), _dp(_f64(vel, 2)), float(radius))


def toi_ball_point(pos, vel, radius, point, t_eps=-1e-10, dot_fma=True):
    """physics.toi_ball_point (physics.py:79-144)"""
    return lib().orc_toi_ball_point(int(bool(dot_fma)), _dp(_f64(pos, 2)), _dp(_f64(vel, 2)), float(radius),
                                    _dp(_f64(point, 2)), float(t_eps))


def _segment8(segment):
    """("segment", start, covector, normal, end) or the four pairs -> 8 doubles"""
    parts = segment[1:5] if segment[0] == "segment" else segment
    return _f64([x for pair in parts for x in pair], 8)


def segment_param(segment, pos, vel, radius, t_eps=-1e-10, dot_fma=True):
    """physics.toi_and_param_ball_segment (physics.py:147-252) -> (t, u) with u a float, 0, 1 or None"""
    u, code = ctypes.c_double(), ctypes.c_int()
    t = lib().orc_segment_param(int(bool(dot_fma)), _dp(_segment8(segment)), _dp(_f64(pos, 2)), _dp(_f64(vel, 2)),
                                float(radius), float(t_eps), ctypes.byref(u), ctypes.byref(code))
    return t, (u.value if code.value == 0 else 0 if code.value == 1 else 1 if code.value == 2 else None)


def segment_detect(segment, pos, vel, radius, dot_fma=True):
    """LineSegment.detect_collision (obstacles.py:172-183) -> (t, u)"""
    arg = ctypes.c_double()
    t = lib().orc_segment_detect(int(bool(dot_fma)), _dp(_segment8(segment)), _dp(_f64(pos, 2)), _dp(_f64(vel, 2)),
                                 float(radius), ctypes.byref(arg))
    return t, arg.value


def segment_resolve(segment, pos, vel, u, dot_fma=True):
    """LineSegment.resolve_collision (obstacles.py:185-198)"""
    w = np.empty(2)
    st = lib().orc_segment_resolve(int(bool(dot_fma)), _dp(_segment8(segment)), _dp(_f64(pos, 2)), _dp(_f64(vel, 2)),
                                   float(u), _dp(w))
    if st:
        raise OracleError(st)
    return w


def ensemble_run(obstacles, pos, vel, n_collisions, radius=0.0, dot_fma=True, count_hits=False, per_system_status=False):
    """Config S: every system = one ball + the shared obstacles; n_collisions x bounce_ball_obstacle().
    ``per_system_status``: a system whose collision cannot be resolved stops there (``out["status"]``), the others
    carry on; otherwise the first such error raises."""
    pos = np.asarray(pos, dtype=np.float64)
    vel = np.asarray(vel, dtype=np.float64)
    n_sys = pos.shape[0]
    obs = (_Obstacle * max(len(obstacles), 1))()
    for q, ob in enumerate(obstacles):
        if ob[0] == "wall":
            obs[q] = _Obstacle(0, float(ob[1][0]), float(ob[1][1]), float(ob[2][0]), float(ob[2][1]), 0.0, 0.0, 0.0, 0.0)
        elif ob[0] == "segment":
            obs[q] = _Obstacle(2, *[float(x) for pair in ob[1:5] for x in pair])
        else:
            obs[q] = _Obstacle(1, float(ob[1][0]), float(ob[1][1]), float(ob[2]), 0.0, 0.0, 0.0, 0.0, 0.0)
    state = np.ascontiguousarray(np.concatenate([pos, vel], axis=1))
    t_out = np.empty(n_sys)
    done = np.zeros(n_sys, dtype=np.int64)
    hits = np.zeros((n_sys, len(obstacles)), dtype=np.int32) if count_hits else None
    status = np.zeros(n_sys, dtype=np.int32) if per_system_status else None
    st = lib().orc_ensemble_run2(int(bool(dot_fma)), obs, len(obstacles), n_sys, float(radius), _dp(state), _dp(t_out),
                                 int(n_collisions), done.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
                                 None if hits is None else hits.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                 None if status is None else status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if st:
        raise OracleError(st)
    out = {"p0": state[:, :2].copy(), "v": state[:, 2:].copy(), "t": t_out, "done": done}
    if status is not None:
        out["status"] = status
    if hits is not None:
        out["hits"] = hits
    return out
