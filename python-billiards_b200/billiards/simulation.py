"""The ``Billiard`` class: a drop-in for the reference's ``billiards.simulation.Billiard``
(/root/reference/src/billiards/simulation.py:19-636) whose event loop runs on a B200.

What lives where
  * device (torch CUDA tensors, lent to ``libbilliards_b200.so`` per call): the state of record
    ``(p0, v, t0, radius, mass)`` as SoA, the symmetric all-pairs table of absolute impact times, the per-ball
    candidate minima and next-obstacle entries;
  * host: user-facing mirrors (``balls_position``, ``balls_velocity``, ... as numpy arrays; ``balls_radius``,
    ``balls_mass`` as lists), refreshed lazily after each device call and pushed back by ``recompute_toi``.

``add_ball`` only queues the ball; the queue is flushed (one table-build launch for all new rows) before anything
can observe the table or change ``time`` -- every queued row is therefore solved at its add time, as the
reference does.  There is no CPU fallback anywhere in this module.
"""

import ctypes
from collections.abc import Mapping
from math import isinf

import numpy as np

from . import _lib
from .obstacles import Obstacle

INF = float("inf")
_I64 = np.int64


class Billiard:
    """A 2-dimensional billiard table simulated event by event on the GPU.

    Same construction, methods and attributes as the reference class (simulation.py:19-57):
    ``add_ball``, ``evolve``, ``bounce_ball_ball``, ``bounce_ball_obstacle``, ``recompute_toi``;
    ``time``, ``count``, ``balls_position``, ``balls_velocity``, ``balls_radius``, ``balls_mass``,
    ``balls_initial_time``, ``balls_initial_position``, ``obstacles``, ``toi_table``, ``next_collision``,
    ``next_ball_ball_collision``, ``next_ball_obstacle_collision``.

    Extra keyword arguments: ``device`` (CUDA device index) and ``dot_fma`` (numpy-dot rounding mode to
    reproduce; default = probe the local numpy, see ``_lib.probe_dot_fma``).
    """

    def __init__(self, obstacles=None, device=None, dot_fma=None):
        if obstacles is None:
            obstacles = []
        self.obstacles = []
        for obs in obstacles:
            if not isinstance(obs, Obstacle):
                raise TypeError(f"{obs} must be an obstacle.Obstacle instance")
            self.obstacles.append(obs)
        # obstacles are described to the device once; a kind without a CUDA implementation fails here, loudly
        self._obs_records = [obs._device_record() for obs in self.obstacles]

        self._device_arg, self._dot_arg = device, dot_fma
        self._handle = None

        self._time = 0.0
        self._n = 0            # balls known to the device
        self._cap = 0          # capacity of the per-ball device buffers
        self._pitch = 0        # row pitch of the table
        self._dev = {}         # name -> torch tensor
        self._pending = []     # chunks (pos[k,2], vel[k,2], add_time) not yet on the device
        self._n_pending = 0

        # host mirrors (valid when _host_valid)
        self._pos = np.empty((0, 2), dtype=np.float64)
        self._vel = np.empty((0, 2), dtype=np.float64)
        self._t0 = np.empty((0, 2), dtype=np.float64)
        self._p0 = np.empty((0, 2), dtype=np.float64)
        self._host_valid = True
        self.balls_radius = []
        self.balls_mass = []
        self._radius_up = []   # what the device currently holds
        self._mass_up = []
        self._classes = {}     # radius -> class index
        self._class_radii = []

        self._table_dirty = False  # table rows written by the event kernel, mirrors not yet refreshed
        self._next = None      # cached BlResult with the next collisions
        self.last_result = None

    # ------------------------------------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------------------------------------
    def _h(self):
        if self._handle is None:
            self._handle = _lib.get_handle(self._device_arg, self._dot_arg)
        return self._handle

    def _alloc(self, n_new):
        """Make room for n_new balls; grows geometrically after the first allocation."""
        import torch
        h = self._h()
        if n_new <= self._cap:
            return
        cap = n_new if self._cap == 0 else max(n_new, 2 * self._cap)
        pitch = (cap + 15) // 16 * 16
        dev = h.torch_device
        new = {}
        for name in ("p0x", "p0y", "vx", "vy", "t0", "radius", "mass", "cover_t", "obs_t", "obs_arg"):
            new[name] = torch.zeros(cap, dtype=torch.float64, device=dev)
        new["rclass"] = torch.zeros(cap, dtype=torch.int32, device=dev)
        new["obs_idx"] = torch.full((cap,), -1, dtype=torch.int32, device=dev)
        new["cover_code"] = torch.zeros(cap, dtype=torch.int64, device=dev)
        new["seq"] = torch.zeros(cap, dtype=torch.int64, device=dev)
        new["seq_counter"] = self._dev.get("seq_counter", None)
        if new["seq_counter"] is None:
            new["seq_counter"] = torch.zeros(1, dtype=torch.int64, device=dev)
        new["toi"] = torch.full((cap, pitch), INF, dtype=torch.float64, device=dev)
        n = self._n
        if n:
            self._table_sync()
            for name in new:
                if name == "seq_counter":
                    continue
                if name == "toi":
                    new["toi"][:n, :n] = self._dev["toi"][:n, :n]
                else:
                    new[name][:n] = self._dev[name][:n]
        for name in ("rsum2", "rsum2_obs", "obstacles"):
            if name in self._dev:
                new[name] = self._dev[name]
        if "obstacles" not in new:
            recs = (_lib.BlObstacle * max(len(self._obs_records), 1))(*self._obs_records)
            raw = np.frombuffer(bytes(recs), dtype=np.uint8).copy()
            new["obstacles"] = torch.from_numpy(raw).to(dev)
        self._dev, self._cap, self._pitch = new, cap, pitch
        self._sys = None

    def _system(self):
        """struct bl_system for the current buffers (cached: filling it costs 19 data_ptr() calls, ~15 us, which is a
        good part of a short evolve call; dropped whenever a buffer or the ball count changes)"""
        if getattr(self, "_sys", None) is not None:
            return self._sys
        d = self._dev
        s = _lib.BlSystem()
        s.n, s.pitch, s.n_classes, s.n_obstacles = self._n, self._pitch, len(self._class_radii), len(self.obstacles)
        for name in ("p0x", "p0y", "vx", "vy", "t0", "radius", "mass", "rclass", "rsum2", "rsum2_obs", "obstacles",
                     "toi", "seq", "seq_counter", "cover_t", "cover_code", "obs_t", "obs_idx", "obs_arg"):
            setattr(s, name, d[name].data_ptr() if name in d else 0)
        self._sys = s
        return s

    def _table_sync(self):
        """The event kernel rewrites only the rows of colliding balls; make the mirrored cells current again
        before anything reads whole rows of the table (bl_symmetrize)."""
        if self._table_dirty and self._n:
            h = self._h()
            s = self._system()
            h.check(h.lib.bl_symmetrize(h.h, ctypes.byref(s), h.stream()))
        self._table_dirty = False

    def _upload_radii(self, first, last):
        """(Re)derive radius classes for balls [first, last) from balls_radius and refresh the pow tables."""
        import torch
        h = self._h()
        grew = False
        radii = self.balls_radius[first:last]
        r = np.asarray(radii, dtype=np.float64)
        inv = None
        if len(radii) > 256 and np.isfinite(r).all():
            # many balls, few radii: classify the distinct values only
            distinct, inv = np.unique(r, return_inverse=True)
            distinct = distinct.tolist()
        else:
            distinct = radii
        for x in distinct:
            if x not in self._classes:
                self._classes[x] = len(self._class_radii)
                self._class_radii.append(x)
                grew = True
        dev = h.torch_device
        if grew or "rsum2" not in self._dev:
            # (R + r)**2 for disks; r**2 = (0.0 + r)**2 for segments (toi_ball_point, physics.py:120); unused for walls
            disk_r = [float(o.radius) if rec.kind == _lib.BL_OBS_DISK else 0.0
                      for o, rec in zip(self.obstacles, self._obs_records)]
            rs, rs_obs = _lib.pow2_table(self._class_radii, disk_r)
            for q, rec in enumerate(self._obs_records):
                if rec.kind == _lib.BL_OBS_WALL:
                    rs_obs[q, :] = 0.0
            self._dev["rsum2"] = torch.from_numpy(rs).to(dev)
            self._dev["rsum2_obs"] = torch.from_numpy(np.ascontiguousarray(rs_obs).reshape(-1)).to(dev) \
                if rs_obs.size else torch.zeros(1, dtype=torch.float64, device=dev)
        if inv is not None:
            c = np.asarray([self._classes[x] for x in distinct], dtype=np.int32)[inv]
        else:
            c = np.asarray([self._classes[x] for x in radii], dtype=np.int32)
        self._sys = None
        self._dev["radius"][first:last] = torch.from_numpy(r).to(dev)
        self._dev["rclass"][first:last] = torch.from_numpy(c).to(dev)
        self._radius_up[first:last] = self.balls_radius[first:last]

    def _upload_masses(self):
        import torch
        m = np.asarray(self.balls_mass, dtype=np.float64)
        self._dev["mass"][:self._n] = torch.from_numpy(m).to(self._h().torch_device)
        self._mass_up = list(self.balls_mass)

    def _flush(self):
        """Put queued balls on the device and solve their table rows at their add time."""
        import torch
        if not self._pending:
            return
        h = self._h()
        self._table_sync()
        first, k = self._n, self._n_pending
        last = first + k
        self._alloc(last)
        dev = h.torch_device
        pos = np.concatenate([c[0] for c in self._pending], axis=0).reshape(k, 2)
        vel = np.concatenate([c[1] for c in self._pending], axis=0).reshape(k, 2)
        t_add = np.concatenate([np.full(c[0].shape[0], c[2], dtype=np.float64) for c in self._pending])
        d = self._dev
        # one host -> device copy for the five state arrays
        packed = torch.from_numpy(np.ascontiguousarray(np.stack([pos[:, 0], pos[:, 1], vel[:, 0], vel[:, 1], t_add]))).to(dev)
        for q, name in enumerate(("p0x", "p0y", "vx", "vy", "t0")):
            d[name][first:last] = packed[q]
        self._n = last
        self._sys = None
        self._snap_fresh = False
        self._radius_up.extend([None] * k)
        self._upload_radii(first, last)
        self._upload_masses()
        self._pending = []
        self._n_pending = 0
        s = self._system()
        h.check(h.lib.bl_recompute_range(h.h, ctypes.byref(s), float(self._time), first, k, h.stream()))
        h.check(h.lib.bl_rebuild_cover(h.h, ctypes.byref(s), h.stream()))
        self._next = None

    def _sync_masses(self):
        """balls_mass is read at resolve time by the reference (simulation.py:572-576): pick up list edits."""
        if self.balls_mass != self._mass_up:
            self._upload_masses()

    def _sync_host(self):
        """Refresh the numpy mirrors from the device (positions at the current time)."""
        import torch
        if self._host_valid:
            return
        self._flush()
        h, n = self._h(), self._n
        if n:
            # one launch + one device -> pinned-host copy (bl_snapshot); the mirrors are fresh arrays, like the
            # reference's, so that references a caller kept to earlier arrays stay what they were
            snap = self._snap_buffers(n)
            if not self._snap_fresh:  # (an evolve call leaves the snapshot behind: nothing to launch then)
                s = self._system()
                h.check(h.lib.bl_snapshot(h.h, ctypes.byref(s), float(self._time), _lib.ptr(snap[1]),
                                          ctypes.c_void_p(snap[2].data_ptr()), h.stream()))
            block = snap[2].numpy()[:8 * n].reshape(4, n, 2).copy()
            self._pos, self._vel, self._p0, self._t0 = block[0], block[1], block[2], block[3]
        self._snap_fresh = False
        self._host_valid = True

    #: set by an evolve launch that left the host mirrors' snapshot in the pinned buffer (same stream, same wait)
    _snap_fresh = False
    _snap = None

    def _snap_buffers(self, n):
        """[4][n][2] doubles on the device and page-locked on the host, grown geometrically, cached per Billiard."""
        import torch
        snap = self._snap
        if snap is None or snap[0] < n:
            cap = max(n, 2 * (snap[0] if snap else 0))
            snap = (cap, torch.empty(8 * cap, dtype=torch.float64, device=self._h().torch_device),
                    torch.empty(8 * cap, dtype=torch.float64, pin_memory=True))
            self._snap = snap
            self._snap_fresh = False
        return snap

    # ------------------------------------------------------------------------------------------------
    # attributes of the reference class
    # ------------------------------------------------------------------------------------------------
    @property
    def time(self):
        """Current simulation time."""
        return self._time

    @time.setter
    def time(self, value):
        self._sync_host()   # mirrors are positions at the old time
        self._flush()       # queued rows must be solved at their add time
        self._time = value

    @property
    def count(self):
        """Number of balls in the billiard."""
        return self._n + self._n_pending

    def _get(self, name):
        self._sync_host()
        return getattr(self, name)

    balls_position = property(lambda self: self._get("_pos"), lambda self, v: self._set("_pos", v))
    balls_velocity = property(lambda self: self._get("_vel"), lambda self, v: self._set("_vel", v))
    balls_initial_time = property(lambda self: self._get("_t0"), lambda self, v: self._set("_t0", v))
    balls_initial_position = property(lambda self: self._get("_p0"), lambda self, v: self._set("_p0", v))

    def _set(self, name, value):
        self._sync_host()
        setattr(self, name, np.asarray(value, dtype=np.float64).reshape(self.count, 2))

    @property
    def toi_table(self):
        """Lower-triangular table of absolute impact times as a list of rows (snapshot of the device table)."""
        self._flush()
        n = self._n
        if n == 0:
            return []
        self._table_sync()
        tab = self._dev["toi"][:n, :n].cpu().numpy()
        return [tab[i, :i].copy() for i in range(n)]

    def _row_minima(self):
        import torch
        self._flush()
        self._table_sync()
        h, n = self._h(), self._n
        bt = torch.empty(max(n, 1), dtype=torch.float64, device=h.torch_device)
        bi = torch.empty(max(n, 1), dtype=torch.int64, device=h.torch_device)
        s = self._system()
        h.check(h.lib.bl_row_minima(h.h, ctypes.byref(s), _lib.ptr(bt), _lib.ptr(bi), h.stream()))
        return bt[:n].cpu().numpy(), bi[:n].cpu().numpy()

    @property
    def _balls_toi(self):
        return self._row_minima()[0]

    @property
    def _balls_idx(self):
        return [_I64(x) for x in self._row_minima()[1]]

    @property
    def _obstacles_toi(self):
        self._flush()
        return self._dev["obs_t"][:self._n].cpu().numpy() if self._n else np.empty((0,), dtype=np.float64)

    def _obs_and_args(self, idx, arg):
        if idx < 0:
            return None
        obs = self.obstacles[idx]
        if self._obs_records[idx].kind == _lib.BL_OBS_WALL:
            return (obs, (np.float64(arg),))
        if self._obs_records[idx].kind == _lib.BL_OBS_SEGMENT:
            return (obs, obs._args(arg))
        return (obs, ())

    @property
    def _obstacles_obs(self):
        self._flush()
        if not self._n:
            return []
        idx = self._dev["obs_idx"][:self._n].cpu().tolist()
        arg = self._dev["obs_arg"][:self._n].cpu().tolist()
        return [self._obs_and_args(i, a) for i, a in zip(idx, arg)]

    def _query_next(self):
        self._flush()
        if self._next is None:
            h = self._h()
            res = _lib.BlResult()
            s = self._system()
            h.check(h.lib.bl_next(h.h, ctypes.byref(s), float(self._time), ctypes.byref(res), h.stream()))
            self._next = res
        return self._next

    @property
    def _next_ball_ball_collision(self):
        r = self._query_next()
        return (r.bb_t, _I64(r.bb_i), _I64(r.bb_j))

    @property
    def _next_ball_obstacle_collision(self):
        r = self._query_next()
        return (r.bo_t, _I64(r.bo_ball), self._obs_and_args(int(r.bo_obs), r.bo_arg))

    @property
    def next_ball_ball_collision(self):
        """Next ball-ball collision as a (time, ball index, ball index)-triplet."""
        r = self._query_next()
        return (float(r.bb_t), int(r.bb_i), int(r.bb_j))

    @property
    def next_ball_obstacle_collision(self):
        """Next ball-obstacle collision as (time, ball index, (obstacle, args))."""
        r = self._query_next()
        return (float(r.bo_t), int(r.bo_ball), self._obs_and_args(int(r.bo_obs), r.bo_arg))

    @property
    def next_collision(self):
        """The earlier of the two; a ball-ball collision wins a tie (simulation.py:131-139)."""
        r = self._query_next()
        if r.bb_t <= r.bo_t:
            return self.next_ball_ball_collision
        return self.next_ball_obstacle_collision

    # ------------------------------------------------------------------------------------------------
    # building the table
    # ------------------------------------------------------------------------------------------------
    def add_ball(self, pos, vel=(0, 0), radius=0.0, mass=1.0):
        """Add a ball at ``pos`` with velocity ``vel``; returns its index (simulation.py:141-224).

        Zero radius = point particle (two of them never collide); zero mass = pushes nobody; infinite mass =
        pushed by nobody.
        """
        p = np.asarray(pos, dtype=np.float64)
        v = np.asarray(vel, dtype=np.float64)
        if p.shape != (2,) or v.shape != (2,):
            raise ValueError(f"position and velocity must be 2D vectors, got shapes {p.shape} and {v.shape}")
        radius, mass = float(radius), float(mass)  # TypeError for None, like the reference
        self._sync_host()
        idx = self.count
        self._pending.append((p.reshape(1, 2).copy(), v.reshape(1, 2).copy(), self._time))
        self._n_pending += 1
        self.balls_radius.append(radius)
        self.balls_mass.append(mass)
        self._pos = np.append(self._pos, [p], axis=0)
        self._vel = np.append(self._vel, [v], axis=0)
        self._p0 = np.append(self._p0, [p], axis=0)
        self._t0 = np.append(self._t0, [(self._time, self._time)], axis=0)
        self._next = None
        return idx

    def recompute_toi(self, indices=None):
        """Recompute the time-of-impact entries of the given ball(s) after editing ``balls_position``,
        ``balls_velocity`` or ``balls_radius`` in place (simulation.py:226-309).  ``None`` = all balls."""
        import torch
        if indices is None:
            idx = list(range(self.count))
        elif isinstance(indices, int):
            idx = [indices]
        else:
            idx = [int(i) for i in iter(indices)]  # TypeError if not iterable
        self._flush()
        self._table_sync()
        h, n = self._h(), self._n
        for i in idx:
            if not -n <= i < n:
                raise IndexError(f"ball index {i} out of range")
        idx = [i % n for i in idx] if n else []
        s = self._system()
        if self._host_valid and n:
            # the mirrors may carry edits: every ball takes the mirrored velocity, edited positions re-base the ball
            dev = h.torch_device
            d_pos = torch.from_numpy(np.ascontiguousarray(self._pos, dtype=np.float64)).to(dev)
            d_vel = torch.from_numpy(np.ascontiguousarray(self._vel, dtype=np.float64)).to(dev)
            h.check(h.lib.bl_apply_edits(h.h, ctypes.byref(s), float(self._time), _lib.ptr(d_pos), _lib.ptr(d_vel),
                                         h.stream()))
        if self.balls_radius != self._radius_up:
            self._upload_radii(0, n)
            s = self._system()
        self._sync_masses()
        if idx:
            d_idx = torch.tensor(sorted(set(idx)), dtype=torch.int32, device=h.torch_device)
            h.check(h.lib.bl_recompute(h.h, ctypes.byref(s), float(self._time), _lib.ptr(d_idx), d_idx.numel(),
                                       h.stream()))
        h.check(h.lib.bl_rebuild_cover(h.h, ctypes.byref(s), h.stream()))
        self._next = None
        self._host_valid = False
        self._snap_fresh = False

    def _detect_ball_collision(self, idx1, idx2):
        """Absolute time of impact of two balls of the simulation (simulation.py:311-330)."""
        from .physics import toi_ball_ball_batch
        pos, vel = self.balls_position, self.balls_velocity
        r1, r2 = self.balls_radius[idx1], self.balls_radius[idx2]
        h = self._h()
        t = toi_ball_ball_batch([pos[idx1]], [vel[idx1]], r1, [pos[idx2]], [vel[idx2]], r2, device=h.device,
                                dot_fma=h.dot_fma)
        return self._time + float(t[0])

    # ------------------------------------------------------------------------------------------------
    # the event loop
    # ------------------------------------------------------------------------------------------------
    def _run(self, end_time, max_events=-1, first_kind=_lib.BL_ANY, log=None):
        if end_time != end_time:
            # the reference handles nothing for NaN (`t <= nan` is False) and then fails its own assert
            # (simulation.py:419, :435); on the device a NaN limit would mean "never stop"
            raise ValueError("end_time must not be NaN")
        self._flush()
        self._sync_masses()
        h = self._h()
        res = _lib.BlResult()
        s = self._system()
        self._snap_fresh = False
        if self._n:
            # evolve + the snapshot of the host mirrors at the clock it stops at, one wait for both: the frame loop of
            # visualize_*.animate reads balls_position / balls_velocity after every evolve call
            snap = self._snap_buffers(self._n)
            rc = h.lib.bl_evolve_snapshot(h.h, ctypes.byref(s), float(self._time), float(end_time), int(max_events),
                                          int(first_kind), ctypes.byref(log) if log is not None else None,
                                          ctypes.byref(res), _lib.ptr(snap[1]), ctypes.c_void_p(snap[2].data_ptr()),
                                          h.stream())
            self._snap_fresh = rc == _lib.BL_OK
        else:
            rc = h.lib.bl_evolve(h.h, ctypes.byref(s), float(self._time), float(end_time), int(max_events),
                                 int(first_kind), ctypes.byref(log) if log is not None else None, ctypes.byref(res),
                                 h.stream())
        self.last_result = res
        if res.n_ball_ball + res.n_ball_obstacle:
            self._table_dirty = True
        if rc == _lib.BL_OK or rc in (_lib.BL_ESEPARATING, _lib.BL_EDIVZERO, _lib.BL_ENAN):
            self._time = res.time
            self._host_valid = False
            self._next = res if rc == _lib.BL_OK else None
        if rc == _lib.BL_ESEPARATING:
            raise ValueError(f"Balls are not moving towards each other: pos * vel > 0 "
                             f"(balls {res.err_i} and {res.err_j} at t = {res.time})")
        if rc == _lib.BL_EDIVZERO:
            raise FloatingPointError(f"divide by zero encountered in the collision of balls {res.err_i} and "
                                     f"{res.err_j}: both have zero mass")
        h.check(rc)
        return res

    #: how ``evolve`` serves callbacks.  "replay" (default): the whole interval is ONE kernel launch that writes a
    #: device event log with the callback payloads; the log is copied back once and replayed on the host (``time`` is
    #: set to each collision's time while its callbacks run; the arrays ``balls_position`` / ``balls_velocity`` /
    #: ``toi_table`` read inside a callback show the END of the launch, the per-collision numbers are the callback's
    #: arguments).  "step": one launch per collision, so the billiard's arrays are current after every collision --
    #: for callbacks that inspect or edit the billiard itself.
    callback_mode = "replay"
    #: collisions logged per launch in replay mode (the launch stops when the log is full and is continued)
    callback_log_capacity = 1 << 16
    #: safety valve for ``evolve``: None = no limit, like the reference.  Some set-ups collide infinitely often in finite
    #: time (a zero-mass ball squeezed between two heavy ones, a moving infinite mass pressing a ball against a wall, a
    #: radius-0 ball on a LineSegment); ``evolve(end_time)`` then never returns -- in the reference a Ctrl-C away, here
    #: inside a kernel that cannot be interrupted.  With a limit, ``evolve`` raises RuntimeError after that many
    #: collisions in one call (the state is left at the last handled collision, bit-exact and resumable).
    evolve_event_limit = None

    def _log_buffers(self, cap):
        """Device event log (t, i, j, payload[12]) as ONE allocation plus a pinned host mirror, cached per Billiard."""
        import torch
        cur = getattr(self, "_logbuf", None)
        if cur is None or cur[0] < cap:
            dev = self._h().torch_device
            d = torch.zeros(15 * cap, dtype=torch.float64, device=dev)
            hbuf = torch.zeros(15 * cap, dtype=torch.float64, pin_memory=True)
            cur = (cap, d, hbuf)
            self._logbuf = cur
        cap, d, hbuf = cur
        log = _lib.BlLog(cap, d[0:].data_ptr(), d[cap:].data_ptr(), d[2 * cap:].data_ptr(), d[3 * cap:].data_ptr())
        return cap, d, hbuf, log

    def _read_log(self, k, cap, d, hbuf, payload=True):
        """first k records of the device log -> numpy (t, i, j, payload) through the pinned mirror: one sync"""
        import torch
        if k == 0:
            return (np.empty(0), np.empty(0, dtype=_I64), np.empty(0, dtype=_I64), np.empty((0, 12)))
        for q in range(3):
            hbuf[q * cap:q * cap + k].copy_(d[q * cap:q * cap + k], non_blocking=True)
        if payload:
            hbuf[3 * cap:3 * cap + 12 * k].copy_(d[3 * cap:3 * cap + 12 * k], non_blocking=True)
        torch.cuda.current_stream(self._h().torch_device).synchronize()
        hn = hbuf.numpy()
        return (hn[0:k].copy(), hn[cap:cap + k].view(_I64).copy(), hn[2 * cap:2 * cap + k].view(_I64).copy(),
                hn[3 * cap:3 * cap + 12 * k].reshape(k, 12).copy() if payload else None)

    def _call_back(self, t, i, j, p, ball_callbacks, obstacle_callbacks):
        """the callbacks of one logged collision, in the reference's order (simulation.py:590-594, :626-631)"""
        if j >= 0:
            if ball_callbacks:
                if i in ball_callbacks:
                    ball_callbacks[i](t, p[0:2].copy(), p[2:4].copy(), p[4:6].copy(), j)
                if j in ball_callbacks:
                    ball_callbacks[j](t, p[6:8].copy(), p[8:10].copy(), p[10:12].copy(), i)
        else:
            obs, args = self._obs_and_args(-1 - j, p[6])
            if ball_callbacks and i in ball_callbacks:
                ball_callbacks[i](t, p[0:2].copy(), p[2:4].copy(), p[4:6].copy(), obs)
            if obstacle_callbacks and obs in obstacle_callbacks:
                obstacle_callbacks[obs](t, p[0:2].copy(), p[2:4].copy(), p[4:6].copy(), args, i)

    def evolve(self, end_time, time_callback=None, ball_callbacks=None, obstacle_callbacks=None):
        """Advance the simulation to ``end_time``; returns ``(ball-ball, ball-obstacle)`` collision counts
        (simulation.py:354-438).

        Without callbacks the whole interval is ONE kernel launch.  With callbacks it still is (``callback_mode``
        "replay"): the kernel logs every collision with its payload, the host replays the log --
        ``ball_callbacks[i](t, pos, vel_before, vel_after, other ball or obstacle)``,
        ``obstacle_callbacks[obs](t, pos, vel_before, vel_after, args, ball)``, then ``time_callback(t)``, in the
        reference's order.  Differences from the reference, by design: callbacks run after the launch, so the
        billiard's own arrays read inside a callback show the end of the launch (``callback_mode = "step"`` gives one
        launch per collision and current arrays; in both modes the velocities are already the post-collision ones,
        whereas the reference calls back just before it stores them).
        """
        if ball_callbacks is not None:
            if not isinstance(ball_callbacks, Mapping):
                raise TypeError(f"Argument 'ball_callbacks' must be a mapping, not a {type(ball_callbacks)}")
            if any(not isinstance(key, int) for key in ball_callbacks):
                raise ValueError("Keys of 'ball_callbacks' must be integers")
        if obstacle_callbacks is not None:
            if not isinstance(obstacle_callbacks, Mapping):
                raise TypeError(f"Argument 'obstacle_callbacks' must be a mapping, not a {type(obstacle_callbacks)}")
            if any(not isinstance(key, Obstacle) for key in obstacle_callbacks):
                raise ValueError("Keys of 'obstacle_callbacks' must be instances of Obstacle")

        limit = -1 if self.evolve_event_limit is None else int(self.evolve_event_limit)
        if time_callback is None and not ball_callbacks and not obstacle_callbacks:
            res = self._run(end_time, max_events=limit)
            if res.stop == _lib.BL_STOP_TIME:
                self._time = end_time  # keep the caller's number (the reference stores end_time itself)
            elif res.stop == _lib.BL_STOP_EVENTS:
                raise RuntimeError(f"evolve_event_limit: {limit} collisions handled and t = {self._time} < end_time = "
                                   f"{end_time}; collisions accumulating in finite time?")
            return (int(res.n_ball_ball), int(res.n_ball_obstacle))

        n_bb = n_bo = 0
        if self.callback_mode == "step":
            while True:
                res = self._step(end_time, _lib.BL_ANY, ball_callbacks, obstacle_callbacks)
                n_bb += int(res.n_ball_ball)
                n_bo += int(res.n_ball_obstacle)
                if res.n_ball_ball + res.n_ball_obstacle == 0:
                    break
                if time_callback is not None:
                    time_callback(self._time)
            self._time = end_time
            return (n_bb, n_bo)

        want_payload = bool(ball_callbacks) or bool(obstacle_callbacks)
        self._flush()
        cap, d, hbuf, log = self._log_buffers(int(self.callback_log_capacity))
        if not want_payload:
            log = _lib.BlLog(cap, log.t, log.i, log.j, 0)
        while True:
            res = self._run(end_time, log=log)
            n_bb += int(res.n_ball_ball)
            n_bo += int(res.n_ball_obstacle)
            t, i, j, p = self._read_log(int(res.n_logged), cap, d, hbuf, want_payload)
            after = self._time
            try:
                for e in range(len(t)):
                    te = float(t[e])
                    self._time = te  # the reference's `time` while the callbacks of this collision run
                    if want_payload:
                        self._call_back(te, int(i[e]), int(j[e]), p[e], ball_callbacks, obstacle_callbacks)
                    if time_callback is not None:
                        time_callback(te)
            finally:
                self._time = after
            if res.stop != _lib.BL_STOP_LOG:
                break
        if res.stop == _lib.BL_STOP_TIME:
            self._time = end_time
        return (n_bb, n_bo)

    def _step(self, end_time, kind, ball_callbacks, obstacle_callbacks):
        """One collision with its callback payload (the device logs pos / v_before / v_after of the event); the
        record comes back through one pinned block."""
        want = bool(ball_callbacks) or bool(obstacle_callbacks)
        log = None
        if want:
            self._flush()
            cap, d, hbuf, log = self._log_buffers(1)
        res = self._run(end_time, max_events=1, first_kind=kind, log=log)
        if want and res.n_logged:
            t, i, j, p = self._read_log(1, cap, d, hbuf)
            self._call_back(float(t[0]), int(i[0]), int(j[0]), p[0], ball_callbacks, obstacle_callbacks)
        return res

    def bounce_ball_ball(self, ball_callbacks=None):
        """Advance to the next ball-ball collision and handle it (simulation.py:440-502)."""
        if isinf(self._query_next().bb_t):
            raise RuntimeError("there is no ball-ball collision to advance to")
        self._step(INF, _lib.BL_BALL_BALL, ball_callbacks, None)

    def bounce_ball_obstacle(self, ball_callbacks=None, obstacle_callbacks=None):
        """Advance to the next ball-obstacle collision and handle it (simulation.py:504-554)."""
        if isinf(self._query_next().bo_t):
            raise RuntimeError("there is no ball-obstacle collision to advance to")
        self._step(INF, _lib.BL_BALL_OBSTACLE, ball_callbacks, obstacle_callbacks)

    # ------------------------------------------------------------------------------------------------
    # extensions used by the benchmarks and parity tests
    # ------------------------------------------------------------------------------------------------
    def add_balls(self, pos, vel, radius=0.0, mass=1.0):
        """Queue many balls at once (same semantics as repeated ``add_ball``)."""
        pos = np.asarray(pos, dtype=np.float64).reshape(-1, 2)
        vel = np.asarray(vel, dtype=np.float64).reshape(-1, 2)
        k = pos.shape[0]
        radius = np.broadcast_to(np.asarray(radius, dtype=np.float64), (k,)).tolist()
        mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (k,)).tolist()
        self._sync_host()
        first = self.count
        t = self._time
        if k:
            self._pending.append((pos.copy(), vel.copy(), t))
            self._n_pending += k
        self.balls_radius.extend(radius)
        self.balls_mass.extend(mass)
        self._pos = np.concatenate([self._pos, pos], axis=0)
        self._vel = np.concatenate([self._vel, vel], axis=0)
        self._p0 = np.concatenate([self._p0, pos], axis=0)
        self._t0 = np.concatenate([self._t0, np.full((k, 2), t)], axis=0)
        self._next = None
        return range(first, first + k)

    def evolve_events(self, max_events, end_time=INF, log_capacity=0, payload=False):
        """Process at most ``max_events`` collisions (count-based stop, Notes.md:11 of the reference) in one launch.

        Returns ``(n_ball_ball, n_ball_obstacle)``; with ``log_capacity`` > 0 also the device event log as
        ``(t, i, j[, payload])`` numpy arrays (j = -1-obstacle_index for obstacle hits)."""
        import torch
        if not log_capacity:
            res = self._run(end_time, max_events=max_events)
            return (int(res.n_ball_ball), int(res.n_ball_obstacle))
        dev = self._h().torch_device
        cap = int(log_capacity)
        buf_t = torch.zeros(cap, dtype=torch.float64, device=dev)
        buf_i = torch.zeros(cap, dtype=torch.int64, device=dev)
        buf_j = torch.zeros(cap, dtype=torch.int64, device=dev)
        buf_p = torch.zeros((cap, 12), dtype=torch.float64, device=dev) if payload else None
        log = _lib.BlLog(cap, buf_t.data_ptr(), buf_i.data_ptr(), buf_j.data_ptr(),
                         buf_p.data_ptr() if payload else 0)
        res = self._run(end_time, max_events=max_events, log=log)
        k = int(res.n_logged)
        out = [buf_t[:k].cpu().numpy(), buf_i[:k].cpu().numpy(), buf_j[:k].cpu().numpy()]
        if payload:
            out.append(buf_p[:k].cpu().numpy())
        return (int(res.n_ball_ball), int(res.n_ball_obstacle)), tuple(out)
