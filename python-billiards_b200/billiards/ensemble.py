"""``BilliardEnsemble``: many independent billiards sharing one obstacle list.

The reference has no ensemble type -- ``examples/sinai_billiard.py:33-38`` fakes one by putting non-interacting
point particles into a single ``Billiard`` (and ``Notes.md:14`` wishes for a ``ParticleBilliard``).  Two shapes:

* **one ball per system** (``pos`` of shape ``(S, 2)``; config S of BASELINE.json): every system is exactly "a
  ``Billiard`` with one ball" stepped with ``bounce_ball_obstacle()`` (simulation.py:504-554), one CUDA thread per
  system (``bl_ensemble_run``);
* **n <= 224 balls per system** (``pos`` of shape ``(S, n, 2)``; e.g. 2^16 pool tables with different break shots; up to
  32 balls a group of lanes serves a system, from 33 balls one CTA with the table in shared memory): every
  system is exactly "a ``Billiard`` with these n balls" run with ``evolve`` (simulation.py:354-438), W = 2..32 lanes of
  a warp per system with the system's time-of-impact table in shared memory (``bl_multi_run``) -- the same kernel
  that serves a single small ``Billiard``, so system s of the ensemble and a ``Billiard`` built from the same numbers
  agree bit for bit.  Radii and masses are per ball slot, the same in every system.

Systems are independent, so an ensemble shards over GPUs with no exchange during the run (``shard_bounds``);
``torch.distributed`` (NCCL over NVLink) is used only to gather the final statistics.
"""

import ctypes

import numpy as np

from . import _lib
from .obstacles import Obstacle

INF = float("inf")


def shard_bounds(n_total, world_size, rank):
    """Contiguous block of systems owned by ``rank``: [lo, hi).  Results are partition-invariant."""
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


_pinned_cache = {}


def _pinned(name, shape, dtype):
    """page-locked host buffer, allocated once per (name, shape, dtype) and reused by later calls"""
    import torch
    key = (name, tuple(shape), dtype)
    buf = _pinned_cache.get(key)
    if buf is None:
        buf = torch.empty(shape, dtype=dtype, pin_memory=True)
        _pinned_cache[key] = buf
    return buf


def run_from_host(obstacles, pos, vel, n_collisions, end_time=INF, radius=0.0, mass=1.0, device=None, dot_fma=None,
                  chunks=4, count_hits=False):
    """Ensembles, host arrays in -> host arrays out, with the copies hidden behind the kernel.

    The systems are split into ``chunks`` contiguous blocks; block c goes host -> device, runs and comes back on its own
    CUDA stream, so the copy engines work on one block while the SMs run the collision loop of another (the kernels are
    bound by instruction issue / the FP64 pipe, the copies by PCIe).  Results are what
    ``BilliardEnsemble(...).run(n_collisions, end_time)`` gives (systems are independent: the partition does not matter).
    ``pos`` / ``vel``: ``(S, 2)`` (one ball per system) or ``(S, n, 2)`` (n <= 224 balls per system; ``radius`` and ``mass`` of shape (S, n) give every system its own; otherwise ``radius`` and ``mass``
    per ball slot), ideally pinned torch CPU tensors (numpy arrays are staged through a pinned buffer first).  Returns a
    dict of numpy arrays -- ``time`` (S,), ``position`` (of the last collision for one-ball systems, at each system's
    clock otherwise), ``velocity``, ``counts`` (S,), ``stats`` -- views of pinned buffers that the next call overwrites.
    """
    import torch
    pos_t, vel_t = _as_f64_tensor(pos), _as_f64_tensor(vel)
    multi = pos_t.dim() == 3
    if not multi:
        pos_t, vel_t = pos_t.reshape(-1, 2), vel_t.reshape(-1, 2)
    S = pos_t.shape[0]
    if not pos_t.is_pinned():
        pos_t = _pinned("in_pos", pos_t.shape, torch.float64).copy_(pos_t)
    if not vel_t.is_pinned():
        vel_t = _pinned("in_vel", vel_t.shape, torch.float64).copy_(vel_t)
    h = _lib.get_handle(device, dot_fma)
    dev = h.torch_device
    out_t = _pinned("out_t", (S,), torch.float64)
    out_p = _pinned("out_p", tuple(pos_t.shape), torch.float64)
    out_v = _pinned("out_v", tuple(pos_t.shape), torch.float64)
    out_c = _pinned("out_c", (S,), torch.int64)
    chunks = max(1, min(int(chunks), S))
    streams = _pinned_cache.setdefault(("streams", dev.index, chunks), [torch.cuda.Stream(dev) for _ in range(chunks)])
    parts = []
    start = torch.cuda.current_stream(dev)
    for c in range(chunks):
        lo, hi = shard_bounds(S, chunks, c)
        st = streams[c]
        st.wait_stream(start)
        with torch.cuda.stream(st):
            r_c = radius[lo:hi] if np.ndim(radius) == 2 else radius   # (per-system radii / masses follow their systems)
            m_c = mass[lo:hi] if np.ndim(mass) == 2 else mass
            ens = BilliardEnsemble(obstacles, pos_t[lo:hi], vel_t[lo:hi], radius=r_c, mass=m_c, device=dev.index,
                                   dot_fma=h.dot_fma, count_hits=count_hits)
            ens.run(n_collisions, end_time, sync=False)
            if multi:
                d = ens._dev
                out_t[lo:hi].copy_(d["time"], non_blocking=True)
                out_p[lo:hi].copy_(ens._positions_device(0.0, 1), non_blocking=True)
                out_v[lo:hi].copy_(torch.stack([d["vx"], d["vy"]], dim=-1), non_blocking=True)
                out_c[lo:hi].copy_(d["n_ball_ball"] + d["n_ball_obstacle"], non_blocking=True)
            else:
                out_t[lo:hi].copy_(ens._time, non_blocking=True)
                out_p[lo:hi].copy_(ens._state[0:2].T, non_blocking=True)
                out_v[lo:hi].copy_(ens._state[2:4].T, non_blocking=True)
                out_c[lo:hi].copy_(ens._counts, non_blocking=True)
            parts.append(ens)
    for st in streams:
        start.wait_stream(st)  # the caller's stream continues after every block is back
    total = torch.stack([p.stats() for p in parts]).sum(dim=0)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    total = total.cpu()  # synchronises the current stream, which waited for every chunk
    bad = sum(int(((p._dev["status"] if multi else p._status) != 0).sum().item()) for p in parts)
    if bad:
        raise ValueError(f"{bad} systems hit a separating/degenerate collision (status != 0)")
    return {"time": out_t.numpy(), "position": out_p.numpy(), "velocity": out_v.numpy(), "counts": out_c.numpy(),
            "stats": total.numpy()}


def _as_f64_tensor(x):
    import torch
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))


class BilliardEnsemble:
    """S independent systems: the obstacles + one ball each (``pos`` (S, 2)) or n <= 224 balls each (``pos`` (S, n, 2)),
    all starting at time 0.

    Args:
        obstacles: list of ``Disk`` / ``InfiniteWall`` / ``LineSegment`` (list order breaks ties, as in ``Billiard``).
        pos, vel: initial positions and velocities, numpy arrays or (pinned) torch CPU tensors.
        radius: ball radius (0 = point particle); per ball slot (n,) for multi-ball systems.
        mass: per ball slot (n,) for multi-ball systems (a single ball never needs one).
        device, dot_fma: as for ``Billiard``.
        count_hits: keep per-obstacle hit counters (one-ball systems).
    """

    def __init__(self, obstacles, pos, vel, radius=0.0, mass=1.0, device=None, dot_fma=None, count_hits=False):
        import torch
        self.obstacles = list(obstacles)
        for obs in self.obstacles:
            if not isinstance(obs, Obstacle):
                raise TypeError(f"{obs} must be an obstacle.Obstacle instance")
        self._records = [o._device_record() for o in self.obstacles]
        self._rec_array = (_lib.BlObstacle * max(len(self._records), 1))(*self._records)
        pos_t, vel_t = _as_f64_tensor(pos), _as_f64_tensor(vel)
        if pos_t.shape != vel_t.shape or pos_t.dtype != torch.float64 or vel_t.dtype != torch.float64 \
                or pos_t.dim() not in (2, 3) or pos_t.shape[-1] != 2:
            raise ValueError(f"pos and vel must be float64 arrays of the same shape (S, 2) or (S, n, 2), got "
                             f"{tuple(pos_t.shape)} and {tuple(vel_t.shape)}")
        self._handle = _lib.get_handle(device, dot_fma)
        self.n = pos_t.shape[0]
        self.balls_per_system = 1 if pos_t.dim() == 2 else int(pos_t.shape[1])
        self._multi = pos_t.dim() == 3
        if self._multi:
            self._init_multi(pos_t, vel_t, radius, mass)
            return
        dev = self._handle.torch_device
        self.radius = float(radius)
        # (R + r)**2 with Python floats, like Disk.detect_collision -> toi_ball_ball does (physics.py:52)
        self._rsum2 = np.asarray([float((o.radius + self.radius) ** 2) if r.kind == _lib.BL_OBS_DISK else
                                  float(self.radius ** 2) if r.kind == _lib.BL_OBS_SEGMENT else 0.0
                                  for o, r in zip(self.obstacles, self._records)] or [0.0], dtype=np.float64)
        self._state = torch.zeros((5, self.n), dtype=torch.float64, device=dev)  # SoA: p0x p0y vx vy t0
        self._state[0:2].copy_(pos_t.to(dev, non_blocking=True).T)
        self._state[2:4].copy_(vel_t.to(dev, non_blocking=True).T)
        self._time = torch.zeros(self.n, dtype=torch.float64, device=dev)
        self._counts = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self._status = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self._hits = torch.zeros((len(self.obstacles), self.n), dtype=torch.int32, device=dev) if count_hits else None

    # ================================================================================================
    # n balls per system
    # ================================================================================================
    def _init_multi(self, pos_t, vel_t, radius, mass):
        import torch
        h = self._handle
        dev = h.torch_device
        S, n = self.n, self.balls_per_system
        if not 1 <= n <= 224:
            raise ValueError(f"multi-ball ensembles hold 1..224 balls per system, got {n} "
                             "(larger systems are single Billiards: the grid-wide event kernel)")
        r_in, m_in = np.asarray(radius, dtype=np.float64), np.asarray(mass, dtype=np.float64)
        self._per_system = r_in.ndim == 2 or m_in.ndim == 2
        if self._per_system:
            return self._init_multi_per_system(pos_t, vel_t, np.broadcast_to(r_in, (S, n)), np.broadcast_to(m_in, (S, n)))
        self.balls_radius = [float(x) for x in np.broadcast_to(r_in, (n,))]
        self.balls_mass = [float(x) for x in np.broadcast_to(m_in, (n,))]
        classes, class_radii = {}, []
        for r in self.balls_radius:
            if r not in classes:
                classes[r] = len(class_radii)
                class_radii.append(r)
        disk_r = [float(o.radius) if rec.kind == _lib.BL_OBS_DISK else 0.0
                  for o, rec in zip(self.obstacles, self._records)]
        rs, rs_obs = _lib.pow2_table(class_radii, disk_r)  # libm pow, like float.__pow__ (physics.py:52)
        for q, rec in enumerate(self._records):
            if rec.kind == _lib.BL_OBS_WALL:
                rs_obs[q, :] = 0.0
        d = {}
        d["rsum2"] = torch.from_numpy(rs).to(dev)
        d["rsum2_obs"] = torch.from_numpy(np.ascontiguousarray(rs_obs).reshape(-1)).to(dev) if rs_obs.size \
            else torch.zeros(1, dtype=torch.float64, device=dev)
        d["obstacles"] = torch.from_numpy(np.frombuffer(bytes(self._rec_array), dtype=np.uint8).copy()).to(dev)
        d["radius"] = torch.tensor(self.balls_radius, dtype=torch.float64, device=dev)
        d["mass"] = torch.tensor(self.balls_mass, dtype=torch.float64, device=dev)
        d["rclass"] = torch.tensor([classes[r] for r in self.balls_radius], dtype=torch.int32, device=dev)
        self._n_classes = len(class_radii)
        self._alloc_multi_state(d, pos_t, vel_t)

    def _init_multi_per_system(self, pos_t, vel_t, radius, mass):
        """Every system its own radii and masses (``radius`` / ``mass`` of shape (S, n)): e.g. Galperin's billiard at many
        mass ratios at once.  Radius classes are per system and per ball slot (K = n); the (r_a + r_b)**2 tables come from
        the host's libm pow() system by system, like ``float ** 2`` in physics.py:52."""
        import torch
        dev = self._handle.torch_device
        S, n = self.n, self.balls_per_system
        self.balls_radius = np.array(radius, dtype=np.float64, order="C")   # (S, n), own writable copies
        self.balls_mass = np.array(mass, dtype=np.float64, order="C")
        disk_r = [float(o.radius) if rec.kind == _lib.BL_OBS_DISK else 0.0
                  for o, rec in zip(self.obstacles, self._records)]
        n_obs = len(self.obstacles)
        rs = np.empty((S, n, n), dtype=np.float64)
        rs_obs = np.zeros((S, max(n_obs, 1), n), dtype=np.float64)
        cache = {}
        for sys_ in range(S):
            key = self.balls_radius[sys_].tobytes()
            if key not in cache:
                a, b = _lib.pow2_table(self.balls_radius[sys_], disk_r)
                for q, rec in enumerate(self._records):
                    if rec.kind == _lib.BL_OBS_WALL:
                        b[q, :] = 0.0
                cache[key] = (a, b)
            rs[sys_] = cache[key][0]
            if n_obs:
                rs_obs[sys_, :n_obs] = cache[key][1]
        d = {}
        d["rsum2"] = torch.from_numpy(rs).to(dev)
        d["rsum2_obs"] = torch.from_numpy(np.ascontiguousarray(rs_obs[:, :max(n_obs, 1)]).reshape(-1)).to(dev)
        d["obstacles"] = torch.from_numpy(np.frombuffer(bytes(self._rec_array), dtype=np.uint8).copy()).to(dev)
        d["radius"] = torch.from_numpy(self.balls_radius).to(dev)
        d["mass"] = torch.from_numpy(self.balls_mass).to(dev)
        d["rclass"] = torch.arange(n, dtype=torch.int32, device=dev).repeat(S, 1).contiguous()
        self._n_classes = n
        self._alloc_multi_state(d, pos_t, vel_t)

    def _alloc_multi_state(self, d, pos_t, vel_t):
        import torch
        dev = self._handle.torch_device
        S, n = self.n, self.balls_per_system
        # state of record, SoA over (system, ball): one H2D copy each for pos and vel, de-interleaved on the device
        p = pos_t.to(dev, non_blocking=True)
        v = vel_t.to(dev, non_blocking=True)
        d["p0x"], d["p0y"] = p[..., 0].contiguous(), p[..., 1].contiguous()
        d["vx"], d["vy"] = v[..., 0].contiguous(), v[..., 1].contiguous()
        d["t0"] = torch.zeros((S, n), dtype=torch.float64, device=dev)
        d["toi"] = torch.empty((S, n, n), dtype=torch.float64, device=dev)
        d["obs_t"] = torch.empty((S, n), dtype=torch.float64, device=dev)
        d["obs_arg"] = torch.empty((S, n), dtype=torch.float64, device=dev)
        d["obs_idx"] = torch.empty((S, n), dtype=torch.int32, device=dev)
        d["time"] = torch.zeros(S, dtype=torch.float64, device=dev)
        d["n_ball_ball"] = torch.zeros(S, dtype=torch.int64, device=dev)
        d["n_ball_obstacle"] = torch.zeros(S, dtype=torch.int64, device=dev)
        d["status"] = torch.zeros(S, dtype=torch.int32, device=dev)
        self._dev = d
        self._fresh = True  # tables not built yet: the first run solves every pair at time 0, like add_ball
        self._hits = None

    def _multi_struct(self):
        m = _lib.BlMulti()
        m.n_sys, m.n, m.n_classes, m.n_obstacles = self.n, self.balls_per_system, self._n_classes, len(self.obstacles)
        m.per_system = 1 if self._per_system else 0
        for name in ("p0x", "p0y", "vx", "vy", "t0", "radius", "mass", "rclass", "rsum2", "rsum2_obs", "obstacles",
                     "toi", "obs_t", "obs_arg", "obs_idx", "time", "n_ball_ball", "n_ball_obstacle", "status"):
            setattr(m, name, self._dev[name].data_ptr())
        return m

    def _run_multi(self, n_collisions, end_time, sync):
        import torch
        h = self._handle
        m = self._multi_struct()
        h.check(h.lib.bl_multi_run(h.h, ctypes.byref(m), int(self._fresh), float(end_time), int(n_collisions),
                                   h.stream()))
        self._fresh = False
        if sync:
            torch.cuda.current_stream(h.torch_device).synchronize()
            st = self._dev["status"]
            bad = int((st != 0).sum().item())
            if bad:
                first = int(torch.nonzero(st)[0].item())
                code = int(st[first].item())
                what = {_lib.BL_ESEPARATING: "balls are not moving towards each other (ValueError in the reference)",
                        _lib.BL_EDIVZERO: "two zero-mass balls collided (divide by zero)",
                        _lib.BL_ENAN: "a NaN impact time"}.get(code, f"status {code}")
                raise ValueError(f"{bad} systems stopped early, first: system {first}: {what}")

    def evolve(self, end_time, max_collisions=None):
        """Every system runs ``Billiard.evolve(end_time)``; returns the (ball-ball, ball-obstacle) collision counts of
        this call per system as two int64 arrays.  ``max_collisions`` bounds the collisions of each system in this call
        (a safety valve: some set-ups collide infinitely often in finite time, and a kernel cannot be interrupted);
        a system that reaches it stops at its last collision, ``time`` tells which."""
        if not self._multi:
            raise TypeError("evolve(end_time) is for multi-ball ensembles; one-ball ensembles use run()")
        before_bb, before_bo = self._dev["n_ball_ball"].clone(), self._dev["n_ball_obstacle"].clone()
        self._run_multi(-1 if max_collisions is None else int(max_collisions), end_time, True)
        return (_lib.to_host(self._dev["n_ball_ball"] - before_bb), _lib.to_host(self._dev["n_ball_obstacle"] - before_bo))

    # ================================================================================================
    # stepping
    # ================================================================================================
    def run(self, n_collisions, end_time=INF, sync=True):
        """Every system handles its next ``n_collisions`` collisions (or stops at ``end_time``; ``n_collisions`` < 0 =
        no limit on the count)."""
        import torch
        if end_time != end_time:
            raise ValueError("end_time must not be NaN")
        if self._multi:
            return self._run_multi(n_collisions, end_time, sync)
        h = self._handle
        h.check(h.lib.bl_ensemble_run(
            h.h, self._rec_array, len(self._records), self._rsum2.ctypes.data_as(_lib.c_double_p), self.radius, self.n,
            _lib.ptr(self._state), _lib.ptr(self._time), int(n_collisions), float(end_time), _lib.ptr(self._counts),
            _lib.ptr(self._hits), _lib.ptr(self._status), h.stream()))
        if sync:
            torch.cuda.current_stream(h.torch_device).synchronize()
            bad = int((self._status != 0).sum().item())
            if bad:
                raise ValueError(f"{bad} systems hit a separating/degenerate collision (status != 0)")

    # ================================================================================================
    # read-out
    # ================================================================================================
    @property
    def time(self):
        """clock of each system: time of its last collision, or the end_time of the last evolve"""
        return _lib.to_host(self._dev["time"] if self._multi else self._time)

    @property
    def counts(self):
        """collisions handled so far, per system"""
        if self._multi:
            return _lib.to_host(self._dev["n_ball_ball"] + self._dev["n_ball_obstacle"])
        return _lib.to_host(self._counts)

    @property
    def counts_ball_ball(self):
        return _lib.to_host(self._dev["n_ball_ball"]) if self._multi else np.zeros(self.n, dtype=np.int64)

    @property
    def counts_ball_obstacle(self):
        return _lib.to_host(self._dev["n_ball_obstacle"]) if self._multi else self.counts

    @property
    def hits(self):
        return None if self._hits is None else _lib.to_host(self._hits).T

    @property
    def balls_initial_position(self):
        if self._multi:
            import torch
            return _lib.to_host(torch.stack([self._dev["p0x"], self._dev["p0y"]], dim=-1))
        return _lib.to_host(self._state[0:2].T.contiguous())  # transposed on the device, one copy of what is asked for

    @property
    def balls_initial_time(self):
        return _lib.to_host(self._dev["t0"]) if self._multi else _lib.to_host(self._state[4])

    @property
    def balls_velocity(self):
        if self._multi:
            import torch
            return _lib.to_host(torch.stack([self._dev["vx"], self._dev["vy"]], dim=-1))
        return _lib.to_host(self._state[2:4].T.contiguous())

    @property
    def balls_position(self):
        """positions at each system's own clock (multi-ball ensembles: what ``Billiard.balls_position`` shows)"""
        if not self._multi:
            s = _lib.to_host(self._state)
            t = _lib.to_host(self._time)
            dt = t - s[4]
            return np.stack([s[0] + s[2] * dt, s[1] + s[3] * dt], axis=1)
        return self._positions(0.0, 1)

    def _positions_device(self, time, at_system_time):
        import torch
        h = self._handle
        out = torch.empty((self.n, self.balls_per_system, 2), dtype=torch.float64, device=h.torch_device)
        m = self._multi_struct()
        h.check(h.lib.bl_multi_positions(h.h, ctypes.byref(m), float(time), int(at_system_time), _lib.ptr(out),
                                         h.stream()))
        return out

    def _positions(self, time, at_system_time):
        return _lib.to_host(self._positions_device(time, at_system_time))

    @property
    def toi_tables(self):
        """(S, n, n) absolute impact times of every system (multi-ball ensembles), both triangles, diagonal inf"""
        if not self._multi:
            raise TypeError("one-ball systems have no ball-ball table")
        if self._fresh:
            self._run_multi(0, INF, True)
        return _lib.to_host(self._dev["toi"])

    def positions_at(self, time):
        """ball positions at a common ``time`` (p0 + v*(time - t0), valid until each system's next collision)"""
        if self._multi:
            return self._positions(time, 0)
        s = _lib.to_host(self._state)
        dt = float(time) - s[4]
        return np.stack([s[0] + s[2] * dt, s[1] + s[3] * dt], axis=1)

    def stats(self):
        """Summed over this process's systems (device reduction): one-ball ensembles [total collisions, hits per
        obstacle...]; multi-ball ensembles [total collisions, ball-ball, ball-obstacle]."""
        import torch
        h = self._handle
        if self._multi:
            bb, bo = self._dev["n_ball_ball"].sum(), self._dev["n_ball_obstacle"].sum()
            return torch.stack([bb + bo, bb, bo])
        out = torch.zeros(1 + len(self.obstacles), dtype=torch.int64, device=h.torch_device)
        h.check(h.lib.bl_ensemble_stats(h.h, _lib.ptr(self._counts), _lib.ptr(self._hits), len(self.obstacles), self.n,
                                        _lib.ptr(out), h.stream()))
        return out

    def gather_stats(self):
        """Whole-job statistics: all-reduce of ``stats()`` over the default process group (NCCL on GPUs; the
        only collective of the ensemble path).  Without an initialised group this is just ``stats()``."""
        import torch.distributed as dist
        out = self.stats()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out.cpu().numpy()


BilliardEnsemble.run_from_host = staticmethod(run_from_host)
