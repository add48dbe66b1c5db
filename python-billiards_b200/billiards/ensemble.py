"""``BilliardEnsemble``: many independent single-ball billiards sharing one obstacle list.

The reference has no ensemble type -- ``examples/sinai_billiard.py:33-38`` fakes one by putting non-interacting
point particles into a single ``Billiard`` (and ``Notes.md:14`` wishes for a ``ParticleBilliard``).  Here every
system is exactly "a ``Billiard`` with one ball" stepped with ``bounce_ball_obstacle()``
(simulation.py:504-554), one CUDA thread per system (``bl_ensemble_run``).  Systems are independent, so an
ensemble shards over GPUs with no exchange during the run; ``torch.distributed`` (NCCL over NVLink) is used only
to gather the final statistics.
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


class BilliardEnsemble:
    """n independent systems: obstacles + ONE ball each (same radius for all), all starting at time 0.

    Args:
        obstacles: list of ``Disk`` / ``InfiniteWall`` (list order breaks ties, as in ``Billiard``).
        pos, vel: arrays (n, 2) -- initial position and velocity of each system's ball.
        radius: ball radius (0 = point particle).
        device, dot_fma: as for ``Billiard``.
        count_hits: keep per-obstacle hit counters.
    """

    def __init__(self, obstacles, pos, vel, radius=0.0, device=None, dot_fma=None, count_hits=False):
        import torch
        self.obstacles = list(obstacles)
        for obs in self.obstacles:
            if not isinstance(obs, Obstacle):
                raise TypeError(f"{obs} must be an obstacle.Obstacle instance")
        self._records = [o._device_record() for o in self.obstacles]
        self._rec_array = (_lib.BlObstacle * max(len(self._records), 1))(*self._records)
        self.radius = float(radius)
        # (R + r)**2 with Python floats, like Disk.detect_collision -> toi_ball_ball does (physics.py:52)
        self._rsum2 = np.asarray([float((o.radius + self.radius) ** 2) if r.kind == _lib.BL_OBS_DISK else
                                  float(self.radius ** 2) if r.kind == _lib.BL_OBS_SEGMENT else 0.0
                                  for o, r in zip(self.obstacles, self._records)] or [0.0], dtype=np.float64)
        # numpy arrays or (pinned) torch CPU tensors of shape (n, 2); the H2D copy is asynchronous for pinned input
        pos_t = pos if isinstance(pos, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float64))
        vel_t = vel if isinstance(vel, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(vel, dtype=np.float64))
        pos_t, vel_t = pos_t.reshape(-1, 2), vel_t.reshape(-1, 2)
        if pos_t.shape != vel_t.shape or pos_t.dtype != torch.float64 or vel_t.dtype != torch.float64:
            raise ValueError(f"pos and vel must be float64 arrays of the same shape, got {tuple(pos_t.shape)} and "
                             f"{tuple(vel_t.shape)}")
        self.n = pos_t.shape[0]
        self._handle = _lib.get_handle(device, dot_fma)
        dev = self._handle.torch_device
        self._state = torch.zeros((5, self.n), dtype=torch.float64, device=dev)  # SoA: p0x p0y vx vy t0
        self._state[0:2].copy_(pos_t.to(dev, non_blocking=True).T)
        self._state[2:4].copy_(vel_t.to(dev, non_blocking=True).T)
        self._time = torch.zeros(self.n, dtype=torch.float64, device=dev)
        self._counts = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self._status = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self._hits = torch.zeros((len(self.obstacles), self.n), dtype=torch.int32, device=dev) if count_hits else None

    # -- stepping ------------------------------------------------------------------------------------
    def run(self, n_collisions, end_time=INF, sync=True):
        """Every system handles its next ``n_collisions`` collisions (or stops at ``end_time``)."""
        import torch
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

    # -- read-out ------------------------------------------------------------------------------------
    @property
    def time(self):
        """time of the last collision of each system"""
        return _lib.to_host(self._time)

    @property
    def counts(self):
        return _lib.to_host(self._counts)

    @property
    def hits(self):
        return None if self._hits is None else _lib.to_host(self._hits).T

    @property
    def balls_initial_position(self):
        return _lib.to_host(self._state[0:2].T.contiguous())  # transposed on the device, one copy of what is asked for

    @property
    def balls_velocity(self):
        return _lib.to_host(self._state[2:4].T.contiguous())

    def positions_at(self, time):
        """ball positions at a common ``time`` (p0 + v*(time - t0), valid until each system's next collision)"""
        s = _lib.to_host(self._state)
        dt = float(time) - s[4]
        return np.stack([s[0] + s[2] * dt, s[1] + s[3] * dt], axis=1)

    def stats(self):
        """[total collisions, hits per obstacle...] summed over this process's systems (device reduction)."""
        import torch
        h = self._handle
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
