"""Static obstacles -- device-backed mirror of the reference's ``billiards.obstacles``.

``Obstacle`` keeps the reference's two-method interface (obstacles.py:21-59); ``Disk`` (:62-76),
``InfiniteWall`` (:79-144) and ``LineSegment`` (:147-198) are the three kinds the CUDA event loop knows.  Their
``detect_collision`` / ``resolve_collision`` methods run the same device functions the event kernel inlines
(``bl_obstacle_detect`` / ``bl_obstacle_resolve``), one problem per call.

Not supported ("Known differences"): user-defined ``Obstacle`` subclasses -- there is no host-side
collision loop to run their Python methods in, ``Billiard()`` raises ``NotImplementedError`` for them -- and edits of
an obstacle's attributes after the ``Billiard`` was built (the device record is taken once, at construction).
"""

from math import sqrt

import numpy as np

from . import _lib
from .physics import INF, _raise_elastic, _vec2  # noqa: F401


class Obstacle:
    """Base class: subclasses provide ``detect_collision(pos, vel, radius) -> (t, args)`` and
    ``resolve_collision(pos, vel, radius, *args) -> new velocity``.

    Only obstacles that can describe themselves to the device (``_device_record``) can be used in a
    ``Billiard``; there is no host-side collision loop to fall back to.
    """

    def detect_collision(self, pos, vel, radius):
        raise NotImplementedError("Subclasses should implement this!")

    def resolve_collision(self, pos, vel, radius, *args):
        raise NotImplementedError("Subclasses should implement this!")

    def _device_record(self):
        raise NotImplementedError(
            f"{type(self).__name__} has no CUDA implementation: the B200 event loop supports Disk, InfiniteWall and "
            "LineSegment (there is no host-side loop that could call a Python detect_collision per event)")

    # -- shared device calls -------------------------------------------------------------------------
    def _detect_on_device(self, pos, vel, radius, rsum2):
        import torch
        h = _lib.get_handle()
        p, v = _vec2(pos, "pos"), _vec2(vel, "vel")
        d_in = torch.from_numpy(np.concatenate([p, v, [float(radius)]])).to(h.torch_device)
        d_rs = torch.tensor([float(rsum2)], dtype=torch.float64, device=h.torch_device)
        d_out = torch.empty(2, dtype=torch.float64, device=h.torch_device)
        rec = self._device_record()
        h.check(h.lib.bl_obstacle_detect(h.h, rec, _lib.ptr(d_in), _lib.ptr(d_rs), _lib.ptr(d_out[:1]),
                                         _lib.ptr(d_out[1:]), 1, h.stream()))
        t, arg = d_out.cpu().tolist()
        return t, arg

    def _resolve_on_device(self, pos, vel, radius, arg):
        import torch
        h = _lib.get_handle()
        p, v = _vec2(pos, "pos"), _vec2(vel, "vel")
        d_in = torch.from_numpy(np.concatenate([p, v, [float(radius)]])).to(h.torch_device)
        d_arg = torch.tensor([float(arg)], dtype=torch.float64, device=h.torch_device)
        d_out = torch.empty(2, dtype=torch.float64, device=h.torch_device)
        d_st = torch.empty(1, dtype=torch.int32, device=h.torch_device)
        rec = self._device_record()
        h.check(h.lib.bl_obstacle_resolve(h.h, rec, _lib.ptr(d_in), _lib.ptr(d_arg), _lib.ptr(d_out), _lib.ptr(d_st),
                                          1, h.stream()))
        _raise_elastic(int(d_st.item()))
        return d_out.cpu().numpy()


class Disk(Obstacle):
    """A circular obstacle; balls stay outside (obstacles.py:62-76)."""

    def __init__(self, center, radius):
        self.center = np.asarray(center)
        self.radius = radius

    def _device_record(self):
        return _lib.BlObstacle(_lib.BL_OBS_DISK, 0, float(self.center[0]), float(self.center[1]), float(self.radius), 0.0)

    def detect_collision(self, pos, vel, radius):
        """Time until the ball touches the disk: a ball-ball problem against a static ball of the disk's radius."""
        t, _ = self._detect_on_device(pos, vel, radius, (self.radius + radius) ** 2)
        return t, ()

    def resolve_collision(self, pos, vel, radius, *args):
        """Reflect the ball off the disk (elastic collision with an immovable ball)."""
        return self._resolve_on_device(pos, vel, radius, 0.0)


class InfiniteWall(Obstacle):
    """An infinite one-sided wall through two points (obstacles.py:79-144).

    Walking from ``start_point`` to ``end_point`` the balls live on the side named by ``exterior``
    ("left" by default): balls arriving from that side bounce, balls crossing from the other side pass.
    """

    def __init__(self, start_point, end_point, exterior="left"):
        self.start_point = np.asarray(start_point)
        self.end_point = np.asarray(end_point)
        delta = self.end_point - self.start_point
        if delta[0] == 0.0 and delta[1] == 0.0:
            raise ValueError("this is not a line")
        if exterior not in ("left", "right"):
            raise ValueError(f'exterior must be "left" or "right", not {exterior}')
        # unit normal pointing to the left of the direction of travel; must round exactly like the reference's
        # constructor (normal / np.linalg.norm(normal), then negated for "right") because it feeds the kernels
        left = np.asarray([-delta[1], delta[0]])
        unit = left / np.linalg.norm(left)
        self._normal = -unit if exterior == "right" else unit

    def _device_record(self):
        return _lib.BlObstacle(_lib.BL_OBS_WALL, 0, float(self.start_point[0]), float(self.start_point[1]),
                               float(self._normal[0]), float(self._normal[1]))

    def detect_collision(self, pos, vel, radius):
        """Time until the ball's perimeter reaches the wall, and the closing speed ("headway")."""
        t, headway = self._detect_on_device(pos, vel, radius, 0.0)
        if t == INF:
            return INF, ()
        return t, (np.float64(headway),)

    def resolve_collision(self, pos, vel, radius, headway):
        """Mirror the velocity component normal to the wall."""
        assert headway > 0  # a colliding ball cannot be moving away from the wall
        return self._resolve_on_device(pos, vel, radius, headway)


class LineSegment(Obstacle):
    """A line segment that balls bounce off from both sides and from its two end points (obstacles.py:147-198)."""

    def __init__(self, start_point, end_point):
        self.start_point = np.asarray(start_point)
        self.end_point = np.asarray(end_point)
        direction = self.end_point - self.start_point
        length_sqrd = direction.dot(direction)
        if length_sqrd == 0.0:
            raise ValueError("this is not a line")
        # the two derived vectors feed the kernels, so they are formed with the same numpy operations as in the
        # reference's constructor: direction / |direction|^2 and the left-hand normal / |direction|
        self._covector = direction / length_sqrd
        self._normal = np.array([-direction[1], direction[0]]) / sqrt(length_sqrd)

    def _device_record(self):
        return _lib.BlObstacle(_lib.BL_OBS_SEGMENT, 0, float(self.start_point[0]), float(self.start_point[1]),
                               float(self._covector[0]), float(self._covector[1]), float(self._normal[0]),
                               float(self._normal[1]), float(self.end_point[0]), float(self.end_point[1]))

    @staticmethod
    def _args(u):
        """the args tuple of the reference: (0,) / (1,) for the end points, (float u,) for the line part"""
        return (int(u),) if u in (0.0, 1.0) else (np.float64(u),)

    def detect_collision(self, pos, vel, radius):
        """Time until the ball touches the segment or one of its end points, and where: (t, (u,))."""
        t, u = self._detect_on_device(pos, vel, radius, radius ** 2)
        if t == INF:
            return INF, (None,)
        return t, self._args(u)

    def resolve_collision(self, pos, vel, radius, u):
        """Velocity after the bounce: elastic collision with an end point for u = 0 / 1, mirror image otherwise."""
        return self._resolve_on_device(pos, vel, radius, float(u))
