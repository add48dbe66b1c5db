"""Collision detection and handling -- device-backed mirror of the reference's ``billiards.physics``.

Same names, argument meaning and error behaviour as /root/reference/src/billiards/physics.py, but the
arithmetic runs in the sm_100a kernels of ``libbilliards_b200.so`` (``bl_toi_ball_ball``,
``bl_elastic_collision``); the scalar functions are batches of one.  The ``*_batch`` variants take arrays
of problems and are what a throughput-minded caller should use.
"""

import numpy as np

from . import _lib

INF = float("inf")


def _vec2(x, name):
    a = np.asarray(x, dtype=np.float64)
    if a.shape != (2,):
        raise ValueError(f"{name} must be a 2D vector, got shape {a.shape}")
    return a


def toi_ball_ball_batch(pos1, vel1, radius1, pos2, vel2, radius2, t_eps=-1e-10, device=None, dot_fma=None):
    """Times of impact for m pairs of moving balls (arrays of shape (m, 2) / (m,)).  physics.py:10-76."""
    import torch
    h = _lib.get_handle(device, dot_fma)
    p1, v1, p2, v2 = (np.asarray(x, dtype=np.float64).reshape(-1, 2) for x in (pos1, vel1, pos2, vel2))
    m = p1.shape[0]
    r1 = np.broadcast_to(np.asarray(radius1, dtype=np.float64), (m,))
    r2 = np.broadcast_to(np.asarray(radius2, dtype=np.float64), (m,))
    rows = np.concatenate([p1, v1, r1[:, None], p2, v2, r2[:, None]], axis=1)
    # (radius1 + radius2) ** 2 is evaluated by Python floats exactly like physics.py:52 does
    rsum2 = np.asarray([(a + b) ** 2 for a, b in zip(r1.tolist(), r2.tolist())], dtype=np.float64)
    d_in = torch.from_numpy(np.ascontiguousarray(rows)).to(h.torch_device)
    d_rs = torch.from_numpy(rsum2).to(h.torch_device)
    d_out = torch.empty(m, dtype=torch.float64, device=h.torch_device)
    h.check(h.lib.bl_toi_ball_ball(h.h, _lib.ptr(d_in), _lib.ptr(d_rs), float(t_eps), _lib.ptr(d_out), m, h.stream()))
    return d_out.cpu().numpy()


def toi_ball_ball(pos1, vel1, radius1, pos2, vel2, radius2, t_eps=-1e-10):
    """Time of impact of two moving balls, infinite if they do not collide (physics.py:10-76).

    Overlapping balls do not collide; ``t_eps`` (slightly negative) keeps grazing contacts that rounding would
    otherwise lose.
    """
    import torch
    h = _lib.get_handle()
    p1, v1, p2, v2 = _vec2(pos1, "pos1"), _vec2(vel1, "vel1"), _vec2(pos2, "pos2"), _vec2(vel2, "vel2")
    rsum2 = float((radius1 + radius2) ** 2)
    row = np.concatenate([p1, v1, [float(radius1)], p2, v2, [float(radius2)]])
    d_in = torch.from_numpy(row).to(h.torch_device)
    d_rs = torch.tensor([rsum2], dtype=torch.float64, device=h.torch_device)
    d_out = torch.empty(1, dtype=torch.float64, device=h.torch_device)
    h.check(h.lib.bl_toi_ball_ball(h.h, _lib.ptr(d_in), _lib.ptr(d_rs), float(t_eps), _lib.ptr(d_out), 1, h.stream()))
    return float(d_out.item())


def _raise_elastic(status, what=""):
    if status == _lib.BL_ESEPARATING:
        raise ValueError("Balls are not moving towards each other: pos * vel > 0" + what)
    if status == _lib.BL_EDIVZERO:
        # numpy raises/warns "divide by zero" for the reference's impulse when both masses vanish
        raise FloatingPointError("divide by zero encountered in elastic collision (total mass is zero)" + what)
    if status != _lib.BL_OK:
        raise _lib.BilliardsCudaError(status, "elastic_collision")


def elastic_collision_batch(pos1, vel1, mass1, pos2, vel2, mass2, device=None, dot_fma=None):
    """m elastic collisions; returns (vel1_after, vel2_after, status) with status[k] a bl_status (0 = ok)."""
    import torch
    h = _lib.get_handle(device, dot_fma)
    p1, v1, p2, v2 = (np.asarray(x, dtype=np.float64).reshape(-1, 2) for x in (pos1, vel1, pos2, vel2))
    m = p1.shape[0]
    m1 = np.broadcast_to(np.asarray(mass1, dtype=np.float64), (m,))
    m2 = np.broadcast_to(np.asarray(mass2, dtype=np.float64), (m,))
    rows = np.ascontiguousarray(np.concatenate([p1, v1, m1[:, None], p2, v2, m2[:, None]], axis=1))
    d_in = torch.from_numpy(rows).to(h.torch_device)
    d_out = torch.empty((m, 4), dtype=torch.float64, device=h.torch_device)
    d_st = torch.empty(m, dtype=torch.int32, device=h.torch_device)
    h.check(h.lib.bl_elastic_collision(h.h, _lib.ptr(d_in), _lib.ptr(d_out), _lib.ptr(d_st), m, h.stream()))
    out = d_out.cpu().numpy()
    return out[:, :2].copy(), out[:, 2:].copy(), d_st.cpu().numpy()


def elastic_collision(pos1, vel1, mass1, pos2, vel2, mass2):
    """Velocities after a perfectly elastic collision of two balls (physics.py:255-282).

    Raises ValueError when the balls are separating (pos*vel > 1e-15) and FloatingPointError when both masses
    are zero (the reference divides by zero there).
    """
    _vec2(pos1, "pos1"), _vec2(vel1, "vel1"), _vec2(pos2, "pos2"), _vec2(vel2, "vel2")
    w1, w2, st = elastic_collision_batch([pos1], [vel1], float(mass1), [pos2], [vel2], float(mass2))
    _raise_elastic(int(st[0]))
    return w1[0], w2[0]


def toi_ball_point(pos, vel, radius, point, t_eps=-1e-10):
    """Time of impact of a moving ball and a static point, infinite if there is none (physics.py:79-144)."""
    import torch
    h = _lib.get_handle()
    p, v, q = _vec2(pos, "pos"), _vec2(vel, "vel"), _vec2(point, "point")
    d_in = torch.from_numpy(np.concatenate([p, v, q])).to(h.torch_device)
    d_r2 = torch.tensor([float(radius ** 2)], dtype=torch.float64, device=h.torch_device)  # physics.py:120
    d_out = torch.empty(1, dtype=torch.float64, device=h.torch_device)
    h.check(h.lib.bl_toi_ball_point(h.h, _lib.ptr(d_in), _lib.ptr(d_r2), float(t_eps), _lib.ptr(d_out), 1, h.stream()))
    return float(d_out.item())


def toi_and_param_ball_segment(pos, vel, radius, line_start, covector, normal, t_eps=-1e-10):
    """Time of impact of a moving ball and an open line segment, with the line parameter of the hit
    (physics.py:147-252).

    Returns ``(t, u)``: ``t`` finite and ``u`` in [0, 1] for a hit on the segment; ``(inf, 0)`` / ``(inf, 1)`` when
    only the start / end point can still be hit (test it with ``toi_ball_point``); ``(inf, None)`` otherwise.
    """
    import torch
    h = _lib.get_handle()
    p, v = _vec2(pos, "pos"), _vec2(vel, "vel")
    s0, cv, nm = _vec2(line_start, "line_start"), _vec2(covector, "covector"), _vec2(normal, "normal")
    rec = _lib.BlObstacle(_lib.BL_OBS_SEGMENT, 0, s0[0], s0[1], cv[0], cv[1], nm[0], nm[1], 0.0, 0.0)
    d_in = torch.from_numpy(np.concatenate([p, v, [float(radius)]])).to(h.torch_device)
    d_out = torch.empty(2, dtype=torch.float64, device=h.torch_device)
    d_code = torch.empty(1, dtype=torch.int32, device=h.torch_device)
    h.check(h.lib.bl_toi_and_param_ball_segment(h.h, rec, _lib.ptr(d_in), float(t_eps), _lib.ptr(d_out[:1]),
                                                _lib.ptr(d_out[1:]), _lib.ptr(d_code), 1, h.stream()))
    t, u = d_out.cpu().tolist()
    code = int(d_code.item())
    return t, (np.float64(u) if code == 0 else 0 if code == 1 else 1 if code == 2 else None)
