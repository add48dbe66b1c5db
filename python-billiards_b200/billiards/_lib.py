"""ctypes binding of ``libbilliards_b200.so`` (C ABI declared in ``include/billiards_b200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, every
entry point raises.  PyTorch is used for device buffers and streams only.
"""

import ctypes
import os
import threading
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BILLIARDS_B200_LIB", os.path.join(_HERE, "libbilliards_b200.so"))

BL_OK = 0
BL_ECUDA, BL_EINVAL, BL_ESEPARATING, BL_EDIVZERO, BL_ENAN, BL_EUNSUPPORTED, BL_EASSERT = -1, -2, -3, -4, -5, -6, -7
BL_STOP_TIME, BL_STOP_EVENTS, BL_STOP_LOG, BL_STOP_ERROR = 0, 1, 2, 3
BL_ANY, BL_BALL_BALL, BL_BALL_OBSTACLE = 0, 1, 2
BL_OBS_WALL, BL_OBS_DISK, BL_OBS_SEGMENT = 0, 1, 2

c_double_p = ctypes.POINTER(ctypes.c_double)
c_void = ctypes.c_void_p


class BlObstacle(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("pad", ctypes.c_int32), ("a", ctypes.c_double), ("b", ctypes.c_double),
                ("c", ctypes.c_double), ("d", ctypes.c_double), ("e", ctypes.c_double), ("f", ctypes.c_double),
                ("g", ctypes.c_double), ("h", ctypes.c_double)]


class BlSystem(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("pitch", ctypes.c_int32), ("n_classes", ctypes.c_int32),
                ("n_obstacles", ctypes.c_int32),
                ("p0x", c_void), ("p0y", c_void), ("vx", c_void), ("vy", c_void), ("t0", c_void),
                ("radius", c_void), ("mass", c_void), ("rclass", c_void), ("rsum2", c_void), ("rsum2_obs", c_void),
                ("obstacles", c_void), ("toi", c_void), ("seq", c_void), ("seq_counter", c_void),
                ("cover_t", c_void), ("cover_code", c_void),
                ("obs_t", c_void), ("obs_idx", c_void), ("obs_arg", c_void)]


class BlResult(ctypes.Structure):
    _fields_ = [("time", ctypes.c_double), ("n_ball_ball", ctypes.c_int64), ("n_ball_obstacle", ctypes.c_int64),
                ("bb_t", ctypes.c_double), ("bb_i", ctypes.c_int64), ("bb_j", ctypes.c_int64),
                ("bo_t", ctypes.c_double), ("bo_ball", ctypes.c_int64), ("bo_obs", ctypes.c_int64),
                ("bo_arg", ctypes.c_double), ("status", ctypes.c_int32), ("stop", ctypes.c_int32),
                ("n_logged", ctypes.c_int64), ("n_rescans", ctypes.c_int64), ("err_i", ctypes.c_int64),
                ("err_j", ctypes.c_int64), ("prof", ctypes.c_int64 * 16), ("window", ctypes.c_double)]


class BlMulti(ctypes.Structure):
    _fields_ = [("n_sys", ctypes.c_int64), ("n", ctypes.c_int32), ("n_classes", ctypes.c_int32),
                ("n_obstacles", ctypes.c_int32), ("per_system", ctypes.c_int32),
                ("p0x", c_void), ("p0y", c_void), ("vx", c_void), ("vy", c_void), ("t0", c_void),
                ("radius", c_void), ("mass", c_void), ("rclass", c_void), ("rsum2", c_void), ("rsum2_obs", c_void),
                ("obstacles", c_void), ("toi", c_void), ("obs_t", c_void), ("obs_arg", c_void), ("obs_idx", c_void),
                ("time", c_void), ("n_ball_ball", c_void), ("n_ball_obstacle", c_void), ("status", c_void)]


class BlLog(ctypes.Structure):
    _fields_ = [("cap", ctypes.c_int64), ("t", c_void), ("i", c_void), ("j", c_void), ("payload", c_void)]


# every symbol include/billiards_b200.h declares: name -> (restype, argtypes)
_H = c_void
_SYS = ctypes.POINTER(BlSystem)
SIGNATURES = {
    "bl_create": (ctypes.c_int, [ctypes.POINTER(c_void), ctypes.c_int, ctypes.c_int]),
    "bl_destroy": (ctypes.c_int, [_H]),
    "bl_last_error": (ctypes.c_char_p, [_H]),
    "bl_version": (ctypes.c_char_p, []),
    "bl_recompute": (ctypes.c_int, [_H, _SYS, ctypes.c_double, c_void, ctypes.c_int, c_void]),
    "bl_recompute_range": (ctypes.c_int, [_H, _SYS, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_void]),
    "bl_apply_edits": (ctypes.c_int, [_H, _SYS, ctypes.c_double, c_void, c_void, c_void]),
    "bl_host_pow2_table": (ctypes.c_int, [c_double_p, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int, c_double_p]),
    "bl_rebuild_cover": (ctypes.c_int, [_H, _SYS, c_void]),
    "bl_symmetrize": (ctypes.c_int, [_H, _SYS, c_void]),
    "bl_row_minima": (ctypes.c_int, [_H, _SYS, c_void, c_void, c_void]),
    "bl_evolve": (ctypes.c_int, [_H, _SYS, ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int,
                                 ctypes.POINTER(BlLog), ctypes.POINTER(BlResult), c_void]),
    "bl_evolve_snapshot": (ctypes.c_int, [_H, _SYS, ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int,
                                          ctypes.POINTER(BlLog), ctypes.POINTER(BlResult), c_void, c_void, c_void]),
    "bl_next": (ctypes.c_int, [_H, _SYS, ctypes.c_double, ctypes.POINTER(BlResult), c_void]),
    "bl_snapshot": (ctypes.c_int, [_H, _SYS, ctypes.c_double, c_void, c_void, c_void]),
    "bl_positions": (ctypes.c_int, [_H, _SYS, ctypes.c_double, c_void, c_void]),
    "bl_toi_ball_ball": (ctypes.c_int, [_H, c_void, c_void, ctypes.c_double, c_void, ctypes.c_int64, c_void]),
    "bl_elastic_collision": (ctypes.c_int, [_H, c_void, c_void, c_void, ctypes.c_int64, c_void]),
    "bl_obstacle_detect": (ctypes.c_int, [_H, ctypes.POINTER(BlObstacle), c_void, c_void, c_void, c_void,
                                          ctypes.c_int64, c_void]),
    "bl_obstacle_resolve": (ctypes.c_int, [_H, ctypes.POINTER(BlObstacle), c_void, c_void, c_void, c_void,
                                           ctypes.c_int64, c_void]),
    "bl_toi_ball_point": (ctypes.c_int, [_H, c_void, c_void, ctypes.c_double, c_void, ctypes.c_int64, c_void]),
    "bl_toi_and_param_ball_segment": (ctypes.c_int, [_H, ctypes.POINTER(BlObstacle), c_void, ctypes.c_double, c_void,
                                                     c_void, c_void, ctypes.c_int64, c_void]),
    "bl_ensemble_run": (ctypes.c_int, [_H, ctypes.POINTER(BlObstacle), ctypes.c_int, c_double_p, ctypes.c_double,
                                       ctypes.c_int64, c_void, c_void, ctypes.c_int64, ctypes.c_double, c_void, c_void,
                                       c_void, c_void]),
    "bl_ensemble_stats": (ctypes.c_int, [_H, c_void, c_void, ctypes.c_int, ctypes.c_int64, c_void, c_void]),
    "bl_selftest_div2": (ctypes.c_int, [_H, ctypes.c_uint64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), c_double_p,
                                        c_void]),
    "bl_multi_run": (ctypes.c_int, [_H, ctypes.POINTER(BlMulti), ctypes.c_int, ctypes.c_double, ctypes.c_int64, c_void]),
    "bl_multi_positions": (ctypes.c_int, [_H, ctypes.POINTER(BlMulti), ctypes.c_double, ctypes.c_int, c_void, c_void]),
    "bl_measure_fp64_peak": (ctypes.c_int, [_H, c_double_p, c_void]),
    "bl_last_kernel_ms": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_float)]),
    "bl_launch_count": (ctypes.c_int64, [_H]),
    "bl_evolve_config": (ctypes.c_int, [_H, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "bl_set_ensemble_generic": (ctypes.c_int, [_H, ctypes.c_int]),
    "bl_set_batch_target": (ctypes.c_int, [_H, ctypes.c_int]),
    "bl_set_small_kernel": (ctypes.c_int, [_H, ctypes.c_int]),
    "bl_set_cta_shape_limit": (ctypes.c_int, [_H, ctypes.c_int]),
}

_lib = None
_lock = threading.Lock()


def load():
    """dlopen the CUDA library; raises (never falls back) when it is not there."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `make -C python-billiards_b200/csrc` "
                    "(or __graft_entry__.build()); billiards_b200 has no CPU fallback")
            L = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


class BilliardsCudaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code


# ---- numpy dot mode --------------------------------------------------------------------------------------

_dot_mode = None


def probe_dot_fma():
    """How does THIS host's numpy compute a 2-vector dot?  

    The reference's arithmetic is ``ndarray.dot`` on 2-vectors, which OpenBLAS evaluates as
    ``fma(a1, b1, fl(a0*b0))`` on FMA-capable CPUs and as ``fl(fl(a0*b0) + fl(a1*b1))`` otherwise.
    The kernels implement both; this probe selects the one that reproduces the local numpy bit for bit,
    so results equal what the reference would give on the same machine.
    Override with BILLIARDS_B200_DOT=fma|plain.
    """
    global _dot_mode
    if _dot_mode is not None:
        return _dot_mode
    env = os.environ.get("BILLIARDS_B200_DOT", "").lower()
    if env in ("fma", "plain"):
        _dot_mode = env == "fma"
        return _dot_mode
    try:
        import math
        fma = getattr(math, "fma", None)
        if fma is None:
            libm = ctypes.CDLL("libm.so.6")
            libm.fma.restype = ctypes.c_double
            libm.fma.argtypes = [ctypes.c_double] * 3
            fma = libm.fma
        rng = np.random.default_rng(20240229)
        a, b = rng.standard_normal((2048, 2)), rng.standard_normal((2048, 2))
        fused = plain = True
        for x, y in zip(a, b):
            d = float(x.dot(y))
            x0, x1, y0, y1 = float(x[0]), float(x[1]), float(y[0]), float(y[1])
            fused = fused and d == fma(x1, y1, x0 * y0)
            plain = plain and d == x0 * y0 + x1 * y1
        if fused == plain:
            warnings.warn("numpy's 2-vector dot matches neither known formula on this host; "
                          "assuming the fused variant", RuntimeWarning)
            _dot_mode = True
        else:
            _dot_mode = fused
    except OSError:
        _dot_mode = True
    return _dot_mode


# ---- device handles ----------------------------------------------------------------------------------------

_handles = {}


class Handle:
    """One bl_handle per (device, dot mode)."""

    def __init__(self, device, dot_fma):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("billiards_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load()
        self.device = int(device)
        self.dot_fma = bool(dot_fma)
        self.torch_device = torch.device("cuda", self.device)
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.torch_device)  # make sure torch's context is the primary context
            h = c_void()
            rc = self.lib.bl_create(ctypes.byref(h), self.device, int(self.dot_fma))
            self.h = h
            if rc != BL_OK:
                msg = self.lib.bl_last_error(h).decode() if h else "bl_create failed"
                if h:
                    self.lib.bl_destroy(h)
                raise BilliardsCudaError(rc, msg)

    def check(self, rc):
        if rc != BL_OK:
            raise BilliardsCudaError(rc, self.lib.bl_last_error(self.h).decode())

    def stream(self):
        import torch
        return c_void(torch.cuda.current_stream(self.torch_device).cuda_stream)

    def launch_count(self):
        return int(self.lib.bl_launch_count(self.h))

    def fp64_peak_tflops(self):
        x = ctypes.c_double()
        self.check(self.lib.bl_measure_fp64_peak(self.h, ctypes.byref(x), self.stream()))
        return float(x.value)

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        self.check(self.lib.bl_last_kernel_ms(self.h, ctypes.byref(ms)))
        return float(ms.value)


def get_handle(device=None, dot_fma=None):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    if dot_fma is None:
        dot_fma = probe_dot_fma()
    key = (int(device), bool(dot_fma))
    if key not in _handles:
        _handles[key] = Handle(*key)
    return _handles[key]


def to_host(t):
    """Device tensor -> numpy array.  Read-outs of 1 MB .. 256 MB land in page-locked memory from torch's caching
    host allocator: the copy runs at the full PCIe rate instead of being staged through the driver's bounce buffer,
    and the block goes back to the cache when the array is dropped (16 MB of an ensemble: ~4 ms -> ~0.4 ms)."""
    import torch
    nbytes = t.numel() * t.element_size()
    if t.is_cuda and (1 << 20) <= nbytes <= (1 << 28):
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t)
        return out.numpy()
    return t.cpu().numpy()


def ptr(t):
    """device address of a torch tensor (or None)"""
    return c_void(0 if t is None else t.data_ptr())


def pow2_table(radii, disk_radii=()):
    """(r_a + r_b)**2 over distinct ball radii and (R_q + r_a)**2 for disks, with the libm pow() CPython's
    ``float ** 2`` uses (physics.py:52)."""
    L = load()
    r = np.ascontiguousarray(radii, dtype=np.float64)
    R = np.ascontiguousarray(disk_radii, dtype=np.float64)
    K, nd = r.size, R.size
    out = np.empty((K, K), dtype=np.float64)
    out_disk = np.empty((max(nd, 1), max(K, 1)), dtype=np.float64)
    rc = L.bl_host_pow2_table(r.ctypes.data_as(c_double_p), K, out.ctypes.data_as(c_double_p),
                              R.ctypes.data_as(c_double_p) if nd else None, nd,
                              out_disk.ctypes.data_as(c_double_p) if nd else None)
    if rc != BL_OK:
        raise BilliardsCudaError(rc, "bl_host_pow2_table")
    return out, out_disk[:nd, :K]
