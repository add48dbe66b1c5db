"""CPU-only checks of the host side: the C ABI library loads and exports every symbol include/billiards_b200.h
declares, host helpers (pow table, dot probe, shard arithmetic, argument validation) behave, the product package
never touches oracle/, and the ensemble sharding + statistics gather work over a world_size-2 gloo group."""

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "python-billiards_b200")


def test_library_exports_every_declared_symbol():
    from billiards import _lib
    header = open(os.path.join(ROOT, "include", "billiards_b200.h")).read()
    declared = set(re.findall(r"\b(bl_[a-z0-9_]+)\s*\(", header))
    declared -= {"bl_handle"}
    L = _lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert b"sm_100a" in L.bl_version()


def test_struct_layouts_match_the_header():
    from billiards import _lib
    assert ctypes.sizeof(_lib.BlObstacle) == 72
    assert ctypes.sizeof(_lib.BlSystem) == 16 + 19 * 8
    assert ctypes.sizeof(_lib.BlResult) == 8 * 10 + 8 + 8 * 4 + 128 + 8
    assert ctypes.sizeof(_lib.BlLog) == 40


def test_pow_table_is_python_pow():
    from billiards import _lib
    rng = np.random.default_rng(0)
    r = rng.uniform(0, 5, 300)
    r[0], r[1] = 3.0, 8.237813583927716 - 3.0     # a sum whose pow(s, 2) differs from s*s on glibc
    R = np.asarray([0.5, 2.0])
    tab, tab_disk = _lib.pow2_table(r, R)
    rl, Rl = r.tolist(), R.tolist()
    want = np.asarray([[(a + b) ** 2 for b in rl] for a in rl])
    assert np.array_equal(tab.view(np.uint64), want.view(np.uint64))
    want_d = np.asarray([[(q + a) ** 2 for a in rl] for q in Rl])
    assert np.array_equal(tab_disk.view(np.uint64), want_d.view(np.uint64))


def test_dot_probe_matches_numpy():
    from billiards import _lib
    fma = _lib.probe_dot_fma()
    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(3)
    for _ in range(500):
        a, b = rng.standard_normal(2), rng.standard_normal(2)
        want = libm.fma(a[1], b[1], a[0] * b[0]) if fma else a[0] * b[0] + a[1] * b[1]
        assert float(a.dot(b)) == want


def test_wall_normal_matches_workloads_restatement():
    import billiards
    import workloads
    rng = np.random.default_rng(5)
    for _ in range(200):
        s, e = rng.uniform(-5, 5, 2), rng.uniform(-5, 5, 2)
        for side in ("left", "right"):
            w = billiards.InfiniteWall(s, e, side)
            assert np.array_equal(w._normal.view(np.uint64), workloads.wall_normal(s, e, side).view(np.uint64))
    w = billiards.InfiniteWall((0, -1), (0, 1), exterior="right")
    assert w._normal.tolist() == [1.0, -0.0]


def test_host_argument_validation_needs_no_gpu():
    import billiards
    with pytest.raises(TypeError):
        billiards.Billiard(obstacles=[42])
    with pytest.raises(ValueError):
        billiards.InfiniteWall((0, 0), (0, 0))
    with pytest.raises(ValueError):
        billiards.InfiniteWall((0, 0), (1, 0), "up")
    with pytest.raises(ValueError):
        billiards.LineSegment((1, 1), (1, 1))
    seg = billiards.LineSegment((0, 0), (3, 4))
    assert seg._covector.tolist() == [3 / 25, 4 / 25] and seg._normal.tolist() == [-4 / 5, 3 / 5]
    b = billiards.Billiard()
    with pytest.raises(ValueError):
        b.add_ball((0, 0, 0))
    with pytest.raises(TypeError):
        b.add_ball((0, 0), (0, 0), None)
    assert b.add_ball((1, 2), (3, 4), 0.5, 2.0) == 0 and b.count == 1
    assert b.balls_position.tolist() == [[1.0, 2.0]] and b.balls_radius == [0.5] and b.balls_mass == [2.0]


def test_queued_balls_mix_single_and_bulk_adds():
    """add_ball / add_balls only queue (host mirrors grow, nothing touches the device until the first query): indices,
    count and the mirrors follow the reference's add order; the queued chunks keep their own copies of the inputs."""
    import billiards
    b = billiards.Billiard()
    assert b.add_ball((0, 0), (1, 0), 0.5) == 0
    pos = np.array([[3.0, 0.0], [6.0, 1.0], [9.0, 2.0]])
    vel = np.array([[0.0, 1.0], [0.0, 2.0], [0.0, 3.0]])
    assert list(b.add_balls(pos, vel, radius=[0.1, 0.2, 0.1], mass=2.0)) == [1, 2, 3]
    pos[:] = -1.0  # the caller's arrays are not aliased
    assert b.add_ball((12, 3), (0, 4), 0.2, float("inf")) == 4
    assert list(b.add_balls(np.zeros((0, 2)), np.zeros((0, 2)))) == [] and b.count == 5
    assert b.balls_position.tolist() == [[0, 0], [3, 0], [6, 1], [9, 2], [12, 3]]
    assert b.balls_velocity.tolist() == [[1, 0], [0, 1], [0, 2], [0, 3], [0, 4]]
    assert b.balls_radius == [0.5, 0.1, 0.2, 0.1, 0.2] and b.balls_mass == [1.0, 2.0, 2.0, 2.0, float("inf")]
    assert b._n_pending == 5 and sum(c[0].shape[0] for c in b._pending) == 5
    queued = np.concatenate([c[0] for c in b._pending])
    assert queued.tolist() == [[0, 0], [3, 0], [6, 1], [9, 2], [12, 3]]


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import billiards
    b = billiards.Billiard()
    b.add_ball((0, 0), (1, 0), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b.evolve(1.0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no oracle", ""), f"{f} mentions the oracle"


def test_shard_bounds_partition():
    from billiards.ensemble import shard_bounds
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 3):
        for w in (1, 2, 4, 8):
            blocks = [shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "python-billiards_b200"))
import numpy as np, torch, torch.distributed as dist
import workloads
from billiards.ensemble import shard_bounds
from oracle import oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_total, n_coll = 96, 300
obstacles, pos, vel = workloads.sinai(n_total, seed=11)
lo, hi = shard_bounds(n_total, world, rank)
# host-side flow of the multi-GPU ensemble with the CPU oracle standing in for the kernel:
# contiguous shard -> independent run -> one all-reduce of the statistics
out = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos[lo:hi], vel[lo:hi], n_coll, count_hits=True)
stats = torch.tensor([int(out["done"].sum())] + out["hits"].sum(axis=0).tolist(), dtype=torch.int64)
dist.all_reduce(stats)
times = [None] * world
dist.all_gather_object(times, out["t"].tolist())
if rank == 0:
    full = oracle.ensemble_run(workloads.oracle_obstacles(obstacles), pos, vel, n_coll, count_hits=True)
    assert stats.tolist() == [int(full["done"].sum())] + full["hits"].sum(axis=0).tolist(), stats
    assert sum(times, []) == full["t"].tolist()          # partition-invariant, bit for bit
    print("gloo ok", stats.tolist())
dist.destroy_process_group()
'''


def test_ensemble_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "gloo ok" in res.stdout


def test_api_surface_matches_the_reference():
    """Drop-in boundary: every public class, method, property, function and parameter of the reference's
    hot-path modules exists here with the same name, kind and default -- recorded from the live reference by
    tests/golden/make_api_surface.py.  Extra trailing keyword parameters (device=, dot_fma=) are allowed."""
    import inspect
    import json
    import billiards
    with open(os.path.join(ROOT, "tests", "golden", "api_surface.json")) as f:
        want = json.load(f)

    def check(fn, params, where):
        have = [[p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
                for p in inspect.signature(fn).parameters.values()]
        assert have[:len(params)] == params, f"{where}: {have} vs reference {params}"
        for extra in have[len(params):]:
            assert extra[2] is not None or extra[1].startswith("VAR_"), f"{where}: extra parameter {extra} has no default"

    def check_class(cls, spec, where):
        for name, params in spec["methods"].items():
            check(getattr(cls, name), params, f"{where}.{name}")
        for name in spec["properties"]:
            assert isinstance(inspect.getattr_static(cls, name), property), f"{where}.{name} is not a property"

    for name in want["package"]:
        assert hasattr(billiards, name), name
    check_class(billiards.simulation.Billiard, want["simulation.Billiard"], "Billiard")
    for cls, spec in want["obstacles"].items():
        check_class(getattr(billiards.obstacles, cls), spec, cls)
        assert issubclass(getattr(billiards.obstacles, cls), billiards.obstacles.Obstacle)
    for name, params in want["physics"].items():
        check(getattr(billiards.physics, name), params, f"physics.{name}")
    b = billiards.Billiard()
    b.add_ball((1, 2), (3, 4), 0.5, 2.0)
    for attr in want["Billiard.attributes"]:
        if attr != "toi_table":            # needs the device
            assert hasattr(b, attr), attr
    assert isinstance(inspect.getattr_static(billiards.Billiard, "toi_table"), property) or hasattr(b, "toi_table")


def test_to_host_passes_cpu_tensors_through():
    import torch
    from billiards import _lib
    t = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    out = _lib.to_host(t)
    assert isinstance(out, np.ndarray) and out.tolist() == t.tolist()
