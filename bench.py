#!/usr/bin/env python3
"""bench.py -- resolved collisions per second of the Billiard.evolve hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's algorithm on the host CPU

Workloads (BASELINE.json):
  `dense_gas_16384` (config D): ONE system of 16384 hard disks at packing fraction 0.5; a step is `--events-per-step`
         collisions of that system (state and the 2.1 GB time-of-impact table stay in HBM).  The headline of a plain
         `python bench.py` (N = 1).  The same line carries `configs` (Galperin, pool, ideal gas N = 1024: configs G, P, I),
         `frame_loop`, `ensemble` (config S on one GPU), `pool_ensemble` and `gas64_ensemble` (multi-ball ensembles).
  `sinai_ensemble` (config S): 2^20 independent Sinai billiards per GPU (`--scaling weak`, default) or in total
         (`--scaling strong`), contiguous shards, no exchange during the run, one NCCL all-reduce of the statistics
         per step; a step is 10^4 collisions of every system.  The headline whenever the process was started by
         torch.distributed.run (WORLD_SIZE in the environment) -- at EVERY N including 1, so that a 1/2/4/8 series is
         one workload -- or with `--workload ensemble`.  The single system is not partitioned (its per-event global
         argmin does not shard): "replicas only".
  `pool_ensemble`: 2^16 pool tables (16 balls each) per GPU, every table evolved to t = 100 (`--workload pool_ensemble`).

`value` is measured with the inputs resident in HBM (CUDA events around K steps, max over ranks); `e2e` is the same
metric through the public Python API starting from HOST arrays every step (H2D of the initial conditions, table
build, the events, D2H of the resulting state).  `roofline` is for the dominant kernel (bl_evolve_kernel: algorithmic
bytes 76*N per ball-ball and 68*N per ball-obstacle event over the kernel time, against the measured HBM copy
bandwidth; bl_ensemble_kernel: algorithmic FP64 operations against the FP64 FMA peak measured in this run).
`cpu_baseline` times the CPU oracle -- a C restatement of the reference's algorithm, bit-identical to it -- on a bounded
prefix of the same workload on this box's host cores.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import workloads  # noqa: E402

INF = float("inf")
FP64_FLOPS_PER_SINAI_COLLISION = 60.0   # algorithmic adds/muls/divs/sqrts per collision (fma = 2)
# the reference itself (pure Python + numpy, 1 core) on the survey box, SURVEY.md section 6 / BASELINE.md
PYTHON_REFERENCE_COLLISIONS_PER_S = {"galperin": 46.6e3, "pool": 4.06e3, "ideal_gas_1024": 175.0, "dense_gas_16384": 5.6,
                                     "sinai_single_particle": 27.8e3}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def load_traffic():
    """DRAM bytes of the two dominant kernels from this round's `ncu --set full` captures (profiles/r2_dram_traffic.json,
    written by tools/ncu_traffic.py from the .ncu-rep files; per collision / per system).  None when absent."""
    path = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        for _ in range(8):  # a timed region shorter than the sampling period: wait for the first sample
            if self.rows:
                break
            time.sleep(0.1)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if not sm:
            return None
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[-1][2]), "samples": len(sm),
                "reasons": sorted(reasons)}


def dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def max_over_ranks(x, device):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def sum_over_ranks(x, device):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    return x


def barrier(device):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    torch.cuda.synchronize(device)


# ----------------------------------------------------------------------------------------------------
# workload D: one dense gas
# ----------------------------------------------------------------------------------------------------
def gas_invariants(bld, cfg, device):
    """What the physics conserves, checked on the state a long run leaves: sum |v|^2 (equal masses), every ball
    inside the box, no two balls overlapping (all pairs, on the GPU)."""
    import torch
    pos = torch.from_numpy(bld.balls_position).to(device)
    vel = bld.balls_velocity
    n, r = cfg.n, float(cfg.radius[0])
    side = float(cfg.obstacles[1][1][0])
    energy = float((vel * vel).sum()) / float((cfg.vel * cfg.vel).sum())
    lo, hi = float(pos.min().item()), float(pos.max().item())
    min_gap = INF
    for a in range(0, n, 2048):
        d = torch.cdist(pos[a:a + 2048], pos, compute_mode="donot_use_mm_for_euclid_dist")
        d[torch.arange(d.shape[0], device=d.device), torch.arange(a, a + d.shape[0], device=d.device)] = INF
        min_gap = min(min_gap, float(d.min().item()))
    out = {"kinetic_energy_ratio": energy, "min_coordinate": lo, "max_coordinate": hi, "box": [0.0, side], "radius": r,
           "min_centre_distance": min_gap, "min_allowed": 2 * r}
    assert abs(energy - 1.0) < 1e-9, out
    assert lo >= r - 1e-9 and hi <= side - r + 1e-9, out
    assert min_gap >= 2 * r - 1e-9, out
    return out


def bench_gas(args, device, fma, steps, warmup, e2e_steps=5):
    import torch
    import billiards
    from billiards import _lib
    side = args.gas_side
    cfg = workloads.dense_gas(side)
    N = cfg.n
    h = _lib.get_handle(device, fma)
    obstacles = workloads.make_obstacles(billiards, cfg.obstacles)
    E = args.events_per_step

    t0 = time.perf_counter()
    bld = billiards.Billiard(obstacles, device=device, dot_fma=fma)
    bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    bld._flush()
    torch.cuda.synchronize(device)
    build_s = time.perf_counter() - t0

    for _ in range(warmup):
        bld.evolve_events(E)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(device)
    sampler.start()
    launches0 = h.launch_count()
    n_bb = n_bo = 0
    kernel_ms = 0.0
    rescans = 0
    prof = [0] * 16
    barrier(device)
    ev0.record()
    for _ in range(steps):
        c = bld.evolve_events(E)
        n_bb += c[0]
        n_bo += c[1]
        kernel_ms += h.last_kernel_ms()
        rescans += int(bld.last_result.n_rescans)
        prof = [a + int(b) for a, b in zip(prof, bld.last_result.prof)]
    ev1.record()
    barrier(device)
    clocks = sampler.stop()
    launches = h.launch_count() - launches0
    total_ms = ev0.elapsed_time(ev1)
    value = (n_bb + n_bo) / (total_ms * 1e-3)
    invariants = gas_invariants(bld, cfg, device)

    # roofline of the dominant kernel: algorithmic bytes / kernel time (CUDA events inside the library, same stream)
    peak, peak_src, _ = load_peaks()
    traffic = load_traffic().get(f"evolve_dram_bytes_per_collision_{N}")
    alg_bytes = n_bb * 76.0 * N + n_bo * 68.0 * N
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"kernel": "bl_evolve_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic * (n_bb + n_bo) / max(steps, 1) if traffic else None,
                "traffic_source": load_traffic().get("evolve_source") if traffic else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(steps, 1),
                "kernel_ms_per_launch": kernel_ms / max(steps, 1),
                "note": "one launch commits batches of causally independent collisions per grid-wide round; the "
                        "rounds are bound by synchronisation + FP64 latency, not by HBM (DESIGN.md section 5)"}

    # e2e: host arrays -> Billiard -> E collisions -> host arrays, every step
    h2d = cfg.pos.nbytes + cfg.vel.nbytes + cfg.radius.nbytes + cfg.mass.nbytes
    d2h = 0
    t_e2e = 0.0
    done = 0
    for it in range(1 + e2e_steps):  # first one is a warm-up
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        b2 = billiards.Billiard(obstacles, device=device, dot_fma=fma)
        b2.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
        c = b2.evolve_events(E)
        pos, vel = b2.balls_position, b2.balls_velocity
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        if it:
            t_e2e += dt
            done += c[0] + c[1]
            d2h = 7 * 8 * N + 128
        del b2
    e2e = {"value": done / t_e2e, "unit": "collisions/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "includes": "H2D of initial conditions, 134M-pair table build, the events, D2H of positions/velocities",
           "steps": e2e_steps}
    # the frame-loop flavour: a live Billiard, every step = evolve + D2H of the state (what animate() does)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    done = 0
    for _ in range(e2e_steps):
        c = bld.evolve_events(E)
        pos, vel = bld.balls_position, bld.balls_velocity
        done += c[0] + c[1]
    torch.cuda.synchronize(device)
    e2e["live_object_value"] = done / (time.perf_counter() - t0)

    return {
        "value": value, "ms_per_step": total_ms / steps, "roofline": roofline, "e2e": e2e, "clocks": clocks,
        "gpu_launches": launches, "steps": steps,
        "config": {"workload": f"dense_gas_{N}", "balls": N, "packing_fraction": 0.5, "obstacles": "4 InfiniteWall",
                   "collisions_per_step": E, "table_bytes": int(N) * ((N + 15) // 16 * 16) * 8,
                   "l2": "inputs larger than L2 (table 2.1 GB, one event touches 4 rows/columns of it)",
                   "ctas_x_threads": list(_evolve_config(h, N)), "dot_fma": bool(fma)},
        "counts": {"ball_ball": n_bb, "ball_obstacle": n_bo, "stale_row_rescans": rescans},
        "state_invariants_after_run": invariants,
        "phase_cycles_per_round": {k: prof[q] / max(prof[5], 1) for k, q in
                                   (("post", 8), ("grid_barrier", 9), ("collect", 11), ("resolve_and_rank", 2),
                                    ("batch_cut_beside_row_stores", 0), ("solve", 10), ("optimistic_commit", 6), ("repair", 3))},
        "repair_phases": prof[4], "rounds": prof[5], "reposted_or_redone_rounds": prof[7],
        "collisions_per_round": (n_bb + n_bo) / max(prof[5], 1),
        "build_seconds": build_s,
    }, cfg


def _evolve_config(h, n):
    import ctypes
    c, t = ctypes.c_int(), ctypes.c_int()
    h.lib.bl_evolve_config(h.h, n, ctypes.byref(c), ctypes.byref(t))
    return c.value, t.value


# ----------------------------------------------------------------------------------------------------
# configs G, P, I and the frame loop (N = 1 line only)
# ----------------------------------------------------------------------------------------------------
def bench_single_config(name, cfg, device, fma, end_time=None, n_events=None, reps=3, expect=None):
    """One of the reference's own examples through the public API: device value = collisions / kernel time of the one
    launch that handles the whole run; e2e = from host arrays (Billiard(), add_ball one by one as a user would, evolve,
    read balls_position)."""
    import torch
    import billiards
    from billiards import _lib
    h = _lib.get_handle(device, fma)
    obstacles = workloads.make_obstacles(billiards, cfg.obstacles)
    best_ms, best_wall, counts = None, None, None
    for _ in range(reps):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        bld = billiards.Billiard(obstacles, device=device, dot_fma=fma)
        if cfg.n <= 64:
            for k in range(cfg.n):
                bld.add_ball(tuple(cfg.pos[k].tolist()), tuple(cfg.vel[k].tolist()), float(cfg.radius[k]), float(cfg.mass[k]))
        else:
            bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
        c = bld.evolve(end_time) if n_events is None else bld.evolve_events(n_events)
        pos = bld.balls_position
        torch.cuda.synchronize(device)
        wall = time.perf_counter() - t0
        ms = h.last_kernel_ms()
        if best_ms is None or ms < best_ms:
            best_ms = ms
        if best_wall is None or wall < best_wall:
            best_wall = wall
        counts = c
    if expect is not None:
        assert tuple(counts) == tuple(expect), f"{name}: {counts} != {expect}"
    total = counts[0] + counts[1]
    return {"value": total / (best_ms * 1e-3), "unit": "collisions/s", "kernel_ms": best_ms, "counts": list(counts),
            "e2e": {"value": total / best_wall, "unit": "collisions/s", "wall_ms": best_wall * 1e3,
                    "h2d_bytes_per_step": int(6 * 8 * cfg.n), "d2h_bytes_per_step": int(7 * 8 * cfg.n + 128),
                    "includes": "Billiard(), add_ball x N, table build, the one evolve launch, D2H of the positions"},
            "python_reference_collisions_per_s_survey_box": PYTHON_REFERENCE_COLLISIONS_PER_S.get(name)}


def cpu_single_config(cfg, fma, end_time=None, n_events=None, seconds=3.0):
    from oracle import oracle
    t0 = time.perf_counter()
    o = cfg.build_oracle(oracle, dot_fma=fma)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    if n_events is None:
        c = o.evolve(end_time)
        done = sum(c)
    else:
        done, chunk = 0, 200
        while time.perf_counter() - t0 < seconds and done < n_events:
            done += sum(o.evolve(INF, max_events=chunk))
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "collisions/s", "cores": 1, "kind": "port",
            "sample": f"{done} collisions in {dt:.3f} s (table build {build_s:.3f} s not counted)",
            "host_cpus": os.cpu_count()}


def bench_configs(args, device, fma):
    out = {}
    g = workloads.galperin()
    out["galperin"] = bench_single_config("galperin", g, device, fma, end_time=16.0, expect=(157080, 157079))
    p = workloads.pool()
    out["pool"] = bench_single_config("pool", p, device, fma, end_time=100.0, expect=(167, 248), reps=5)
    i = workloads.ideal_gas()
    out["ideal_gas_1024"] = bench_single_config("ideal_gas_1024", i, device, fma, n_events=10**6, reps=2)
    if not args.no_cpu_baseline:
        out["galperin"]["cpu_baseline"] = cpu_single_config(g, fma, end_time=16.0)
        out["pool"]["cpu_baseline"] = cpu_single_config(p, fma, end_time=100.0)
        out["ideal_gas_1024"]["cpu_baseline"] = cpu_single_config(i, fma, n_events=10**6, seconds=3.0)
    return out


def bench_frame_loop(device, fma, frames=600):
    """What visualize_matplotlib.animate (:711-723) and the pyglet window (:737-766) of the reference do: many short
    evolve calls, each followed by a read of balls_position.  N = 1024 ideal gas, dt = 1/60."""
    import torch
    import billiards
    cfg = workloads.ideal_gas()
    bld = billiards.Billiard(workloads.make_obstacles(billiards, cfg.obstacles), device=device, dot_fma=fma)
    bld.add_balls(cfg.pos, cfg.vel, cfg.radius, cfg.mass)
    bld.evolve(0.05)
    _ = bld.balls_position
    torch.cuda.synchronize(device)
    dt = 1.0 / 60.0
    t0 = time.perf_counter()
    n = 0
    for k in range(frames):
        c = bld.evolve(0.05 + (k + 1) * dt)
        pos = bld.balls_position  # noqa: F841
        n += c[0] + c[1]
    wall = time.perf_counter() - t0
    return {"workload": "ideal_gas_1024, 600 x (evolve(t + 1/60) + balls_position)", "frames": frames, "collisions": n,
            "calls_per_s": frames / wall, "ms_per_call": wall / frames * 1e3}


# ----------------------------------------------------------------------------------------------------
# workload S: Sinai ensemble, one shard per rank
# ----------------------------------------------------------------------------------------------------
def bench_ensemble(args, device, fma, rank, world, steps, warmup):
    import torch
    import billiards
    from billiards import _lib
    from billiards.ensemble import run_from_host, shard_bounds
    n_coll = args.collisions_per_system
    strong = args.scaling == "strong"
    n_total = args.systems_per_gpu if strong else args.systems_per_gpu * world
    lo, hi = shard_bounds(n_total, world, rank)
    n_local = hi - lo
    h = _lib.get_handle(device, fma)
    # initial conditions are generated per block of 2^17 systems (seed = 1000 + block), so that a rank's contiguous
    # shard is the same bits whatever the partition: results are partition-invariant
    blk = 1 << 17
    pos = np.empty((n_local, 2))
    vel = np.empty((n_local, 2))
    for b in range(lo // blk, (hi + blk - 1) // blk):
        ob, p, v = workloads.sinai(blk, seed=1000 + b)
        s0, s1 = max(lo, b * blk), min(hi, (b + 1) * blk)
        pos[s0 - lo:s1 - lo] = p[s0 - b * blk:s1 - b * blk]
        vel[s0 - lo:s1 - lo] = v[s0 - b * blk:s1 - b * blk]
    obstacles = workloads.make_obstacles(billiards, workloads.SINAI_OBSTACLES)
    fp64_peak = h.fp64_peak_tflops()  # measured now, on this GPU (8 independent DFMA chains per thread)
    ens = billiards.BilliardEnsemble(obstacles, pos, vel, device=device, dot_fma=fma, count_hits=True)
    for _ in range(warmup):
        ens.run(n_coll)
        ens.gather_stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(device)
    sampler.start()
    launches0 = h.launch_count()
    before = int(ens.stats()[0].item())
    kernel_ms = 0.0
    barrier(device)
    ev0.record()
    for _ in range(steps):
        ens.run(n_coll, sync=False)
        stats = ens.gather_stats()      # NCCL all-reduce of [collisions, hits per obstacle] -- the only collective
        kernel_ms += h.last_kernel_ms()
    ev1.record()
    barrier(device)
    clocks = sampler.stop()
    launches = h.launch_count() - launches0 - 1
    total_ms = max_over_ranks(ev0.elapsed_time(ev1), device)
    local_done = int(ens.stats()[0].item()) - before
    done_all = sum_over_ranks(float(local_done), device)
    value = done_all / (total_ms * 1e-3)
    achieved = local_done * FP64_FLOPS_PER_SINAI_COLLISION / (kernel_ms * 1e-3) / 1e12
    traffic = load_traffic().get("ensemble_dram_bytes_per_system")
    roofline = {"kernel": "bl_ensemble_kernel_u", "bound": "fp64", "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": traffic * n_local if traffic else None,
                "traffic_source": load_traffic().get("ensemble_source") if traffic else None,
                "peak_source": "FP64 FMA peak measured in this run (bl_measure_fp64_peak: 8 independent DFMA chains "
                               "per thread, 2048 threads per SM, best of 3)",
                "algorithmic_flops_per_collision": FP64_FLOPS_PER_SINAI_COLLISION,
                "kernel_ms_per_launch": kernel_ms / max(steps, 1)}
    # e2e: pinned host arrays in, host results out, every step, through the public pipelined call
    e2e_steps = max(1, min(steps, 3))
    pos_pin = torch.from_numpy(pos).pin_memory()
    vel_pin = torch.from_numpy(vel).pin_memory()
    run_from_host(obstacles, pos_pin, vel_pin, n_coll, device=device, dot_fma=fma, chunks=args.e2e_chunks)  # warm-up
    barrier(device)
    t0 = time.perf_counter()
    done = 0
    for _ in range(e2e_steps):
        out = run_from_host(obstacles, pos_pin, vel_pin, n_coll, device=device, dot_fma=fma, chunks=args.e2e_chunks)
        done += int(out["stats"][0])
    barrier(device)
    dt = max_over_ranks(time.perf_counter() - t0, device)
    e2e = {"value": done / dt, "unit": "collisions/s", "h2d_bytes_per_step": int(n_total * 4 * 8),
           "d2h_bytes_per_step": int(n_total * 6 * 8 + 48 * world), "steps": e2e_steps, "chunks": args.e2e_chunks,
           "includes": "H2D of positions/velocities from pinned memory, the run, D2H of time/position/velocity/counts "
                       "of every system, the statistics all-reduce; copies of one block overlap the kernel of another"}
    return {"value": value, "ms_per_step": total_ms / steps, "roofline": roofline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches, "steps": steps,
            "config": {"workload": "sinai_ensemble", "systems_per_gpu": n_local, "systems_total": n_total,
                       "collisions_per_system_per_step": n_coll, "obstacles": "4 InfiniteWall + Disk(0.5)",
                       "sharding": "contiguous block per rank, no exchange; one NCCL all-reduce of statistics per step",
                       "l2": "no memory traffic inside the collision loop (state in registers)",
                       "dot_fma": bool(fma)},
            "stats": [int(x) for x in stats]}


# ----------------------------------------------------------------------------------------------------
# multi-ball ensemble: pool tables
# ----------------------------------------------------------------------------------------------------
def pool_tables(n_tables, seed):
    cfg = workloads.pool()
    rng = np.random.default_rng(seed)
    pos = np.ascontiguousarray(np.broadcast_to(cfg.pos[None], (n_tables,) + cfg.pos.shape))
    vel = np.ascontiguousarray(np.broadcast_to(cfg.vel[None], (n_tables,) + cfg.vel.shape))
    speed = cfg.vel[15, 0] * rng.uniform(0.8, 1.2, n_tables)
    ang = rng.uniform(-0.02, 0.02, n_tables)
    vel[:, 15, 0] = speed * np.cos(ang)
    vel[:, 15, 1] = speed * np.sin(ang)
    return cfg, pos, vel


def bench_gas_ensemble(device, fma, side=8, systems=8192, end_time=2.0, steps=3):
    """Ensemble of mid-size systems, one CTA per system (csrc/bl_evolve_cta.cu): `systems` dense gases of side*side balls
    with different velocity directions, every one built (table) and evolved to `end_time` in one launch."""
    import torch
    import billiards
    from billiards import _lib
    h = _lib.get_handle(device, fma)
    cfg = workloads.dense_gas(side)
    rng = np.random.default_rng(side)
    pos = np.ascontiguousarray(np.broadcast_to(cfg.pos[None], (systems,) + cfg.pos.shape))
    th = rng.uniform(0, 2 * np.pi, (systems, cfg.n))
    vel = np.stack([np.cos(th), np.sin(th)], axis=2)
    obstacles = workloads.make_obstacles(billiards, cfg.obstacles)
    pos_pin, vel_pin = torch.from_numpy(pos).pin_memory(), torch.from_numpy(vel).pin_memory()
    ms, total = 0.0, 0
    for step in range(steps + 1):
        ens = billiards.BilliardEnsemble(obstacles, pos_pin, vel_pin, radius=float(cfg.radius[0]), mass=1.0, device=device,
                                         dot_fma=fma)
        bb, bo = ens.evolve(end_time)
        if step:  # the first one warms up
            ms += h.last_kernel_ms()
            total += int(bb.sum() + bo.sum())
    return {"value": total / (ms * 1e-3), "unit": "collisions/s", "ms_per_step": ms / steps, "steps": steps,
            "config": {"workload": "gas_ensemble", "systems": systems, "balls_per_system": cfg.n, "end_time": end_time,
                       "kernel": "bl_evolve_cta_kernel, one CTA per system, table build included"}}


def bench_pool_ensemble(args, device, fma, rank, world, steps, warmup):
    import torch
    import billiards
    from billiards import _lib
    h = _lib.get_handle(device, fma)
    S = args.pool_tables_per_gpu
    cfg, pos, vel = pool_tables(S, 2000 + rank)
    obstacles = workloads.make_obstacles(billiards, cfg.obstacles)
    pos_pin, vel_pin = torch.from_numpy(pos).pin_memory(), torch.from_numpy(vel).pin_memory()

    def one(sync_host):
        ens = billiards.BilliardEnsemble(obstacles, pos_pin, vel_pin, radius=cfg.radius, mass=cfg.mass, device=device,
                                         dot_fma=fma)
        ens.run(-1, end_time=100.0, sync=False)
        if sync_host:
            return ens, ens.balls_position, ens.balls_velocity, int(ens.gather_stats()[0])
        return ens, None, None, None

    for _ in range(max(warmup, 1)):
        one(False)
    torch.cuda.synchronize(device)
    kernel_ms, local = 0.0, 0
    sampler = ClockSampler(device)
    sampler.start()
    launches0 = h.launch_count()
    barrier(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_ms = 0.0
    for _ in range(steps):
        # a step = every table from its break shot to t = 100; building the ensemble (H2D) is outside the device-timed part
        ens = billiards.BilliardEnsemble(obstacles, pos_pin, vel_pin, radius=cfg.radius, mass=cfg.mass, device=device,
                                         dot_fma=fma)
        torch.cuda.synchronize(device)
        ev0.record()
        ens.run(-1, end_time=100.0, sync=False)
        ev1.record()
        torch.cuda.synchronize(device)
        tot_ms += ev0.elapsed_time(ev1)
        kernel_ms += h.last_kernel_ms()
        local += int(ens.stats()[0].item())
    barrier(device)
    clocks = sampler.stop()
    launches = h.launch_count() - launches0
    total_ms = max_over_ranks(tot_ms, device)
    done_all = sum_over_ranks(float(local), device)
    from billiards.ensemble import run_from_host
    run_from_host(obstacles, pos_pin, vel_pin, -1, end_time=100.0, radius=cfg.radius, mass=cfg.mass, device=device,
                  dot_fma=fma, chunks=args.e2e_chunks)  # warm-up: pinned result buffers, streams
    barrier(device)
    t0 = time.perf_counter()
    done = 0
    e2e_steps = max(1, min(steps, 3))
    for _ in range(e2e_steps):
        out = run_from_host(obstacles, pos_pin, vel_pin, -1, end_time=100.0, radius=cfg.radius, mass=cfg.mass,
                            device=device, dot_fma=fma, chunks=args.e2e_chunks)
        done += int(out["stats"][0])
    barrier(device)
    dt = max_over_ranks(time.perf_counter() - t0, device)
    return {"value": done_all / (total_ms * 1e-3), "ms_per_step": total_ms / steps, "clocks": clocks,
            "gpu_launches": launches, "steps": steps, "unit": "collisions/s",
            "kernel_ms_per_launch": kernel_ms / steps,
            "e2e": {"value": done / dt, "unit": "collisions/s", "h2d_bytes_per_step": int(pos.nbytes + vel.nbytes) * world,
                    "d2h_bytes_per_step": int(pos.nbytes + vel.nbytes + 24) * world, "steps": e2e_steps,
                    "includes": "H2D of the break shots from pinned memory, table build + every collision to t = 100, D2H of "
                                "clocks / positions / velocities / counts, the statistics all-reduce; blocks of tables are "
                                "pipelined over streams (ensemble.run_from_host)", "chunks": args.e2e_chunks},
            "config": {"workload": "pool_ensemble", "tables_per_gpu": S, "tables_total": S * world, "balls_per_table": 16,
                       "end_time": 100.0, "kernel": "bl_evolve_small_kernel<W=16>: 2 tables per warp, 8 per CTA",
                       "collisions_per_table": local / steps / S, "dot_fma": bool(fma),
                       "l2": "tables live in shared memory during the launch (2.2 KB per table)"},
            "python_reference_collisions_per_s_survey_box": PYTHON_REFERENCE_COLLISIONS_PER_S["pool"]}


# ----------------------------------------------------------------------------------------------------
# CPU legs (the oracle is test infrastructure; this is one of the places allowed to run it)
# ----------------------------------------------------------------------------------------------------
def cpu_gas_baseline(side, seconds, fma):
    from oracle import oracle
    cfg = workloads.dense_gas(side)
    t0 = time.perf_counter()
    o = cfg.build_oracle(oracle, dot_fma=fma)
    build_s = time.perf_counter() - t0
    done, t_run = 0, 0.0
    chunk = 10
    while t_run < seconds:
        t0 = time.perf_counter()
        c = o.evolve(INF, max_events=chunk)
        t_run += time.perf_counter() - t0
        done += sum(c)
    return {"value": done / t_run, "unit": "collisions/s", "cores": 1, "kind": "port",
            "sample": f"first {done} collisions of dense_gas_{cfg.n} (reference algorithm incl. its O(N^2) row "
                      f"re-argmin per event), {t_run:.1f} s; table build {build_s:.1f} s not counted",
            "host_cpus": os.cpu_count(),
            "python_reference_collisions_per_s_survey_box": PYTHON_REFERENCE_COLLISIONS_PER_S.get(f"dense_gas_{cfg.n}")}


def cpu_sinai_baseline(seconds, fma):
    from oracle import oracle
    obstacles, pos, vel = workloads.sinai(64, seed=1000)
    obs = workloads.oracle_obstacles(obstacles)
    done, t_run, n_coll = 0, 0.0, 20000
    while t_run < seconds:
        t0 = time.perf_counter()
        out = oracle.ensemble_run(obs, pos, vel, n_coll, dot_fma=fma)
        t_run += time.perf_counter() - t0
        done += int(out["done"].sum())
    return {"value": done / t_run, "unit": "collisions/s", "cores": 1, "kind": "port",
            "sample": f"64 Sinai systems x {n_coll} collisions, repeated for {t_run:.1f} s", "host_cpus": os.cpu_count(),
            "python_reference_collisions_per_s_survey_box": PYTHON_REFERENCE_COLLISIONS_PER_S["sinai_single_particle"]}


def cpu_pool_ensemble_baseline(fma, tables=64):
    from oracle import oracle
    cfg, pos, vel = pool_tables(tables, 2000)
    t0 = time.perf_counter()
    done = 0
    for s in range(tables):
        o = oracle.OracleBilliard(workloads.oracle_obstacles(cfg.obstacles), dot_fma=fma)
        o.add_balls(pos[s], vel[s], cfg.radius, cfg.mass)
        done += sum(o.evolve(100.0))
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "collisions/s", "cores": 1, "kind": "port",
            "sample": f"{tables} pool tables to t = 100, {done} collisions in {dt:.3f} s", "host_cpus": os.cpu_count()}


def pick_workload(args):
    """gas for a plain run, the Sinai ensemble whenever torch.distributed.run started the process (at every N, so that
    a scaling series is one workload)"""
    if args.workload != "auto":
        return args.workload
    return "ensemble" if "WORLD_SIZE" in os.environ else "gas"


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (C port of it, bit-identical; see oracle/) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    fma = oracle.probe_dot_fma()
    workload = pick_workload(args)
    if workload == "gas":
        cfg = workloads.dense_gas(args.gas_side)
        o = cfg.build_oracle(oracle, dot_fma=fma)
        per_step = args.reference_events_per_step
        for _ in range(args.warmup):
            o.evolve(INF, max_events=per_step)
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            done += sum(o.evolve(INF, max_events=per_step))
        dt = time.perf_counter() - t0
        config = {"workload": f"dense_gas_{cfg.n}", "balls": cfg.n, "packing_fraction": 0.5,
                  "collisions_per_step": per_step}
        sample = f"{done} collisions of dense_gas_{cfg.n} after {args.warmup * per_step} warm-up collisions"
    elif workload == "pool_ensemble":
        tables = 16
        for _ in range(args.warmup):
            cpu_pool_ensemble_baseline(fma, tables)
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            b = cpu_pool_ensemble_baseline(fma, tables)
            done += int(b["sample"].split(", ")[1].split(" ")[0])
        dt = time.perf_counter() - t0
        config = {"workload": "pool_ensemble", "tables": tables, "balls_per_table": 16, "end_time": 100.0}
        sample = f"{tables} pool tables to t = 100 per step (bounded sample of the 2^16-per-GPU job)"
    else:
        obstacles, pos, vel = workloads.sinai(64, seed=1000)
        obs = workloads.oracle_obstacles(obstacles)
        n_coll = 5000
        for _ in range(args.warmup):
            oracle.ensemble_run(obs, pos, vel, n_coll, dot_fma=fma)
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            done += int(oracle.ensemble_run(obs, pos, vel, n_coll, dot_fma=fma)["done"].sum())
        dt = time.perf_counter() - t0
        config = {"workload": "sinai_ensemble", "systems": 64, "collisions_per_system_per_step": n_coll}
        sample = f"64 Sinai systems x {n_coll} collisions per step (bounded sample of the 2^20-systems job)"
    value = done / dt
    line = {"impl": "reference", "metric": "resolved collisions/sec", "value": value, "unit": "collisions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "collisions/s", "cores": 1, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(), "dot_fma": bool(fma),
                             "note": "python-billiards is single-threaded pure Python; the port follows its algorithm "
                                     "step for step (bit-identical event logs) and is >= as fast as the original"},
            "e2e": {"value": value, "unit": "collisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "gas", "ensemble", "pool_ensemble"],
                    help="auto: dense gas for a plain run, Sinai ensemble under torch.distributed.run")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="ensemble: --systems-per-gpu systems per GPU (weak) or in total (strong)")
    ap.add_argument("--gas-side", type=int, default=128, help="dense gas has side**2 balls (128 -> 16384)")
    ap.add_argument("--events-per-step", type=int, default=1000000,
                    help="collisions per step; the default 10 steps are the 10^7 collisions of config D")
    ap.add_argument("--reference-events-per-step", type=int, default=10)
    ap.add_argument("--systems-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--collisions-per-system", type=int, default=10000)
    ap.add_argument("--pool-tables-per-gpu", type=int, default=1 << 16)
    ap.add_argument("--e2e-chunks", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N = 1 gas line without configs / frame loop / ensembles")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from billiards import _lib
    rank, world, local = dist_setup()
    device = local if world > 1 else torch.cuda.current_device()
    fma = _lib.probe_dot_fma()
    workload = pick_workload(args)

    line = {"metric": "resolved collisions/sec", "unit": "collisions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic"}
    keys = ("value", "ms_per_step", "config", "roofline", "e2e", "clocks", "gpu_launches")
    if workload == "gas":
        if world > 1:
            raise SystemExit("the single dense gas is not partitioned across GPUs (replicas only): run it with --gpus 1")
        gas, _ = bench_gas(args, device, fma, args.steps, args.warmup)
        line.update({k: gas[k] for k in keys})
        for key in ("counts", "state_invariants_after_run", "phase_cycles_per_round", "repair_phases", "rounds",
                    "reposted_or_redone_rounds", "collisions_per_round"):
            line[key] = gas[key]
        line["table_build_seconds"] = gas["build_seconds"]
        line["note"] = ("single-system workload: not partitioned across GPUs (replicas only).  The N = 1 point of the 1/2/4/8 "
                        "scaling series is `ensemble.value` of this line (or this script under torch.distributed.run with "
                        "--gpus 1): multi-GPU lines report the sharded Sinai ensemble.")
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_gas_baseline(args.gas_side, args.cpu_seconds, fma)
        if not args.no_extras:
            line["configs"] = bench_configs(args, device, fma)
            line["frame_loop"] = bench_frame_loop(device, fma)
            ens_steps = max(1, min(args.steps, 3))
            ens = bench_ensemble(args, device, fma, rank, world, ens_steps, 1)
            line["ensemble"] = {k: ens[k] for k in ("value", "ms_per_step", "config", "roofline", "e2e", "gpu_launches", "steps")}
            line["ensemble"]["unit"] = "collisions/s"
            pe = bench_pool_ensemble(args, device, fma, rank, world, 3, 1)
            line["pool_ensemble"] = pe
            line["gas64_ensemble"] = bench_gas_ensemble(device, fma)
            if not args.no_cpu_baseline:
                line["ensemble"]["cpu_baseline"] = cpu_sinai_baseline(min(args.cpu_seconds, 5.0), fma)
                line["pool_ensemble"]["cpu_baseline"] = cpu_pool_ensemble_baseline(fma)
    elif workload == "ensemble":
        ens = bench_ensemble(args, device, fma, rank, world, args.steps, args.warmup)
        line.update({k: ens[k] for k in keys})
        line["ensemble"] = {"value": ens["value"], "unit": "collisions/s", "stats": ens["stats"]}
        line["note"] = ("the scaling series is the sharded Sinai ensemble at every N; the single-system workload "
                        "(dense_gas_16384) is not partitioned across GPUs: replicas only")
        if world == 1 and rank == 0:
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_sinai_baseline(min(args.cpu_seconds, 10.0), fma)
            if not args.no_extras:
                gas, _ = bench_gas(args, device, fma, 3, 1, e2e_steps=1)
                line["gas"] = {k: gas[k] for k in ("value", "ms_per_step", "config", "roofline", "e2e", "gpu_launches", "steps")}
                line["gas"]["unit"] = "collisions/s"
    else:
        pe = bench_pool_ensemble(args, device, fma, rank, world, args.steps, args.warmup)
        line.update({k: pe[k] for k in ("value", "ms_per_step", "config", "e2e", "clocks", "gpu_launches")})
        line["roofline"] = {"kernel": "bl_evolve_small_kernel", "bound": "issue", "achieved": None, "peak": None,
                            "unit": "warp instructions/s", "frac": None, "traffic": None,
                            "note": "one group of 16 lanes per table, tables in shared memory: bound by instruction issue and "
                                    "FP64 latency (ncu summary in profiles/), not by HBM or the FP64 pipe's throughput"}
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_pool_ensemble_baseline(fma)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
