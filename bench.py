#!/usr/bin/env python3
"""bench.py -- resolved collisions per second of the Billiard.evolve hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's algorithm on the host CPU

Workloads (BASELINE.json):
  N = 1  `dense_gas_16384`: ONE system of 16384 hard disks at packing fraction 0.5 (config D); a step is
         `--events-per-step` collisions of that system (state and the 2.1 GB time-of-impact table stay in HBM).
         The same line also carries `ensemble` (config S on one GPU) so that the 1/2/4/8 scaling series has its
         N = 1 point.
  N > 1  `sinai_ensemble`: 2^20 independent Sinai billiards PER GPU (weak scaling, contiguous shards, no exchange
         during the run; one NCCL all-reduce of the statistics per step); a step is 10^4 collisions of every system.
         The single system is not partitioned (its per-event global argmin does not shard): "replicas only".

`value` is measured with the inputs resident in HBM (CUDA events around K steps, max over ranks); `e2e` is the same
metric through the public Python API starting from HOST arrays every step (H2D of the initial conditions, table
build, the events, D2H of the resulting state).  `roofline` is for the dominant kernel (bl_evolve_kernel: algorithmic
bytes 76*N per ball-ball and 68*N per ball-obstacle event over the kernel time, against the measured HBM copy
bandwidth; bl_ensemble_kernel: algorithmic FP64 operations against the FP64 pipe peak).  `cpu_baseline` times the
CPU oracle -- a C restatement of the reference's algorithm, bit-identical to it -- on a bounded prefix of the same
workload on this box's host cores.
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
SM_COUNT, FP64_LANES_PER_SM = 148, 64
# FP64 FMA peak measured on this pool's B200 with tools/microbench.cu (profiles/r1_microbench.log): 37.13 TFLOP/s,
# 64.0 FMA/clk/SM -- the datasheet formula at the observed clock; used when MEASURED_PEAKS.json has no FP64 entry
MEASURED_FP64_TFLOPS = 37.13
# DRAM traffic of the two dominant kernels from their `ncu --set full` captures (profiles/r1_final_evolve_kernel_ncu_summary.md:
# dram__bytes_read.sum + dram__bytes_write.sum = 1.841 GB + 5.043 GB for a 20 000-collision launch at N = 16384;
# profiles/r1_final_ensemble_ncu_summary.md: 58.8 MB + 8.0 MB for 2^20 systems) -- scaled to the launch timed here
NCU_EVOLVE_DRAM_BYTES_PER_COLLISION_16384 = (1.947861e9 + 5.039994e9) / 20000
NCU_ENSEMBLE_DRAM_BYTES_PER_SYSTEM = (58.755840e6 + 8.023552e6) / (1 << 20)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


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


def dist_setup(n_gpus):
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


def barrier(device):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    torch.cuda.synchronize(device)


# ----------------------------------------------------------------------------------------------------
# workload D: one dense gas
# ----------------------------------------------------------------------------------------------------
def bench_gas(args, device, fma):
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

    for _ in range(args.warmup):
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
    for _ in range(args.steps):
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

    # roofline of the dominant kernel: algorithmic bytes / kernel time (CUDA events inside the library, same stream)
    peak, peak_src, _ = load_peaks()
    alg_bytes = n_bb * 76.0 * N + n_bo * 68.0 * N
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"kernel": "bl_evolve_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": (NCU_EVOLVE_DRAM_BYTES_PER_COLLISION_16384 * (n_bb + n_bo) / max(args.steps, 1)
                            if N == 16384 else None),
                "traffic_source": "ncu --set full capture of a 20000-collision launch (profiles/), per collision, "
                                  "scaled to this launch; row stores 252 KB + 97 KB read (stale-candidate repairs) per collision",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(args.steps, 1),
                "kernel_ms_per_launch": kernel_ms / max(args.steps, 1),
                "note": "one launch commits batches of causally independent collisions per grid-wide round; the "
                        "rounds are bound by synchronisation + FP64 latency, not by HBM (DESIGN.md section 5)"}

    # e2e: host arrays -> Billiard -> E collisions -> host arrays, every step
    e2e_steps = max(1, min(args.steps, 5))
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
        "value": value, "ms_per_step": total_ms / args.steps, "roofline": roofline, "e2e": e2e, "clocks": clocks,
        "gpu_launches": launches,
        "config": {"workload": f"dense_gas_{N}", "balls": N, "packing_fraction": 0.5, "obstacles": "4 InfiniteWall",
                   "collisions_per_step": E, "table_bytes": int(N) * ((N + 15) // 16 * 16) * 8,
                   "l2": "inputs larger than L2 (table 2.1 GB, one event touches 4 rows/columns of it)",
                   "ctas_x_threads": list(_evolve_config(h, N)), "dot_fma": bool(fma)},
        "counts": {"ball_ball": n_bb, "ball_obstacle": n_bo, "stale_row_rescans": rescans},
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
# workload S: Sinai ensemble, one shard per rank
# ----------------------------------------------------------------------------------------------------
def bench_ensemble(args, device, fma, rank, world, steps, warmup):
    import torch
    import billiards
    from billiards import _lib
    n_local = args.systems_per_gpu
    n_coll = args.collisions_per_system
    h = _lib.get_handle(device, fma)
    # global system index = rank*n_local + i: generate the whole job's initial conditions deterministically and keep
    # this rank's contiguous block (partition-invariant results)
    obstacles_u, pos, vel = workloads.sinai(n_local, seed=1000 + rank)
    obstacles = workloads.make_obstacles(billiards, obstacles_u)
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
    done_all = local_done * world  # every rank runs the same number of collisions per system
    value = done_all / (total_ms * 1e-3)
    fp64_peak = MEASURED_FP64_TFLOPS   # TFLOP/s counting fma = 2
    achieved = local_done * FP64_FLOPS_PER_SINAI_COLLISION / (kernel_ms * 1e-3) / 1e12
    roofline = {"kernel": "bl_ensemble_kernel", "bound": "fp64", "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": NCU_ENSEMBLE_DRAM_BYTES_PER_SYSTEM * n_local,
                "traffic_source": "ncu --set full capture (profiles/r1_final_ensemble_ncu_summary.md): state load + store only",
                "peak_source": "measured DFMA peak, tools/microbench.cu (profiles/r1_microbench.log); equals "
                               "148 SM x 64 lanes x 2 x 1.96 GHz",
                "algorithmic_flops_per_collision": FP64_FLOPS_PER_SINAI_COLLISION,
                "kernel_ms_per_launch": kernel_ms / max(steps, 1)}
    # e2e: host arrays in, host results out, every step
    e2e_steps = max(1, min(steps, 3))
    pos_pin = torch.from_numpy(pos).pin_memory()
    vel_pin = torch.from_numpy(vel).pin_memory()
    barrier(device)
    t0 = time.perf_counter()
    done = 0
    for _ in range(e2e_steps):
        e2 = billiards.BilliardEnsemble(obstacles, pos_pin, vel_pin, device=device, dot_fma=fma)
        e2.run(n_coll)
        p0, v, t = e2.balls_initial_position, e2.balls_velocity, e2.time
        done += int(e2.gather_stats()[0])
    barrier(device)
    dt = max_over_ranks(time.perf_counter() - t0, device)
    e2e = {"value": done / dt, "unit": "collisions/s", "h2d_bytes_per_step": int(pos.nbytes + vel.nbytes) * world,
           "d2h_bytes_per_step": int(n_local * 5 * 8 + 48) * world, "steps": e2e_steps}
    return {"value": value, "ms_per_step": total_ms / steps, "roofline": roofline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches,
            "config": {"workload": "sinai_ensemble", "systems_per_gpu": n_local, "systems_total": n_local * world,
                       "collisions_per_system_per_step": n_coll, "obstacles": "4 InfiniteWall + Disk(0.5)",
                       "sharding": "contiguous block per rank, no exchange; one NCCL all-reduce of statistics per step",
                       "l2": "no memory traffic inside the collision loop (state in registers)",
                       "dot_fma": bool(fma)},
            "stats": [int(x) for x in stats]}


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
            "host_cpus": os.cpu_count()}


def cpu_sinai_baseline(seconds, fma):
    from oracle import oracle
    obstacles, pos, vel = workloads.sinai(64, seed=1000)
    obs = workloads.oracle_obstacles(obstacles)
    done, t_run, n_coll = 0, 0.0, 20000
    state_p, state_v = pos, vel
    while t_run < seconds:
        t0 = time.perf_counter()
        out = oracle.ensemble_run(obs, state_p, state_v, n_coll, dot_fma=fma)
        t_run += time.perf_counter() - t0
        done += int(out["done"].sum())
    return {"value": done / t_run, "unit": "collisions/s", "cores": 1, "kind": "port",
            "sample": f"64 Sinai systems x {n_coll} collisions, repeated for {t_run:.1f} s", "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (C port of it, bit-identical; see oracle/) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fma = True
    from oracle import oracle
    if args.gpus == 1:
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
        sample = f"64 Sinai systems x {n_coll} collisions per step (bounded sample of the 2^20-per-GPU job)"
    value = done / dt
    line = {"impl": "reference", "metric": "resolved collisions/sec", "value": value, "unit": "collisions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "collisions/s", "cores": 1, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(),
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
    ap.add_argument("--gas-side", type=int, default=128, help="dense gas has side**2 balls (128 -> 16384)")
    ap.add_argument("--events-per-step", type=int, default=1000000,
                    help="collisions per step; the default 10 steps are the 10^7 collisions of config D")
    ap.add_argument("--reference-events-per-step", type=int, default=10)
    ap.add_argument("--systems-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--collisions-per-system", type=int, default=10000)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from billiards import _lib
    rank, world, local = dist_setup(args.gpus)
    device = local if world > 1 else torch.cuda.current_device()
    fma = _lib.probe_dot_fma()

    line = {"metric": "resolved collisions/sec", "unit": "collisions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic"}
    if world == 1:
        gas, _ = bench_gas(args, device, fma)
        line.update({k: gas[k] for k in ("value", "ms_per_step", "config", "roofline", "e2e", "clocks", "gpu_launches")})
        line["counts"] = gas["counts"]
        for key in ("phase_cycles_per_round", "repair_phases", "rounds", "reposted_or_redone_rounds", "collisions_per_round"):
            line[key] = gas[key]
        line["table_build_seconds"] = gas["build_seconds"]
        ens_steps = max(1, min(args.steps, 3))
        ens = bench_ensemble(args, device, fma, rank, world, ens_steps, 1)
        line["ensemble"] = {k: ens[k] for k in ("value", "ms_per_step", "config", "roofline", "e2e", "gpu_launches")}
        line["ensemble"]["unit"] = "collisions/s"
        line["ensemble"]["steps"] = ens_steps
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_gas_baseline(args.gas_side, args.cpu_seconds, fma)
            line["ensemble"]["cpu_baseline"] = cpu_sinai_baseline(min(args.cpu_seconds, 5.0), fma)
    else:
        ens = bench_ensemble(args, device, fma, rank, world, args.steps, args.warmup)
        line.update({k: ens[k] for k in ("value", "ms_per_step", "config", "roofline", "e2e", "clocks", "gpu_launches")})
        line["ensemble"] = {"value": ens["value"], "unit": "collisions/s", "stats": ens["stats"]}
        line["note"] = ("single-system workload (dense_gas_16384) is not partitioned across GPUs: replicas only; "
                        "multi-GPU lines report the sharded Sinai ensemble")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
