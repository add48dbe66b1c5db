#!/usr/bin/env python3
"""profiles/r2_scaling.md from the bench lines of a scaling series (gpurun_out/s_<N>_<weak|strong>.json, s_8_pool.json):
    python tools/scaling_table.py gpurun_out profiles"""
import json
import os
import sys

src, dst = sys.argv[1], sys.argv[2]


def load(path):
    d = None
    for line in open(path):
        if line.startswith("{"):
            d = json.loads(line)
    return d


rows = []
for n in (1, 2, 4, 8):
    for sc in ("weak", "strong"):
        p = os.path.join(src, f"s_{n}_{sc}.json")
        if os.path.exists(p):
            d = load(p)
            rows.append(d)
            json.dump(d, open(os.path.join(dst, f"r2_bench_n{n}_{sc}.json"), "w"))
base = rows[0]
out = ["# Round 2 -- scaling of the Sinai ensemble on one B200 box (bench.py under torch.distributed.run, 5 steps after 3 warm-ups)",
       "",
       "Every line is the same workload (config S: 10^4 collisions of every system per step, contiguous shards, no exchange during the",
       "run, one NCCL all-reduce of the statistics per step); device-timed with CUDA events, max over ranks.  Efficiency = value / (N x the",
       "N = 1 value).  JSON lines: `profiles/r2_bench_n<N>_{weak,strong}.json`.",
       "",
       "| GPUs | scaling | systems per GPU | collisions/s (device) | ms per step | kernel ms | efficiency | e2e collisions/s | e2e efficiency | SM clock, throttle reasons |",
       "|---|---|---|---|---|---|---|---|---|---|"]
for d in rows:
    n = d["n_gpus"]
    cl = d.get("clocks") or {}
    out.append(f"| {n} | {d['scaling'] if n > 1 else 'weak = strong'} | {d['config']['systems_per_gpu']} | {d['value'] / 1e9:.1f} G | "
               f"{d['ms_per_step']:.2f} | {d['roofline']['kernel_ms_per_launch']:.2f} | {d['value'] / (n * base['value']):.3f} | "
               f"{d['e2e']['value'] / 1e9:.1f} G | {d['e2e']['value'] / (n * base['e2e']['value']):.3f} | {cl.get('sm_mhz')} MHz, {cl.get('reasons')} |")
p = os.path.join(src, "s_8_pool.json")
if os.path.exists(p):
    d = load(p)
    json.dump(d, open(os.path.join(dst, "r2_bench_n8_pool_ensemble.json"), "w"))
    out += ["", f"Multi-ball ensemble, 8 GPUs, weak (2^16 pool tables of 16 balls per GPU, every table from the break shot to t = 100 in one "
                f"launch): {d['value'] / 1e9:.2f} G collisions/s device-timed ({d['ms_per_step']:.2f} ms per step), e2e {d['e2e']['value'] / 1e9:.2f} G."]
out += ["",
        "Strong scaling at 8 GPUs runs 2^17 systems per GPU = 1024 CTAs, all resident from the start (148 SMs x 6 CTAs = 888 slots, the rest follows",
        "at once): the warps of an SM finish at different times (wall and disk hits diverge) and the SM drains at falling occupancy, with no further",
        "wave to fill it -- the kernel takes longer than 1/8 of the one-GPU kernel -- plus ~0.2 ms of launch and all-reduce per step.  This is the part",
        "weak scaling hides.  The e2e column moves 32 B in and 48 B out per system through pinned memory in four pipelined blocks per GPU",
        "(`ensemble.run_from_host`)."]
open(os.path.join(dst, "r2_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
