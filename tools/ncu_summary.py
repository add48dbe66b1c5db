#!/usr/bin/env python3
"""Turn an `ncu --set full --import-source on` report into the markdown summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/evolve.ncu-rep "title" "command line" > profiles/<name>.md

Needs the `ncu` binary (reads the report with `--page raw --csv` and `--page source --csv`)."""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__icc_request_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"Command: `{cmd}`")
    print("(plain run first, exit 0; numbers below are from the profiled replay and are NOT bench values)\n")
    print("| metric | unit | value |\n|---|---|---|")
    for m in RAW:
        if m in col:
            print(f"| {m} | {units[col[m]]} | {vals[col[m]]} |")
    stalls = []
    for h, i in col.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
        if m:
            try:
                stalls.append((float(vals[i]), m.group(1)))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    print("\n## Warp stall reasons (stalled warps per issue-active cycle; share of the total)\n")
    print("| reason | warps per issue | share |\n|---|---|---|")
    for v, nme in sorted(stalls, reverse=True)[:9]:
        print(f"| {nme} | {v:.2f} | {100 * v / tot:.1f} % |")

    src = page(rep, "source")
    if len(src) > 2:
        h2 = src[1]
        i_src, i_s, i_ex = h2.index("Source"), h2.index("Warp Stall Sampling (All Samples)"), h2.index("Instructions Executed")
        data = []
        for r in src[2:]:
            try:
                data.append((int(r[i_s] or 0), int(r[i_ex] or 0), r[i_src].strip()))
            except (ValueError, IndexError):
                pass
        total = sum(d[0] for d in data) or 1
        texec = sum(d[1] for d in data) or 1
        print("\n## Hottest instructions (warp stall sampling)\n")
        print("| samples | executed | SASS |\n|---|---|---|")
        for s_, e_, t_ in sorted(data, reverse=True)[:14]:
            print(f"| {100 * s_ / total:.1f} % | {e_} | `{t_[:70]}` |")
        ops = collections.Counter()
        for _, e_, t_ in data:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", t_)
            ops[m.group(2) if m else "?"] += e_
        print("\n## Instruction mix (warp instructions executed)\n")
        print("| opcode | share |\n|---|---|")
        for k, v in ops.most_common(12):
            print(f"| {k} | {100 * v / texec:.1f} % |")


if __name__ == "__main__":
    main()
