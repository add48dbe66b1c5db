#!/usr/bin/env python3
"""profiles/r2_dram_traffic.json from `ncu --set full` reports: dram__bytes_read.sum + dram__bytes_write.sum of the
dominant kernels, per collision (event kernel) and per system (ensemble kernel).  bench.py reads the file for
`roofline.traffic`.

    python tools/ncu_traffic.py --evolve gpurun_out/x.ncu-rep --evolve-collisions 1000000 --balls 16384 \
                                --ensemble gpurun_out/y.ncu-rep --ensemble-systems 1048576
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if kernel_regex not in name:
            continue
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            k = hdr.index(key)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[k]]
            tot += float(r[k].replace(",", "")) * scale
        dur = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
        if best is None or dur > best[1]:
            best = (tot, dur, name, units[hdr.index("gpu__time_duration.sum")])
    return best


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--evolve")
    ap.add_argument("--evolve-collisions", type=int, default=10**6)
    ap.add_argument("--balls", type=int, default=16384)
    ap.add_argument("--ensemble")
    ap.add_argument("--ensemble-systems", type=int, default=1 << 20)
    a = ap.parse_args()
    path = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    if a.evolve:
        tot, dur, name, unit = dram_bytes(a.evolve, "bl_evolve_kernel")
        d[f"evolve_dram_bytes_per_collision_{a.balls}"] = tot / a.evolve_collisions
        d["evolve_source"] = (f"ncu --set full, {os.path.basename(a.evolve)}: dram__bytes_read.sum + dram__bytes_write.sum = "
                              f"{tot:.4g} B for one launch of {a.evolve_collisions} collisions at N = {a.balls} ({name[:60]})")
    if a.ensemble:
        tot, dur, name, unit = dram_bytes(a.ensemble, "bl_ensemble_kernel")
        d["ensemble_dram_bytes_per_system"] = tot / a.ensemble_systems
        d["ensemble_source"] = (f"ncu --set full, {os.path.basename(a.ensemble)}: dram__bytes_read.sum + dram__bytes_write.sum = "
                                f"{tot:.4g} B for one launch over {a.ensemble_systems} systems ({name[:60]})")
    json.dump(d, open(path, "w"), indent=1)
    print(json.dumps(d, indent=1))
