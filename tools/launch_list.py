#!/usr/bin/env python3
"""Turn the csv of `ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv <cmd>`
into the markdown launch list kept under profiles/ (launches shorter than 0.2 ms are summed per kernel at the end).
    python tools/launch_list.py launches.csv "title" "command" > profiles/<name>.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
print(f"# {sys.argv[2]}\n")
print(f"Command: `{sys.argv[3]}`")
print("(per-launch times are serialised and cold-cache; what must agree with bench.py is each kernel's SHARE of its step)\n")
print("| # | kernel | grid | block | duration |\n|---|---|---|---|---|")
small = collections.defaultdict(lambda: [0, 0.0])
total = collections.defaultdict(float)
for r in rows:
    name = re.sub(r"^void ", "", r[4])
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)[:70]
    ms = float(r[14]) / 1e6
    total[name] += ms
    if ms >= 0.2:
        print(f"| {r[0]} | {name} | {r[8]} | {r[7]} | {ms:.3f} ms |")
    else:
        small[name][0] += 1
        small[name][1] += ms
print("\n## Launches shorter than 0.2 ms, summed\n\n| kernel | launches | total |\n|---|---|---|")
for k, (c, ms) in sorted(small.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {c} | {ms:.3f} ms |")
allms = sum(total.values())
print("\n## Share of the profiled run by kernel\n\n| kernel | total | share |\n|---|---|---|")
for k, ms in sorted(total.items(), key=lambda kv: -kv[1])[:8]:
    print(f"| {k} | {ms:.3f} ms | {100 * ms / allms:.1f} % |")
