#!/usr/bin/env python3
"""Static instruction mix of one kernel (or one address range of it) from a compiled object -- what was used to trim
bl_ensemble_kernel_u before spending GPU time (B200_PROFILING.md: check cuobjdump -sass here first).

    python tools/sass_mix.py python-billiards_b200/csrc/bl_ensemble.o 'kernel_uILb1ELi4ELi1E' [0x5e0 0x2540]

Prints the opcode histogram (predicated instructions counted under their opcode), the FP64-pipe share
(DADD/DMUL/DFMA/DSETP/MUFU.*64H) and every branch / call with its target, so that loop bounds can be read off."""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    inside, rows = False, []
    for line in sass.splitlines():
        if "Function :" in line:
            inside = re.search(pat, line) is not None
            if inside:
                print(line.strip())
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if inside and m:
            addr, text = int(m.group(1), 16), m.group(2).strip()
            if lo <= addr <= hi:
                rows.append((addr, text))
    mix = collections.Counter()
    for addr, text in rows:
        tok = text.split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        mix[op.split(".")[0]] += 1
        if op.startswith(("BRA", "CALL", "EXIT", "RET")):
            print(f"  {addr:#06x}  {text}")
    n = sum(mix.values())
    fp64 = sum(c for op, c in mix.items() if op in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU"))
    print(f"{n} instructions, {fp64} on the FP64 / XU pipes ({100.0 * fp64 / max(n, 1):.0f} %)")
    print("  ".join(f"{op} {c}" for op, c in mix.most_common(24)))


if __name__ == "__main__":
    main()
