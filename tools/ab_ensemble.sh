#!/bin/bash
# Same-box A/B of ensemble-kernel variants (how profiles/r1_ensemble_ab.log was made).
#   1. here:  tools/ab_ensemble.sh build NAME path/to/variant_of_bl_ensemble.cu     (repeat per variant)
#             -> tools/bin/lib_NAME.so = the in-tree objects with that one file swapped (tools/bin/ is git-ignored but
#                travels with gpurun)
#   2. here:  gpurun --timeout 200 -- 'tools/ab_ensemble.sh run NAME1 NAME2 ...'
#             -> runs the in-tree library first, then each variant, 2^20 systems x 10^4 collisions, and compares
#                t / v / p0 / hits of every system bit for bit with the in-tree result; prints the kernel times
# Remove tools/bin/lib_*.so afterwards (they are 10 MB each and ride along on every gpurun push).
set -e
cd "$(dirname "$0")/.."
CSRC=python-billiards_b200/csrc
if [ "$1" = build ]; then
    mkdir -p tools/bin
    cp "$3" $CSRC/_variant_$2.cu
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true \
         -ftz=false -Xcompiler -fPIC -Xptxas -v -c $CSRC/_variant_$2.cu -o /tmp/_variant_$2.o 2>&1 | grep -E "error|spill" | sort | uniq -c
    rm -f $CSRC/_variant_$2.cu
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -lcudart -o tools/bin/lib_$2.so \
         $CSRC/bl_api.o $CSRC/bl_evolve.o $CSRC/bl_evolve_small.o $CSRC/bl_table.o $CSRC/bl_batch.o /tmp/_variant_$2.o
    ls -la tools/bin/lib_$2.so
elif [ "$1" = run ]; then
    shift
    python tools/ensemble_crosscheck.py run /tmp/ab_base.npz 1048576 10000
    for v in "$@"; do
        BILLIARDS_B200_LIB=$PWD/tools/bin/lib_$v.so python tools/ensemble_crosscheck.py run /tmp/ab_$v.npz 1048576 10000
        echo "$v vs in-tree: $(python tools/ensemble_crosscheck.py compare /tmp/ab_$v.npz /tmp/ab_base.npz | tr '\n' ' ')"
    done
else
    sed -n 2,12p "$0"
fi
