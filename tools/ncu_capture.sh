#!/bin/bash
# One `ncu --set full` capture of the hot kernels with source correlation; exports raw metrics and the
# per-line source page as CSV into gpurun_out/ (the .ncu-rep itself is too large to travel).
#   tools/ncu_capture.sh <tag> [kernel-regex] [env assignments...]
tag=$1; shift
regex=${1:-"k_col3|k_col2|k_row_inv_prox3"}; shift
for kv in "$@"; do export "$kv"; done
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:"$regex" -s 40 -c 2 -f -o /tmp/ncu/$tag \
    python bench.py --steps 28 --warmup 3 --no-cpu --no-configs > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>> gpurun_out/${tag}_ncu.log
ncu -i /tmp/ncu/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>> gpurun_out/${tag}_ncu.log
ls -la /tmp/ncu gpurun_out/${tag}_* >> gpurun_out/${tag}_ncu.log
tail -3 gpurun_out/${tag}_ncu.log
