#!/bin/bash
# Usage (GPU box): bash profiles/cap_kernel.sh TAG KERNEL_REGEX [SKIP] [COUNT]  — one `ncu --set full` capture of a steady-state launch of one kernel
# (bench warm-up frames are skipped with --launch-skip counted per matching kernel).
TAG=$1; K=$2; SKIP=${3:-4}; COUNT=${4:-1}
mkdir -p gpurun_out
timeout 280 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $SKIP -c $COUNT -f -o gpurun_out/${TAG} python bench.py --steps 2 --warmup 5 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}.err
ls -la gpurun_out/${TAG}.ncu-rep
