#!/bin/bash
# usage: profiles/ncu_one.sh <kernel-regex> <out-name> [skip] [count]   (run under gpurun, 1 GPU)
# One `ncu --set full` capture of a kernel family of the C2 bench step (B200_PROFILING.md recipe).
K=$1; OUT=$2; SKIP=${3:-1}; CNT=${4:-1}
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/$OUT \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/$OUT.log 2>&1
