#!/bin/bash
# usage: tools/gpurun_retry_n.sh <ngpus> <timeout-seconds> '<command>'   -- multi-GPU variant of gpurun_retry.sh
N=$1; T=$2; shift 2
for i in $(seq 1 80); do
    OUT=$(/usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@" 2>&1)
    if echo "$OUT" | grep -q "status=transient"; then sleep 45; continue; fi
    echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
