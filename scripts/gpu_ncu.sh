#!/bin/bash
# ncu full capture of selected kernels. Usage: bash scripts/gpu_ncu.sh <tag> <kernel-regex> [pairs]
TAG=$1; KRE=$2; PAIRS=${3:-500000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 2 -c 3 -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --pairs $PAIRS --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
