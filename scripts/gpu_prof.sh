#!/bin/bash
# ncu launch list + full capture of selected kernels. Usage: bash scripts/gpu_prof.sh <tag> <kernel-regex> [pairs] [extra bench args]
TAG=$1; KRE=$2; PAIRS=${3:-500000}; shift 3
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --pairs $PAIRS --cpu-sample-pairs 1000 "$@" > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 10 -c 10 -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --pairs $PAIRS --cpu-sample-pairs 1000 "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
