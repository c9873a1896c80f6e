#!/bin/bash
# r12: A/B of the warp-aggregated allocators + forked NW size classes against the r10 build, pipeline trace, ncu of phase B
TAG=${1:-r12}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
python scripts/gpu_ab.py base: nostreams:KB_NW_STREAMS=0 w16:KB_ALIGN_WARPS=2368 w20:KB_ALIGN_WARPS=2960 sub400k:KB_PIPE_SUB_READS=400000 sub500k:KB_PIPE_SUB_READS=500000 sub1m:KB_PIPE_SUB_READS=1000000 minb8:KB_SEED_MINB=8 > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err; cat gpurun_out/${TAG}_ab.jsonl; tail -3 gpurun_out/${TAG}_ab.err
python scripts/gpu_ab.py --lib build/r10/libkartb200.so r10: > gpurun_out/${TAG}_ab_r10.jsonl 2>> gpurun_out/${TAG}_ab.err; cat gpurun_out/${TAG}_ab_r10.jsonl
KB_PIPE_TRACE=1 python scripts/gpu_ab.py --reps 1 trace: > /dev/null 2> gpurun_out/${TAG}_pipe_trace.txt; tail -14 gpurun_out/${TAG}_pipe_trace.txt
SYN=data/_gen/syn/syn100
python scripts/gpu_ab.py --prefix $SYN --error 0.01 base: nostreams:KB_NW_STREAMS=0 w16:KB_ALIGN_WARPS=2368 > gpurun_out/${TAG}_ab_syn100.jsonl 2>> gpurun_out/${TAG}_ab.err; cat gpurun_out/${TAG}_ab_syn100.jsonl
python scripts/gpu_ab.py --prefix $SYN --error 0.01 --lib build/r10/libkartb200.so r10: >> gpurun_out/${TAG}_ab_syn100.jsonl 2>> gpurun_out/${TAG}_ab.err; tail -1 gpurun_out/${TAG}_ab_syn100.jsonl
python scripts/gpu_modes.py --se 1000000 --pb 5000 --ref-se 0 --ref-pb 0 --check 0 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_ab.err; cat gpurun_out/${TAG}_modes.jsonl
KART_B200_LIB=build/r10/libkartb200.so python scripts/gpu_modes.py --se 1000000 --pb 5000 --ref-se 0 --ref-pb 0 --check 0 > gpurun_out/${TAG}_modes_r10.jsonl 2>> gpurun_out/${TAG}_ab.err; cat gpurun_out/${TAG}_modes_r10.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_segments$|k_align_part|k_nw_tile|k_align_gather|k_assemble$|k_cand_pair' -s 11 -c 11 -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
ls -la gpurun_out
