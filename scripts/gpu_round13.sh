#!/bin/bash
# r13: parity + bench of HEAD, whole-program CLI comparison with the three-stage host pipeline (stage trace), SE / -pacbio shapes,
# 100 Mbp HBM-regime bench, ncu launch list + full capture
TAG=${1:-r13}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
nproc >> gpurun_out/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
python scripts/cli_compare.py --pairs 1000000 --t1 --extra=--full-sa > gpurun_out/${TAG}_cli_c2.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_cli_c2.json
KART_B200_TRACE=1 python scripts/cli_compare.py --pairs 1000000 --extra=--full-sa --ours-only > gpurun_out/${TAG}_cli_c2_trace.json 2> gpurun_out/${TAG}_cli_trace.txt; cat gpurun_out/${TAG}_cli_c2_trace.json; tail -40 gpurun_out/${TAG}_cli_trace.txt
SYN=data/_gen/syn/syn100
if [ -f $SYN.bwt ]; then
python bench.py --steps 5 --warmup 3 --prefix $SYN --error 0.01 --cpu-sample-pairs 100000 > gpurun_out/${TAG}_bench_syn100.json 2>> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench_syn100.json
python scripts/cli_compare.py --pairs 1000000 --prefix $SYN --error 0.01 > gpurun_out/${TAG}_cli_syn100.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_cli_syn100.json
fi
python scripts/gpu_modes.py --se 1000000 --pb 5000 --ref-se 100000 --ref-pb 500 --check 200 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_modes.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_segments$|k_cand_pair|k_align_part|k_nw_tile|k_align_gather|k_assemble$|k_rescue' -s 13 -c 13 -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
tail -20 gpurun_out/${TAG}_bench.err
ls -la gpurun_out
