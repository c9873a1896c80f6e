#!/bin/bash
# r16: parity + bench of HEAD (r14 + task-parallel rescue, unique-tail seeding with lane queue, CLI start-up),
# whole-program CLI comparison, SE / -pacbio shapes, 100 Mbp bench, ncu launch list + full captures (C2 and 100 Mbp)
TAG=${1:-r16}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
nproc >> gpurun_out/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
python scripts/cli_compare.py --pairs 1000000 --t1 --diff-out gpurun_out/${TAG}_cli_c2_diff.txt > gpurun_out/${TAG}_cli_c2.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_cli_c2.json
SYN=data/_gen/syn/syn100
if [ -f $SYN.bwt ]; then
python bench.py --steps 5 --warmup 3 --prefix $SYN --error 0.01 --cpu-sample-pairs 100000 > gpurun_out/${TAG}_bench_syn100.json 2>> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench_syn100.json
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_segments$|k_cand_pair|k_align_part|k_nw_tile|k_nw_warp|k_align_gather|k_assemble$|k_rescue|k_finalize' -s 17 -c 17 -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
if [ -f $SYN.bwt ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}syn_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 --prefix $SYN --error 0.01 > gpurun_out/${TAG}syn_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_rescue|k_cand_pair|k_segments$|k_align_part' -s 7 -c 7 -o gpurun_out/${TAG}syn_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 --prefix $SYN --error 0.01 > gpurun_out/${TAG}syn_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}syn_ncu_full.log | cut -c1-200
fi
tail -20 gpurun_out/${TAG}_bench.err
ls -la gpurun_out
python scripts/gpu_modes.py --se 1000000 --pb 20000 --ref-se 100000 --ref-pb 500 --check 200 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_modes.jsonl
