#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list, ncu full capture of the top kernels.  Usage: bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
if [ "$2" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_segments|k_align|k_assemble|k_sa_locate|k_rescue|k_cand_pair|k_finalize' -s 8 -c 8 -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 1000 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
fi
