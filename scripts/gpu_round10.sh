#!/bin/bash
# r10: full C2 round + 100 Mbp (HBM regime) bench/profile + whole-program CLI comparison
TAG=${1:-r10}
bash scripts/gpu_round.sh $TAG
SYN=data/_gen/syn/syn100
if [ -f $SYN.bwt ]; then
python bench.py --steps 5 --warmup 3 --prefix $SYN --error 0.01 --cpu-sample-pairs 100000 > gpurun_out/${TAG}_bench_syn100.json 2>> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench_syn100.json
bash scripts/gpu_prof.sh ${TAG}syn 'k_fm_seed|k_rescue|k_segments|k_align' 500000 --prefix $SYN --error 0.01
fi
python scripts/cli_compare.py --pairs 1000000 --t1 --extra=--full-sa > gpurun_out/${TAG}_cli_c2.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_cli_c2.json
ls -la gpurun_out
