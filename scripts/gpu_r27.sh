#!/bin/bash
# r27 (run with gpurun --gpus 2): the multi-device tests (CLI worker pool on two GPUs byte-identical to kart -t 1, index clone device to device),
# bench.py under torchrun at N = 2 on the C3 workload (both arms, as the driver launches them), whole program at C3 on one and on two GPUs.
TAG=${1:-r27}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt
( python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or clone_index" 2>&1 | tail -8 ) > gpurun_out/${TAG}_pytest_multi.txt 2>&1; cat gpurun_out/${TAG}_pytest_multi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print("N=%d device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)" % (d["n_gpus"], d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6), d["config"].get("host_binding"))
PY
tail -3 gpurun_out/${TAG}_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample-pairs 200000 > gpurun_out/${TAG}_bench_ref_n2.json 2>> gpurun_out/${TAG}_bench_n2.err; cut -c1-300 gpurun_out/${TAG}_bench_ref_n2.json
PREFIX=data/_gen/syn/syn3100
for n in 1 2; do
  KART_B200_TRACE=1 python scripts/cli_compare.py --pairs 4000000 --prefix $PREFIX --error 0.01 --ours-only --extra "--gpus $n --full-sa" > gpurun_out/${TAG}_cli_c3_g$n.json 2> gpurun_out/${TAG}_cli_trace_g$n.txt; cat gpurun_out/${TAG}_cli_c3_g$n.json; grep -c "gpu1 " gpurun_out/${TAG}_cli_trace_g$n.txt
done
