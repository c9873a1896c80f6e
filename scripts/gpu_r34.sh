#!/bin/bash
# r34 (gpurun --gpus 2): the multi-device tests (CLI worker pool on two GPUs byte-identical to kart -t 1, index clone device to device), bench.py under
# torchrun at N = 2 on the C3 workload as the driver launches it, whole program at C3 on two GPUs with the stage trace.
TAG=${1:-r34}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt; cat gpurun_out/${TAG}_gpus.txt
( python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or clone_index or multi_device" 2>&1 | tail -8 ) > gpurun_out/${TAG}_pytest_multi.txt 2>&1; cat gpurun_out/${TAG}_pytest_multi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print("N=%d device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)" % (d["n_gpus"], d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6), d["config"].get("host_binding"))
PY
tail -n 3 gpurun_out/${TAG}_bench_n2.err
PREFIX=data/_gen/syn/syn3100
KART_B200_TRACE=1 python scripts/cli_compare.py --pairs 4000000 --prefix $PREFIX --error 0.01 --ours-only --extra "--gpus 2 --full-sa" > gpurun_out/${TAG}_cli_c3_g2.json 2> gpurun_out/${TAG}_cli_trace_g2.txt; cat gpurun_out/${TAG}_cli_c3_g2.json; grep -c "gpu1 " gpurun_out/${TAG}_cli_trace_g2.txt; grep "device 1 up\|index uploaded\|kb_init" gpurun_out/${TAG}_cli_trace_g2.txt | tail -4; grep "gpu[01] " gpurun_out/${TAG}_cli_trace_g2.txt | tail -8
