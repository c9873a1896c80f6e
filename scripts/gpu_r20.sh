#!/bin/bash
# r20: GPU index builder (tests + the 3.1 Gbp genome from its .pac, compared with the host-built files), rescue A/B, slot-pipeline
# plans at C3 with packed reads, CLI start-up at C3.
TAG=${1:-r20}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_index_build.py -m gpu -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_index.txt; cat gpurun_out/${TAG}_pytest_index.txt
PREFIX=data/_gen/syn/syn3100
python - <<'PY' > gpurun_out/${TAG}_index_host.txt 2>&1
import sys, time; sys.path.insert(0, "tests")
import parity_util as pu
t = time.time(); p = pu.ensure_syn_index(3100, 24, 12345); print("host-built index", p, "%.1f s" % (time.time() - t))
PY
cat gpurun_out/${TAG}_index_host.txt
mkdir -p /tmp/gidx && cp $PREFIX.pac $PREFIX.ann $PREFIX.amb /tmp/gidx/ && for e in pac ann amb; do mv /tmp/gidx/syn3100.$e /tmp/gidx/g.$e; done
( time KB_INDEX_TRACE=1 kart_b200/bin/kart index -gpu -pac /tmp/gidx/g ) > gpurun_out/${TAG}_index_gpu.txt 2>&1; tail -25 gpurun_out/${TAG}_index_gpu.txt
cmp /tmp/gidx/g.bwt $PREFIX.bwt && echo "BWT IDENTICAL" >> gpurun_out/${TAG}_index_gpu.txt; cmp /tmp/gidx/g.sa $PREFIX.sa && echo "SA IDENTICAL" >> gpurun_out/${TAG}_index_gpu.txt; tail -2 gpurun_out/${TAG}_index_gpu.txt
rm -rf /tmp/gidx
python bench.py --steps 3 --warmup 2 --cpu-sample-pairs 0 --program-pairs 1000000 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r20_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms  e2e %.3f ms  e2e_text %.3f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_text"]["ms_per_step"]), {k: round(v, 3) for k, v in d["stage_ms"].items()}, "pack", d["e2e"]["host_pack_ms_outside_timed_region"]); print(d.get("e2e_program"))
PY
SWEEP_PREFIX=$PREFIX SWEEP_ERR=0.01 SWEEP_PACKED=1 python scripts/gpu_pipe_sweep.py 1250000 2500002 625000,312500,200,0 1250000 1250000,250000,500,250000 1000000,250000,400,250000 833334 > gpurun_out/${TAG}_pipe_sweep.txt 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_pipe_sweep.txt
KB_PIPE_TRACE=1 SWEEP_PREFIX=$PREFIX SWEEP_ERR=0.01 SWEEP_PACKED=1 python scripts/gpu_pipe_sweep.py 1250000 625000,312500,200,0 2> gpurun_out/${TAG}_pipe_trace.txt | tail -2; tail -12 gpurun_out/${TAG}_pipe_trace.txt
KART_B200_TRACE=1 python scripts/cli_compare.py --pairs 2500000 --prefix $PREFIX --error 0.01 --ours-only > gpurun_out/${TAG}_cli_c3.json 2> gpurun_out/${TAG}_cli_trace.txt; cat gpurun_out/${TAG}_cli_c3.json; grep -v "read \|format" gpurun_out/${TAG}_cli_trace.txt | tail -8
tail -12 gpurun_out/${TAG}_bench.err
