#!/bin/bash
# r18: full -m gpu suite (stage-level tests, C1 md5, pipelined chunk vs oracle, C3/C4/C5 on the 3.1 Gbp index built on the box),
# bench.py at its default (C3) with the reference arm, A/B of the warp-per-window rescue and of the L2 fetch granularity,
# ncu launch list + full capture at C3.
TAG=${1:-r18}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt; nproc >> gpurun_out/${TAG}_gpu.txt; free -g | head -2 >> gpurun_out/${TAG}_gpu.txt
( time python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cat gpurun_out/${TAG}_pytest.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench.err; tail -c 1200 gpurun_out/${TAG}_bench_ref.json
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err; tail -c 4500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
for v in "KB_RESCUE_FAST=0" "KB_L2_FETCH=32" "KB_L2_FETCH=128"; do
  echo "== $v" >> gpurun_out/${TAG}_ab.jsonl
  env $v python bench.py --steps 3 --warmup 2 --cpu-sample-pairs 0 --program-pairs 0 >> gpurun_out/${TAG}_ab.jsonl 2>> gpurun_out/${TAG}_bench.err
done
python - <<'PY'
import json
for ln in open("gpurun_out/r18_ab.jsonl"):
    if ln.startswith("=="): print(ln.strip()); continue
    d = json.loads(ln); print(round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}c3_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_rescue|k_cand_pair|k_segments$|k_align_part|k_assemble$|k_finalize' -s 10 -c 10 -o gpurun_out/${TAG}c3_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}c3_ncu_full.log | cut -c1-200
python scripts/ncu_summary.py ${TAG}c3 500000 syn3100 | tail -2
cp profiles/${TAG}c3_* gpurun_out/ 2>/dev/null
tail -20 gpurun_out/${TAG}_bench.err
