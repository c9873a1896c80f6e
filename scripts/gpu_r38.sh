#!/bin/bash
# r38: A/B at C3 of the four-seeds-per-round candidate scan of the heavy items (KB_CAND_WIDE).
TAG=${1:-r38}
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import parity_util as pu
print(pu.ensure_syn_index(3100, 24, 12345))
PY
PREFIX=data/_gen/syn/syn3100
python scripts/gpu_ab.py --pairs 1250000 --prefix $PREFIX --error 0.01 --reps 4 --no-e2e base: wide0:KB_CAND_WIDE=0 base2: wide0b:KB_CAND_WIDE=0 > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_ab.jsonl"):
    d = json.loads(ln); print(d["config"], d["device_ms"], {k: d["stage_ms"][k] for k in ("fm_seed", "cand_pair", "rescue", "segments", "align", "assemble", "finalize")}, d["same_result"])
PY
tail -n 3 gpurun_out/${TAG}_ab.err
