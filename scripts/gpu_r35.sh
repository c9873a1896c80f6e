#!/bin/bash
# r35: A/B at C3 of k_finalize on local copies (KB_FIN_LOCAL), seeding quorums next to the four-trip default, the segx slab, rescue ticket batches.
TAG=${1:-r35}
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import parity_util as pu
print(pu.ensure_syn_index(3100, 24, 12345))
PY
PREFIX=data/_gen/syn/syn3100
python scripts/gpu_ab.py --pairs 1250000 --prefix $PREFIX --error 0.01 --reps 4 --no-e2e \
  base: fin0:KB_FIN_LOCAL=0 qp12:KB_SEED_QP=12 qs8:KB_SEED_QS=8 qp12qs8:KB_SEED_QP=12,KB_SEED_QS=8 trips6:KB_SEED_TRIPS=6 slab256:KB_SEG_SLAB=256 rfb16:KB_RF_BATCH=16 base2: \
  > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_ab.jsonl"):
    d = json.loads(ln); print(d["config"], d["device_ms"], {k: d["stage_ms"][k] for k in ("fm_seed", "rescue", "segments", "assemble", "finalize")}, d["same_result"])
PY
tail -n 3 gpurun_out/${TAG}_ab.err
