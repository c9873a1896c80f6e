#!/bin/bash
# r37: sweep at C3 around the new defaults: segx slab size, seeding trips.
TAG=${1:-r37}
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import parity_util as pu
print(pu.ensure_syn_index(3100, 24, 12345))
PY
PREFIX=data/_gen/syn/syn3100
python scripts/gpu_ab.py --pairs 1250000 --prefix $PREFIX --error 0.01 --reps 4 --no-e2e \
  base: slab128:KB_SEG_SLAB=128 slab512:KB_SEG_SLAB=512 slab1024:KB_SEG_SLAB=1024 trips8:KB_SEED_TRIPS=8 trips12:KB_SEED_TRIPS=12 qp6:KB_SEED_QP=6 base2: \
  > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_ab.jsonl"):
    d = json.loads(ln); print(d["config"], d["device_ms"], {k: d["stage_ms"][k] for k in ("fm_seed", "rescue", "segments", "align", "assemble", "finalize")}, d["same_result"])
PY
tail -n 3 gpurun_out/${TAG}_ab.err
