#!/bin/bash
# r39: A/B at C3 (and the C4 / C5 shapes) of kb_unique_tail's one-strand path (KB_SEED_TAIL_FAST).
TAG=${1:-r39}
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import parity_util as pu
print(pu.ensure_syn_index(3100, 24, 12345))
PY
PREFIX=data/_gen/syn/syn3100
python scripts/gpu_ab.py --pairs 1250000 --prefix $PREFIX --error 0.01 --reps 4 --no-e2e base: tail0:KB_SEED_TAIL_FAST=0 base2: tail0b:KB_SEED_TAIL_FAST=0 > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_ab.jsonl"):
    d = json.loads(ln); print(d["config"], d["device_ms"], {k: d["stage_ms"][k] for k in ("fm_seed", "cand_pair", "rescue", "segments", "align", "assemble", "finalize")}, d["same_result"])
PY
for v in "KB_SEED_TAIL_FAST=1" "KB_SEED_TAIL_FAST=0"; do
  echo "== $v" >> gpurun_out/${TAG}_modes.jsonl
  env $v python scripts/gpu_modes.py --prefixes $PREFIX --modes pacbio --se 0 --pb 50000 --ref-se 0 --ref-pb 0 --check 50 >> gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_ab.err
done
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_modes.jsonl"):
    if ln.startswith("=="): print(ln.strip()); continue
    d = json.loads(ln); print(" ", d["mode"], round(d["device_ms"], 2), {k: round(v, 2) for k, v in d["stage_ms"].items()}, d.get("oracle_mismatches"))
PY
tail -n 3 gpurun_out/${TAG}_ab.err
