#!/bin/bash
# r22: full -m gpu suite (index-builder fix, staged k_unpack, split k_cand_heavy, windowed kb_pair, grouped rescue scan, 16-byte seeding table),
# seeding schedule / table size / heavy-list A/B at C3, bench at its default.
TAG=${1:-r22}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cut -c1-1500 gpurun_out/${TAG}_pytest.txt | tail -30
PREFIX=data/_gen/syn/syn3100
python scripts/gpu_ab.py --prefix $PREFIX --error 0.01 --pairs 1250000 --reps 3 --no-e2e base: q12:KB_SEED_QP=12,KB_SEED_QS=6 q16:KB_SEED_QP=16,KB_SEED_QS=8 q24:KB_SEED_QP=24,KB_SEED_QS=12 \
  q16t1:KB_SEED_QP=16,KB_SEED_QS=8,KB_SEED_TRIPS=1 q16t4:KB_SEED_QP=16,KB_SEED_QS=8,KB_SEED_TRIPS=4 q8s8:KB_SEED_QP=8,KB_SEED_QS=8 tail4:KB_SEED_TAIL=4 k15:KB_KTAB_K=15 k13:KB_KTAB_K=13 noheavy:KB_CAND_HEAVY=0 \
  > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err; cut -c1-420 gpurun_out/${TAG}_ab.jsonl; tail -3 gpurun_out/${TAG}_ab.err
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"])
print(d["cpu_baseline"]); print(d.get("e2e_program"))
PY
tail -5 gpurun_out/${TAG}_bench.err
