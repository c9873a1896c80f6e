#!/bin/bash
# r33: the seeding kernel with a read's packed words staged in shared memory (KB_SEED_STAGE): parity of the load paths, A/B at C3 with the
# same-result signature, the kernel's full ncu row, C4 / C5, and the bench as the driver runs it (three chunks in flight, whole-program leg on the step's batch).
TAG=${1:-r33}
mkdir -p gpurun_out
PREFIX=data/_gen/syn/syn3100
( time python -m pytest tests -m gpu -q -k "seeding_load or rescue_sampled or c3 or C3" 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cut -c1-1500 gpurun_out/${TAG}_pytest.txt | tail -8
python scripts/gpu_ab.py --pairs 1250000 --prefix $PREFIX --error 0.01 --reps 4 --no-e2e \
  stage_hint: nostage:KB_SEED_STAGE=0 stage_nohint:KB_SEED_LD_HINT=0 neither:KB_SEED_STAGE=0,KB_SEED_LD_HINT=0 qp12:KB_SEED_QP=12 qs8:KB_SEED_QS=8 trips4:KB_SEED_TRIPS=4 \
  > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_ab.jsonl"):
    d = json.loads(ln); print(d["config"], d["device_ms"], {k: d["stage_ms"][k] for k in ("fm_seed", "sa_locate", "rescue", "cand_pair", "segments")}, d["same_result"])
PY
ncu --set full --clock-control none -k regex:'k_fm_seed' -s 2 -c 1 -o /tmp/${TAG}seed -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}_ncu_seed.log 2>&1
ncu -i /tmp/${TAG}seed.ncu-rep --page raw --csv > gpurun_out/${TAG}_seed_raw_stage.csv 2>/dev/null
python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 0 --ref-pb 0 --check 100 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_modes.jsonl"):
    d = json.loads(ln); print(" ", d["mode"], round(d["device_ms"], 2), {k: round(v, 2) for k, v in d["stage_ms"].items()}, d.get("oracle_mismatches"))
PY
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"], d["roofline"]["frac"], d["roofline"]["traffic_source"])
print(d["cpu_baseline"]); print(d.get("e2e_program"))
PY
grep real gpurun_out/${TAG}_bench.err; tail -n 5 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_ab.err
