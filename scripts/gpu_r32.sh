#!/bin/bash
# r32: the sampled rescue scan (KB_RF_STRIDE), the kept mate index (KB_RF_REUSE / KB_RF_BATCH) and the L1 no-allocate seeding loads (KB_SEED_LD_HINT)
# on the GPU: parity tests of the rescue paths and of C3 on the 3.1 Gbp index, an A/B at C3 (same-result signature per configuration), the bench with
# two and three chunks in flight, C4 / C5 with the load hint, and the CLI stage trace (pin / pack / map per batch) at C3.
TAG=${1:-r32}
mkdir -p gpurun_out
PREFIX=data/_gen/syn/syn3100
( time python -m pytest tests -m gpu -q -k "rescue or c3 or C3" 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cut -c1-1500 gpurun_out/${TAG}_pytest.txt | tail -8
python scripts/gpu_ab.py --pairs 1250000 --prefix $PREFIX --error 0.01 --reps 4 \
  old:KB_RF_STRIDE=1,KB_RF_REUSE=0,KB_RF_BATCH=1 stride3:KB_RF_REUSE=0,KB_RF_BATCH=1 reuse4: reuse8:KB_RF_BATCH=8 reuse2:KB_RF_BATCH=2 \
  hint:KB_SEED_LD_HINT=1 hint_old:KB_SEED_LD_HINT=1,KB_RF_STRIDE=1,KB_RF_REUSE=0,KB_RF_BATCH=1 \
  > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_ab.jsonl"):
    d = json.loads(ln); print(d["config"], d["device_ms"], "e2e", d["e2e_ms"], {k: d["stage_ms"][k] for k in ("fm_seed", "rescue", "cand_pair", "segments")}, d["same_result"])
PY
for f in 2 3; do
  python bench.py --steps 5 --warmup 3 --cpu-sample-pairs 0 --program-pairs 0 --in-flight $f > gpurun_out/${TAG}_bench_f$f.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_f$f.json").read().strip().splitlines()[-1])
print("in flight $f: device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"], d["roofline"]["frac"])
PY
done
for h in 0 1; do
  KB_SEED_LD_HINT=$h ncu --set full --clock-control none -k regex:'k_fm_seed' -s 2 -c 1 -o /tmp/${TAG}seed$h -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}_ncu_seed$h.log 2>&1
  ncu -i /tmp/${TAG}seed$h.ncu-rep --page raw --csv > gpurun_out/${TAG}_seed_raw_hint$h.csv 2>/dev/null
done
ls -la gpurun_out/${TAG}_seed_raw_hint*.csv
for v in "KB_SEED_LD_HINT=0" "KB_SEED_LD_HINT=1"; do
  echo "== $v" >> gpurun_out/${TAG}_modes.jsonl
  env $v python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 0 --ref-pb 0 --check 100 >> gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err
done
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_modes.jsonl"):
    if ln.startswith("=="): print(ln.strip()); continue
    d = json.loads(ln); print(" ", d["mode"], round(d["device_ms"], 2), {k: round(v, 2) for k, v in d["stage_ms"].items()}, d.get("oracle_mismatches"))
PY
KART_B200_TRACE=1 python scripts/cli_compare.py --pairs 2500000 --prefix $PREFIX --error 0.01 --ours-only > gpurun_out/${TAG}_cli_c3.json 2> gpurun_out/${TAG}_cli_trace.txt; cat gpurun_out/${TAG}_cli_c3.json; tail -42 gpurun_out/${TAG}_cli_trace.txt
tail -5 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_ab.err
