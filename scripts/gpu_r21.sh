#!/bin/bash
# r21: full -m gpu suite (index built by the GPU builder now), bench at its default with the chunks-in-flight e2e and k_cand_heavy,
# whole program against kart -t 1 on 3 M reads at C3 (three batches: the EstDistance recurrence across batches), ncu at C3.
TAG=${1:-r21}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cat gpurun_out/${TAG}_pytest.txt | cut -c1-1500
ls -la data/_gen/syn/ | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r21_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, "pack", round(d["e2e"]["host_pack_ms_outside_timed_region"], 1), d["e2e"]["records_equal_text_entry"])
print(d["cpu_baseline"]); print(d.get("e2e_program"))
PY
tail -5 gpurun_out/${TAG}_bench.err
PREFIX=data/_gen/syn/syn3100
python scripts/cli_compare.py --pairs 1500000 --prefix $PREFIX --error 0.01 --t1 --diff-out gpurun_out/${TAG}_cli_diff.txt > gpurun_out/${TAG}_cli_c3.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_cli_c3.json; head -c 1500 gpurun_out/${TAG}_cli_diff.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}c3_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_rescue_fast|k_cand_pair|k_cand_heavy|k_segments$|k_align_part|k_assemble$|k_finalize|k_unpack$' -s 9 -c 9 -o gpurun_out/${TAG}c3_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}c3_ncu_full.log | cut -c1-200
python scripts/ncu_summary.py ${TAG}c3 500000 syn3100 | tail -2
cp profiles/${TAG}c3_* gpurun_out/ 2>/dev/null
tail -12 gpurun_out/${TAG}_bench.err
