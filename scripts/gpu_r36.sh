#!/bin/bash
# r36: the evidence run of the final build: full -m gpu suite, both bench arms as the driver runs them, ncu launch list and full capture at C3 (traffic
# JSON with the hash of the CUDA sources, summaries and hot lines as text), C4 / C5 shapes with the reference timed next to them, whole program at 16 M reads.
TAG=${1:-r36}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt; nproc >> gpurun_out/${TAG}_gpu.txt; free -g | head -2 >> gpurun_out/${TAG}_gpu.txt
( time python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cut -c1-1500 gpurun_out/${TAG}_pytest.txt | tail -10
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench_ref.json | tail -2
PREFIX=data/_gen/syn/syn3100
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}c3_launches.csv python bench.py --steps 2 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3_ncu_bench.log 2>&1
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_fm_seed|k_sa_locate|k_rescue_fast|k_cand_pair|k_cand_heavy|k_segments$|k_align_part|k_assemble$|k_finalize|k_unpack$|k_nw_tile' -s 16 -c 16 -o gpurun_out/${TAG}c3_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3_ncu_full.log 2>&1
tail -n 2 gpurun_out/${TAG}c3_ncu_full.log | cut -c1-200
python scripts/ncu_summary.py ${TAG}c3 500000 syn3100 | tail -n 2
for k in k_fm_seed k_segments$ k_rescue_fast k_cand_heavy_finish k_assemble$ k_finalize; do python scripts/ncu_hot_lines.py gpurun_out/${TAG}c3_prof.ncu-rep "$k" 30 > gpurun_out/${TAG}c3_$(echo $k | tr -d '$')_hot_lines.txt 2>&1; done
rm -f gpurun_out/${TAG}c3_prof.ncu-rep
cp profiles/${TAG}c3_* gpurun_out/ 2>/dev/null
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"])
print(d["cpu_baseline"]); print(d.get("e2e_program")); print(d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["traffic_source"])
PY
grep real gpurun_out/${TAG}_bench.err
python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 50000 --ref-pb 500 --check 200 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err; cut -c1-700 gpurun_out/${TAG}_modes.jsonl
python scripts/cli_compare.py --pairs 8000000 --prefix $PREFIX --error 0.01 --no-compare > gpurun_out/${TAG}_cli_c3_16m.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_cli_c3_16m.json
du -sh gpurun_out; tail -n 5 gpurun_out/${TAG}_bench.err
