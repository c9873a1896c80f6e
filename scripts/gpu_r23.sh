#!/bin/bash
# r23: full -m gpu suite (k_finalize through shared memory, rescue plan/commit grids), bench at its default, C4 / C5 shapes on the 3.1 Gbp
# index with ncu captures of their align / candidate kernels, ncu of the rescue kernels at C3.
TAG=${1:-r23}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.txt 2>&1; cut -c1-1500 gpurun_out/${TAG}_pytest.txt | tail -12
PREFIX=data/_gen/syn/syn3100
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)  e2e_sync %.3f  e2e_text %.3f" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e_sync"]["ms_per_step"], d["e2e_text"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"])
print(d["cpu_baseline"]); print(d.get("e2e_program"))
PY
python scripts/gpu_modes.py --prefixes $PREFIX --se 200000 --pb 50000 --ref-se 0 --ref-pb 0 --check 100 > gpurun_out/${TAG}_modes.jsonl 2>> gpurun_out/${TAG}_bench.err; cut -c1-700 gpurun_out/${TAG}_modes.jsonl
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_align_part|k_nw_tile|k_nw_warp|k_segments$|k_assemble$|k_cand_pair|k_align_gather' -s 13 -c 13 -o gpurun_out/${TAG}c4_prof -f python scripts/gpu_modes.py --prefixes $PREFIX --modes se100 --se 200000 --pb 0 --ref-se 0 --check 0 --reps 1 > gpurun_out/${TAG}c4_ncu.log 2>&1; tail -2 gpurun_out/${TAG}c4_ncu.log | cut -c1-200
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_align_part|k_nw_tile|k_nw_warp|k_cand_pacbio|k_fm_seed|k_assemble_slow|k_segments_slow|k_align_gather' -s 14 -c 14 -o gpurun_out/${TAG}c5_prof -f python scripts/gpu_modes.py --prefixes $PREFIX --modes pacbio --se 0 --pb 20000 --ref-pb 0 --check 0 --reps 1 > gpurun_out/${TAG}c5_ncu.log 2>&1; tail -2 gpurun_out/${TAG}c5_ncu.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:'k_rescue|k_finalize|k_unpack$|k_cand_heavy' -s 7 -c 7 -o gpurun_out/${TAG}c3r_prof -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3r_ncu.log 2>&1; tail -2 gpurun_out/${TAG}c3r_ncu.log | cut -c1-200
ls -la gpurun_out | grep ${TAG}
tail -5 gpurun_out/${TAG}_bench.err
