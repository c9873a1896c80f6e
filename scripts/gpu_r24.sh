#!/bin/bash
# r24: where k_align_part / k_cand_pacbio spend their time on the C4 / C5 shapes and k_rescue_fast at C3 (ncu --set full with source; the
# reports stay on the box, their raw pages and hot lines come back as text), and the bench after k_finalize went back to direct stores.
TAG=${1:-r24}
mkdir -p gpurun_out
PREFIX=data/_gen/syn/syn3100
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import parity_util as pu
print(pu.ensure_syn_index(3100, 24, 12345))
PY
hot() { python scripts/ncu_hot_lines.py $1 "$2" 45 > gpurun_out/${TAG}_$3_hot.txt 2>&1; }
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_align_part|k_nw_tile|k_segments$|k_assemble$|k_align_gather' -s 9 -c 9 -o /tmp/${TAG}c4 -f python scripts/gpu_modes.py --prefixes $PREFIX --modes se100 --se 200000 --pb 0 --ref-se 0 --check 0 --reps 1 > gpurun_out/${TAG}c4_ncu.log 2>&1
ncu -i /tmp/${TAG}c4.ncu-rep --page raw --csv > gpurun_out/${TAG}c4_raw.csv 2>/dev/null
hot /tmp/${TAG}c4.ncu-rep k_align_part c4_align_part; hot /tmp/${TAG}c4.ncu-rep 'k_segments$' c4_segments; hot /tmp/${TAG}c4.ncu-rep 'k_assemble$' c4_assemble
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_align_part|k_nw_tile|k_nw_warp|k_cand_pacbio|k_fm_seed|k_assemble_slow|k_segments_slow|k_align_gather' -s 12 -c 12 -o /tmp/${TAG}c5 -f python scripts/gpu_modes.py --prefixes $PREFIX --modes pacbio --se 0 --pb 10000 --ref-pb 0 --check 0 --reps 1 > gpurun_out/${TAG}c5_ncu.log 2>&1
ncu -i /tmp/${TAG}c5.ncu-rep --page raw --csv > gpurun_out/${TAG}c5_raw.csv 2>/dev/null
hot /tmp/${TAG}c5.ncu-rep k_align_part c5_align_part; hot /tmp/${TAG}c5.ncu-rep k_cand_pacbio c5_cand_pacbio; hot /tmp/${TAG}c5.ncu-rep k_assemble_slow c5_assemble_slow; hot /tmp/${TAG}c5.ncu-rep 'k_fm_seed' c5_fm_seed
ncu --set full --clock-control none --import-source on -k regex:'k_rescue_fast|k_rescue_plan|k_rescue_commit' -s 3 -c 3 -o /tmp/${TAG}c3r -f python bench.py --steps 1 --warmup 1 --pairs 500000 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}c3r_ncu.log 2>&1
ncu -i /tmp/${TAG}c3r.ncu-rep --page raw --csv > gpurun_out/${TAG}c3r_raw.csv 2>/dev/null
hot /tmp/${TAG}c3r.ncu-rep k_rescue_fast c3_rescue_fast; hot /tmp/${TAG}c3r.ncu-rep k_rescue_plan c3_rescue_plan
python bench.py --steps 3 --warmup 3 --cpu-sample-pairs 0 --program-pairs 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("device %.3f ms (%.1f M/s)  e2e %.3f ms (%.1f M/s)" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6))
print({k: round(v, 3) for k, v in d["stage_ms"].items()}, d["e2e"]["records_equal_text_entry"])
PY
ls -la gpurun_out; du -sh gpurun_out
