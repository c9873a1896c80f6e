#!/bin/bash
# r31: k_align_part / k_fm_seed_q on the C5 shape after the packed partition and the adaptive quorums: ncu --set full with source, text back.
TAG=${1:-r31}
mkdir -p gpurun_out
PREFIX=data/_gen/syn/syn3100
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import parity_util as pu
print(pu.ensure_syn_index(3100, 24, 12345))
PY
hot() { python scripts/ncu_hot_lines.py $1 "$2" 70 > gpurun_out/${TAG}_$3_hot.txt 2>&1; }
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_align_part|k_fm_seed|k_assemble_slow' -s 3 -c 3 -o /tmp/${TAG}c5 -f python scripts/gpu_modes.py --prefixes $PREFIX --modes pacbio --se 0 --pb 50000 --ref-pb 0 --check 0 --reps 1 > gpurun_out/${TAG}c5_ncu.log 2>&1
ncu -i /tmp/${TAG}c5.ncu-rep --page raw --csv > gpurun_out/${TAG}c5_raw.csv 2>/dev/null
hot /tmp/${TAG}c5.ncu-rep k_align_part c5_align_part; hot /tmp/${TAG}c5.ncu-rep 'k_fm_seed' c5_fm_seed; hot /tmp/${TAG}c5.ncu-rep 'k_assemble_slow' c5_assemble_slow
ls -la gpurun_out
