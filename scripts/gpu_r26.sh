#!/bin/bash
# r26: k_align_part on the C5 shape at full size (50 k x 7 kbp) and on C4 after r25's changes: ncu --set full with source, text pages back.
TAG=${1:-r26}
mkdir -p gpurun_out
PREFIX=data/_gen/syn/syn3100
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import parity_util as pu
print(pu.ensure_syn_index(3100, 24, 12345))
PY
hot() { python scripts/ncu_hot_lines.py $1 "$2" 60 > gpurun_out/${TAG}_$3_hot.txt 2>&1; }
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_align_part|k_cand_pacbio' -s 3 -c 3 -o /tmp/${TAG}c5 -f python scripts/gpu_modes.py --prefixes $PREFIX --modes pacbio --se 0 --pb 50000 --ref-pb 0 --check 0 --reps 1 > gpurun_out/${TAG}c5_ncu.log 2>&1
ncu -i /tmp/${TAG}c5.ncu-rep --page raw --csv > gpurun_out/${TAG}c5_raw.csv 2>/dev/null
hot /tmp/${TAG}c5.ncu-rep k_align_part c5_align_part; hot /tmp/${TAG}c5.ncu-rep 'k_cand_pacbio$' c5_cand_pacbio; hot /tmp/${TAG}c5.ncu-rep 'k_cand_pacbio_sort' c5_cand_pacbio_sort
KB_NW_STREAMS=0 ncu --set full --clock-control none --import-source on -k regex:'k_align_part' -s 1 -c 1 -o /tmp/${TAG}c4 -f python scripts/gpu_modes.py --prefixes $PREFIX --modes se100 --se 200000 --pb 0 --ref-se 0 --check 0 --reps 1 > gpurun_out/${TAG}c4_ncu.log 2>&1
ncu -i /tmp/${TAG}c4.ncu-rep --page raw --csv > gpurun_out/${TAG}c4_raw.csv 2>/dev/null
hot /tmp/${TAG}c4.ncu-rep k_align_part c4_align_part
python - <<'PY'
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, parity_util as pu
from kart_b200 import KartIndex, Mapper, synth
idx = KartIndex("data/_gen/syn/syn3100"); g = pu.pac_genome(idx)
m = Mapper(); m.upload_index(idx, expand_sa=True); m.set_params(pacbio=True, paired=False)
r, _, _ = synth.simulate(g, 50000, 7000, 0.15, seed=3, paired=False, indel=0.01)
flat, off = Mapper.pack_reads(r); m.stage(flat, off); m.run(); m.run()
c = m.debug(9, np.uint32, 32)
print("counters", [int(x) for x in c])
jobs = m.debug(10, np.dtype([("gpos", "<i8"), ("read", "<u4"), ("rpos", "<i4"), ("rlen", "<i4"), ("glen", "<i4"), ("run_off", "<u4"), ("nruns", "<i4"), ("ident", "<i4"), ("aligned", "<i4")]), int(c[11]))
mx = np.maximum(jobs["rlen"], jobs["glen"])
print("jobs", len(jobs), "partition jobs", int(c[23]), "pieces", int(c[24]))
print("job size percentiles (max side)", [int(np.percentile(mx, p)) for p in (50, 90, 99, 99.9, 100)])
big = mx > 30
print("sum rl*gl over jobs with both > 30:", int((jobs["rlen"].astype(np.int64) * jobs["glen"])[(jobs["rlen"] > 30) & (jobs["glen"] > 30)].sum()))
print("hist of max side for partition jobs", np.histogram(mx[(jobs["rlen"] > 30) & (jobs["glen"] > 30)], bins=[30, 50, 100, 200, 300, 500, 1000, 2000, 4000, 100000])[0].tolist())
PY
ls -la gpurun_out | head -20
