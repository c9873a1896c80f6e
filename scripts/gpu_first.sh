set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import sys, time; sys.path.insert(0,'tests')
import parity_util as pu, numpy as np
from kart_b200 import KartIndex, Mapper, synth
idx = KartIndex(pu.default_prefix()); g = pu.genome_of(idx)
for expand in (False, True):
    m = pu.make_mapper(idx, expand_sa=expand, paired=True)
    r1,r2,_ = synth.simulate(g, 500000, 150, 0.02, seed=1)
    flat, off = Mapper.pack_reads(pu.interleave(r1,r2))
    for it in range(3):
        t=time.time(); aln,pairs,cig = m.map_chunk(flat, off, 1500); dt=time.time()-t
        print('expand',expand,'iter',it,'e2e %.3fs'%dt, (len(off)-1)/dt, m.stage_ms(), m.work(), flush=True)
PY
