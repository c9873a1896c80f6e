"""GPU experiment: e2e time of kb_map_chunk vs sub-batch size (KB_PIPE_SUB_READS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import parity_util as pu
from kart_b200 import KartIndex, Mapper, synth
from kart_b200.binding import ALN_DTYPE, PAIR_DTYPE
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
idx = KartIndex(pu.default_prefix()); g = pu.genome_of(idx)
r1, r2, _ = synth.simulate(g, pairs, 150, 0.02, seed=1)
reads = pu.interleave(r1, r2); n = reads.shape[0]
seq_pin = torch.empty(reads.size, dtype=torch.uint8).pin_memory(); seq_pin.numpy()[:] = reads.reshape(-1)
off_pin = torch.empty(n + 1, dtype=torch.int64).pin_memory(); off_pin.numpy()[:] = np.arange(n + 1, dtype=np.int64) * 150
flat, off = seq_pin.numpy(), off_pin.numpy().view(np.uint64)
est = np.full(n // 2, 1500, dtype=np.int32)
aln_pin = torch.empty(n * ALN_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
pair_pin = torch.empty((n // 2) * PAIR_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
cig_pin = torch.empty(4 * n + 1024, dtype=torch.int32).pin_memory()
out = (aln_pin.numpy().view(ALN_DTYPE), pair_pin.numpy().view(PAIR_DTYPE), cig_pin.numpy().view(np.uint32))
for sub in [0, 2 * pairs + 2, 1000000, 500000, 333334, 250000, 200000, 131072, 65536]:
    if sub: os.environ["KB_PIPE_SUB_READS"] = str(sub)
    os.environ["KB_PIPE_MIN_READS"] = "1000" if sub != 2 * pairs + 2 else "2000000000"
    m = Mapper(device=0); m.upload_index(idx, expand_sa=True); m.set_params(paired=True)
    for _ in range(2): m.map_chunk(flat, off, est, out=out)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4): m.map_chunk(flat, off, est, out=out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 4
    sm = m.stage_ms()
    print("sub %8d  e2e %.2f ms  (%.1f M reads/s)  kernels total %.2f ms  stages %s launches %d" % (sub, dt * 1e3, n / dt / 1e6, sm["total"], {k: round(v, 2) for k, v in sm.items() if k != "total"}, m.work()["launches"]), flush=True)
    del m
