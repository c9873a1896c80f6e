"""GPU experiment: e2e time of kb_map_chunk vs sub-batch size (KB_PIPE_SUB_READS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import parity_util as pu
from kart_b200 import KartIndex, Mapper, synth
from kart_b200.binding import ALN_DTYPE, PAIR_DTYPE
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
PREFIX = os.environ.get("SWEEP_PREFIX") or pu.default_prefix()      # SWEEP_PREFIX=data/_gen/syn/syn3100 SWEEP_ERR=0.01 SWEEP_PACKED=1 for the C3 workload
idx = KartIndex(PREFIX); g = pu.pac_genome(idx)
r1, r2, _ = synth.simulate(g, pairs, 150, float(os.environ.get("SWEEP_ERR", "0.02")), seed=1)
reads = pu.interleave(r1, r2); n = reads.shape[0]
seq_pin = torch.empty(reads.size, dtype=torch.uint8).pin_memory(); seq_pin.numpy()[:] = reads.reshape(-1)
off_pin = torch.empty(n + 1, dtype=torch.int64).pin_memory(); off_pin.numpy()[:] = np.arange(n + 1, dtype=np.int64) * 150
flat, off = seq_pin.numpy(), off_pin.numpy().view(np.uint64)
est = np.full(n // 2, 1500, dtype=np.int32)
aln_pin = torch.empty(n * ALN_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
pair_pin = torch.empty((n // 2) * PAIR_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
cig_pin = torch.empty(4 * n + 1024, dtype=torch.int32).pin_memory()
out = (aln_pin.numpy().view(ALN_DTYPE), pair_pin.numpy().view(PAIR_DTYPE), cig_pin.numpy().view(np.uint32))
plans = [dict(sub=2 * pairs + 2), dict(sub=500000), dict(sub=400000),
         dict(sub=500000, first=65536, grow=200, tail=0), dict(sub=500000, first=65536, grow=200, tail=65536),
         dict(sub=500000, first=131072, grow=200, tail=131072), dict(sub=400000, first=100000, grow=200, tail=100000),
         dict(sub=600000, first=65536, grow=300, tail=65536), dict(sub=700000, first=131072, grow=250, tail=131072),
         dict(sub=500000, first=32768, grow=200, tail=32768), dict(sub=333334, first=65536, grow=150, tail=65536)]
if len(sys.argv) > 2:
    plans = [dict(zip(("sub", "first", "grow", "tail"), (int(x) for x in a.split(",")))) for a in sys.argv[2:]]
for pl in plans:
    sub = pl["sub"]
    os.environ["KB_PIPE_SUB_READS"] = str(sub)
    os.environ["KB_PIPE_FIRST"] = str(pl.get("first", -1)); os.environ["KB_PIPE_GROW"] = str(pl.get("grow", 200)); os.environ["KB_PIPE_TAIL"] = str(pl.get("tail", 0))
    os.environ["KB_PIPE_MIN_READS"] = "1000" if sub != 2 * pairs + 2 else "2000000000"
    m = Mapper(device=0); m.upload_index(idx, expand_sa=True); m.set_params(paired=True)
    pk = m.pack(flat, off, threads=16) if os.environ.get("SWEEP_PACKED") else None
    for _ in range(2): m.map_chunk(flat, off, est, out=out, packed=pk)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): m.map_chunk(flat, off, est, out=out, packed=pk)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print("plan %-60s e2e %.2f ms  (%.1f M reads/s)  launches %d" % (pl, dt * 1e3, n / dt / 1e6, m.work()["launches"]), flush=True)
    del m
