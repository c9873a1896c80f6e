"""GPU experiment harness: one C2-shaped workload, several runtime configurations (environment variables read by kb_init),
device-resident stage times and e2e (pinned host buffers) for each. One JSON line per configuration.
Usage: python scripts/gpu_ab.py [--pairs N] [--prefix P] [--error E] [--lib path.so] CONFIG [CONFIG ...]
CONFIG = name:VAR=val,VAR=val (empty list = defaults), e.g.  base:  nostreams:KB_NW_STREAMS=0  w16:KB_ALIGN_WARPS=2368"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import parity_util as pu
from kart_b200 import KartIndex, Mapper, synth
from kart_b200.binding import ALN_DTYPE, PAIR_DTYPE

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1000000)
ap.add_argument("--prefix", default=None)
ap.add_argument("--error", type=float, default=0.02)
ap.add_argument("--lib", default=None)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--no-e2e", action="store_true")
ap.add_argument("configs", nargs="+")
a = ap.parse_args()
prefix = a.prefix or pu.default_prefix()
idx = KartIndex(prefix); g = pu.pac_genome(idx) if idx.l_pac > 500_000_000 else pu.genome_of(idx)
r1, r2, _ = synth.simulate(g, a.pairs, 150, a.error, seed=1)
reads = pu.interleave(r1, r2); n = reads.shape[0]
seq_pin = torch.empty(reads.size, dtype=torch.uint8).pin_memory(); seq_pin.numpy()[:] = reads.reshape(-1)
off_pin = torch.empty(n + 1, dtype=torch.int64).pin_memory(); off_pin.numpy()[:] = np.arange(n + 1, dtype=np.int64) * 150
flat, off = seq_pin.numpy(), off_pin.numpy().view(np.uint64)
est = np.full(n // 2, 1500, dtype=np.int32)
aln_pin = torch.empty(n * ALN_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
pair_pin = torch.empty((n // 2) * PAIR_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
cig_pin = torch.empty(4 * n + 1024, dtype=torch.int32).pin_memory()
out = (aln_pin.numpy().view(ALN_DTYPE), pair_pin.numpy().view(PAIR_DTYPE), cig_pin.numpy().view(np.uint32))
ref_sig = None
for spec in a.configs:
    name, _, kv = spec.partition(":")
    env = dict(x.split("=", 1) for x in kv.split(",") if x)
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    m = Mapper(device=0, lib_path=a.lib); m.upload_index(idx, expand_sa=True); m.set_params(paired=True)
    m.stage(flat, off, est)
    for _ in range(2): m.run()
    st = {}
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(a.reps):
        m.run()
        for k, v in m.stage_ms().items(): st[k] = st.get(k, 0.0) + v / a.reps
    torch.cuda.synchronize(); dev = (time.perf_counter() - t) / a.reps
    for _ in range(1 if a.no_e2e else 2): aln, pr, cig = m.map_chunk(flat, off, est, out=out)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(0 if a.no_e2e else a.reps): aln, pr, cig = m.map_chunk(flat, off, est, out=out)
    torch.cuda.synchronize(); e2e = (time.perf_counter() - t) / a.reps
    sig = (int(aln["pos"].sum()), int(aln["score"].sum()), int(aln["flag"].sum()), int(aln["cig_len"].sum()), int(aln["mapq"].sum()))
    if ref_sig is None: ref_sig = sig
    print(json.dumps({"config": name, "env": env, "reads": n, "device_ms": round(dev * 1e3, 3), "device_mreads_s": round(n / dev / 1e6, 1), "e2e_ms": round(e2e * 1e3, 3),
                      "e2e_mreads_s": round(n / e2e / 1e6, 1), "stage_ms": {k: round(v, 3) for k, v in st.items()}, "launches": m.work()["launches"], "same_result": sig == ref_sig}), flush=True)
    m.close(); del m
    for k, v in saved.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
