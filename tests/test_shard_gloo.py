"""CPU: the N>1 path. Reads shard in contiguous chunk-aligned ranges, every rank maps its own shard against
its own copy of the index (no data-path collective), rank 0 gathers the records. Two `gloo` ranks on 127.0.0.1, each driving the
host-emulation build of the kernels (test infrastructure; the GPU version of this is `bench.py --gpus N`), must reproduce the
single-process result record for record."""
import os
import socket
import sys

import numpy as np
import pytest

import parity_util as pu
from kart_b200 import KartIndex, Mapper, synth
from kart_b200.shard import gather_results, shard_ranges, slice_reads

CHUNK = 400   # small chunk so that 2 ranks get several chunks each out of a test-sized read set


def test_shard_ranges_properties():
    for n, world, chunk in [(0, 4, 4000), (1, 2, 4000), (3999, 2, 4000), (4000, 2, 4000), (4002, 2, 4000), (20000, 8, 4000), (100_000, 3, 4000), (2_000_000, 8, 4000), (33, 4, 10)]:
        rs = shard_ranges(n, world, chunk)
        assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == n
        for (a, b), (c, d) in zip(rs, rs[1:]):
            assert b == c and a <= b
        for lo, hi in rs:
            assert lo % chunk == 0 or lo == n             # shards start on chunk boundaries: EstDistance chunks never straddle ranks
            assert lo % 2 == 0 or lo == n                 # pairs are never split
        sizes = [(hi - lo + chunk - 1) // chunk for lo, hi in rs]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_ranges(10, 0)


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = KartIndex(pu.MINI_PREFIX)
        g = pu.genome_of(idx)
        r1, r2, _ = synth.simulate(g, 1100, 150, 0.03, seed=77, indel=0.003)   # every rank derives the same stream, then keeps its shard
        reads = pu.interleave(r1, r2)
        flat, off = Mapper.pack_reads(reads)
        lo, hi = shard_ranges(len(reads), world, CHUNK)[rank]
        f, o = slice_reads(flat, off, lo, hi)
        m = pu.make_mapper(idx, emul=True, paired=True)
        aln, pairs, cig = m.map_chunk(f, o, 1500)
        got = gather_results(aln, cig, pairs[:(hi - lo) // 2], rank, world)
        if rank == 0:
            a, c, p = got
            np.savez(out_path, aln=a, cig=c, pairs=p)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(built, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    r1, r2, _ = synth.simulate(g, 1100, 150, 0.03, seed=77, indel=0.003)
    reads = pu.interleave(r1, r2)
    flat, off = Mapper.pack_reads(reads)
    m = pu.make_mapper(idx, emul=True, paired=True)
    aln, pairs, cig = m.map_chunk(flat, off, 1500)
    a = got["aln"]
    assert len(a) == len(aln) == 2200
    for f in aln.dtype.names:
        if f != "cig_off":
            assert np.array_equal(a[f], aln[f]), f
    for r in range(len(aln)):   # cigar offsets differ (per-rank arenas), the ops must not
        x = got["cig"][a["cig_off"][r]:a["cig_off"][r] + a["cig_len"][r]]
        y = cig[aln["cig_off"][r]:aln["cig_off"][r] + aln["cig_len"][r]]
        assert np.array_equal(x, y), r
    assert np.array_equal(got["pairs"], pairs[:1100])
    assert int((aln["score"] > 0).sum()) > 2000
