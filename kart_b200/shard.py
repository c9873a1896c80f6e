"""Read sharding over the GPUs of one box: contiguous, chunk-aligned ranges of the read stream, one per rank;
the FM-index is replicated; there is NO collective on the data path. `torch.distributed` is only used to bring the per-rank
results (kb_aln_t records + cigar ops) back to rank 0, which is where the reference's single writer sits (Mapping.cpp:597-599).

Chunk alignment matters for parity: the reference derives EstDistance per 4000-read chunk from the chunks before it
(Mapping.cpp:533-540), so a shard boundary inside a chunk would change which pairs see which EstDistance. Shards therefore
start on chunk boundaries and pairs are never split."""
from __future__ import annotations

import numpy as np

CHUNK_READS = 4000   # ReadChunkSize (reference src/GetData.cpp:140); 10 in -pacbio mode (:216)


def shard_ranges(n_reads: int, world: int, chunk: int = CHUNK_READS):
    """[(lo, hi)) read ranges, one per rank: contiguous, starting on chunk boundaries, sizes differing by at most one chunk.
    A trailing partial chunk goes to the last non-empty rank; ranks beyond the number of chunks get empty ranges."""
    if world <= 0 or chunk <= 0 or n_reads < 0:
        raise ValueError("shard_ranges: bad arguments")
    n_chunks = (n_reads + chunk - 1) // chunk
    base, extra = divmod(n_chunks, world)
    out, c = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        lo, hi = min(c * chunk, n_reads), min((c + k) * chunk, n_reads)
        out.append((lo, hi))
        c += k
    return out


def slice_reads(flat: np.ndarray, off: np.ndarray, lo: int, hi: int):
    """The (flat, off) pair of reads [lo, hi) with offsets rebased to zero."""
    o = np.asarray(off).astype(np.int64)
    b, e = int(o[lo]), int(o[hi])
    return np.ascontiguousarray(flat[b:e]), np.ascontiguousarray((o[lo:hi + 1] - b).astype(np.uint64))


def gather_results(aln: np.ndarray, cig: np.ndarray, pairs, rank: int, world: int, group=None):
    """Brings every rank's (aln, cigar, pair stats) to rank 0 in rank order and rebases kb_aln_t.cig_off into the concatenated
    cigar array. Returns (aln, cigar, pairs) on rank 0 and None elsewhere. Host-side only (results already left the GPU)."""
    import torch
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object((aln.tobytes(), cig.tobytes(), None if pairs is None else pairs.tobytes()), parts, dst=0, group=group)
    if rank != 0:
        return None
    alns, cigs, prs, base = [], [], [], 0
    for a, c, p in parts:
        a = np.frombuffer(a, dtype=aln.dtype).copy()
        c = np.frombuffer(c, dtype=np.uint32)
        a["cig_off"] += np.uint32(base)
        base += len(c)
        alns.append(a)
        cigs.append(c)
        if p is not None:
            prs.append(np.frombuffer(p, dtype=pairs.dtype))
    return np.concatenate(alns), np.concatenate(cigs), (np.concatenate(prs) if prs else None)
