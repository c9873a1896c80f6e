"""Seeded synthetic inputs for the parity tests and bench.py.

The reference's own simulator (wgsim/wgsim.c) cannot be seeded (getopt string at wgsim.c:430 has no
seed option, seed = time(0) at :447), so the workloads BASELINE.json names are regenerated here from
fixed numpy seeds with the same error model: uniform fragment position, insert size ~ N(500, 50),
random strand flip, substitution-only sequencing errors `c -> (c+1)&3` at rate `err` (wgsim.c:366-369),
plus per-read haplotype variation (SNPs and short indels; wgsim -r 0.001 -R 0.15 -X 0.3).
FASTQ records are fixed width so that 10^6 pairs are written with pure numpy.
"""
from __future__ import annotations

import os
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCATGCA"):
    _COMP[a] = b


def revcomp_bytes(a: np.ndarray) -> np.ndarray:
    """Reverse-complement rows of a 2-D uint8 array (or a 1-D array)."""
    return _COMP[a[..., ::-1]]


def read_fasta(path: str):
    """Returns (names, list of uint8 arrays with upper-case bases)."""
    names, seqs, cur = [], [], []
    with open(path, "rb") as fh:
        for ln in fh:
            if ln.startswith(b">"):
                if cur:
                    seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
                    cur = []
                names.append(ln[1:].split()[0].decode())
            else:
                cur.append(ln.strip().upper())
    if cur:
        seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
    return names, seqs


def write_fasta(path: str, names, seqs, width: int = 70) -> None:
    with open(path, "wb") as fh:
        for n, s in zip(names, seqs):
            fh.write(b">" + n.encode() + b"\n")
            b = s.tobytes()
            fh.write(b"\n".join(b[i:i + width] for i in range(0, len(b), width)) + b"\n")


def make_genome_codes(total_bp: int, n_contigs: int, seed: int, repeats=((3000, 20, 0.02), (300, 200, 0.05))):
    """i.i.d. uniform contigs with injected repeat families (len, copies, per-copy divergence) as 2-bit codes (A 0, C 1, G 2, T 3):
    (names, contig lengths, codes of the concatenated genome). Mirrors the pilot genome of SURVEY.md Appendix B at any scale."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=total_bp, dtype=np.uint8)
    for rep_len, copies, div in repeats:
        if rep_len * 2 >= total_bp:
            continue
        elem = rng.integers(0, 4, size=rep_len, dtype=np.uint8)
        for p in rng.integers(0, total_bp - rep_len, size=copies):
            c = elem.copy()
            m = rng.random(rep_len) < div
            c[m] = rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)
            g[p:p + rep_len] = c
    cuts = np.linspace(0, total_bp, n_contigs + 1).astype(np.int64)
    # uneven contigs so that "longer chromosome wins" ties are exercised
    if n_contigs > 1:
        jitter = rng.integers(-total_bp // (8 * n_contigs), total_bp // (8 * n_contigs) + 1, size=n_contigs - 1)
        cuts[1:-1] += jitter
    names = ["chr%d" % (i + 1) for i in range(n_contigs)]
    return names, np.diff(cuts), g


def make_genome(total_bp: int, n_contigs: int, seed: int, repeats=((3000, 20, 0.02), (300, 200, 0.05))):
    """The same genome as upper-case characters: (names, list of contigs)."""
    names, lens, g = make_genome_codes(total_bp, n_contigs, seed, repeats)
    g = _ACGT[g]
    cuts = np.concatenate([[0], np.cumsum(lens)])
    return names, [g[cuts[i]:cuts[i + 1]] for i in range(n_contigs)]


def write_pac_ann(prefix: str, names, lens, codes: np.ndarray) -> None:
    """prefix.pac / .ann / .amb exactly as the index builders write them for an ACGT-only FASTA without header comments
    (BWT_Index/bntseq.c:59-89,192-205): the input of `kart index -gpu -pac prefix`, without a FASTA detour."""
    L = int(len(codes))
    pad = (-L) % 4
    c = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]) if pad else codes
    pac = (c[0::4] << 6) | (c[1::4] << 4) | (c[2::4] << 2) | c[3::4]
    with open(prefix + ".pac", "wb") as fh:
        pac.astype(np.uint8).tofile(fh)
        if L % 4 == 0:
            fh.write(b"\x00")
        fh.write(bytes([L % 4]))
    with open(prefix + ".ann", "w") as fh:
        fh.write("%d %d %u\n" % (L, len(names), 11))
        off = 0
        for n, ln in zip(names, lens):
            fh.write("0 %s (null)\n%d %d 0\n" % (n, off, int(ln)))
            off += int(ln)
    with open(prefix + ".amb", "w") as fh:
        fh.write("%d %d 0\n" % (L, len(names)))


class PacText:
    """The forward strand of a BWA .pac file (2 bit / base, MSB first) addressed like a uint8 array of upper-case bases, decoded
    on access: read simulation on a multi-Gbp genome then touches only the windows it samples instead of a decoded copy."""

    def __init__(self, pac: np.ndarray, l_pac: int):
        self.pac, self.l_pac = pac, int(l_pac)

    def __len__(self):
        return self.l_pac

    def __getitem__(self, key):
        if isinstance(key, slice):
            idx = np.arange(*key.indices(self.l_pac), dtype=np.int64)
        else:
            idx = np.asarray(key, dtype=np.int64)
        return _ACGT[(self.pac[idx >> 2] >> ((~idx & 3) << 1).astype(np.uint8)) & 3]


class PacGenome:
    """(contig lengths, PacText): accepted by simulate() in place of the list of decoded contigs"""

    def __init__(self, pac: np.ndarray, l_pac: int, lens):
        self.text, self.lens = PacText(pac, l_pac), np.asarray(lens, dtype=np.int64)


def _apply_errors(reads: np.ndarray, rng, err: float, snp: float) -> None:
    code = np.zeros(256, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    if err > 0:
        m = rng.random(reads.shape) < err
        reads[m] = _ACGT[(code[reads[m]] + 1) & 3]
    if snp > 0:
        m = rng.random(reads.shape) < snp
        reads[m] = _ACGT[(code[reads[m]] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3]


def simulate(genome_seqs, n: int, read_len: int, err: float, seed: int, paired: bool = True,
             insert_mean: float = 500.0, insert_sd: float = 50.0, snp: float = 0.00085,
             indel: float = 0.00015, indel_ext: float = 0.3, n_rate: float = 0.0):
    """Returns (r1, r2, pos) : uint8 arrays [n, read_len] in sequencer orientation (r2 None if not paired)
    and the 0-based forward-strand start of each fragment in the concatenated genome."""
    rng = np.random.default_rng(seed)
    if isinstance(genome_seqs, PacGenome):
        lens, cat = genome_seqs.lens, genome_seqs.text
    else:
        lens = np.array([len(s) for s in genome_seqs], dtype=np.int64)
        cat = np.concatenate(genome_seqs)
    starts = np.concatenate([[0], np.cumsum(lens)])
    frag = np.maximum(read_len, np.rint(rng.normal(insert_mean, insert_sd, size=n)).astype(np.int64)) if paired \
        else np.full(n, read_len, dtype=np.int64)
    slack = 64  # room for deletions
    # choose contig proportional to length, then a start such that the fragment (+slack) stays inside it
    w = np.maximum(lens - frag.max() - slack, 1).astype(np.float64)
    ctg = rng.choice(len(lens), size=n, p=w / w.sum())
    off = (rng.random(n) * np.maximum(lens[ctg] - frag - slack, 1)).astype(np.int64)
    pos = starts[ctg] + off
    idx = np.arange(read_len, dtype=np.int64)[None, :]

    def windows(start):   # [n, read_len] bases, gathered in slices so that the index temporaries stay small
        out = np.empty((n, read_len), dtype=np.uint8)
        step = max(1, (1 << 24) // max(read_len, 1))
        for a in range(0, n, step):
            out[a:a + step] = cat[start[a:a + step, None] + idx]
        return out
    left = windows(pos)
    right = None
    if paired:
        right = windows(pos + frag - read_len)
    # short indels, applied per read (python loop over the few affected reads)
    if indel > 0:
        for arr, base in ((left, pos), (right, (pos + frag - read_len) if paired else None)):
            if arr is None:
                continue
            hit = np.nonzero(rng.random(n) < indel * read_len)[0]
            for i in hit:
                k = int(rng.integers(5, read_len - 5))
                ln = int(rng.geometric(1.0 - indel_ext))
                ln = min(ln, 30, read_len - k - 1)
                src = cat[base[i]:base[i] + read_len + ln + 1]
                if rng.random() < 0.5:   # deletion from the reference
                    new = np.concatenate([src[:k], src[k + ln:k + ln + read_len - k]])
                else:                    # insertion into the read
                    ins = _ACGT[rng.integers(0, 4, size=ln, dtype=np.uint8)]
                    new = np.concatenate([src[:k], ins, src[k:]])[:read_len]
                if len(new) == read_len:
                    arr[i] = new
    _apply_errors(left, rng, err, snp)
    if paired:
        _apply_errors(right, rng, err, snp)
    if n_rate > 0:
        for arr in (left, right):
            if arr is not None:
                arr[rng.random(arr.shape) < n_rate] = ord("N")
    flip = rng.random(n) < 0.5
    if paired:
        r1 = np.where(flip[:, None], revcomp_bytes(right), left)
        r2 = np.where(flip[:, None], left, revcomp_bytes(right))
        return np.ascontiguousarray(r1), np.ascontiguousarray(r2), pos
    r1 = np.where(flip[:, None], revcomp_bytes(left), left)
    return np.ascontiguousarray(r1), None, pos


def _digits(vals: np.ndarray, width: int) -> np.ndarray:
    out = np.empty((len(vals), width), dtype=np.uint8)
    v = vals.astype(np.int64).copy()
    for k in range(width - 1, -1, -1):
        out[:, k] = 48 + v % 10
        v //= 10
    return out


def fastq_bytes(reads: np.ndarray, pos: np.ndarray, mate: int, err: float, first_id: int = 0) -> np.ndarray:
    """Fixed-width FASTQ records as one flat uint8 array: '@r<8 digits>_P<10 digits>\\t/<mate>\\n<seq>\\n+\\n<qual>\\n'."""
    n, L = reads.shape
    q = 33 + (int(-10.0 * np.log10(err) + 0.499) if err > 0 else 40)
    hdr = np.concatenate([
        np.full((n, 2), [ord("@"), ord("r")], dtype=np.uint8), _digits(np.arange(first_id, first_id + n), 8),
        np.full((n, 2), [ord("_"), ord("P")], dtype=np.uint8), _digits(pos, 10),
        np.full((n, 4), [9, ord("/"), 48 + mate, 10], dtype=np.uint8)], axis=1)
    rec = np.concatenate([hdr, reads, np.full((n, 3), [10, ord("+"), 10], dtype=np.uint8),
                          np.full((n, L), q, dtype=np.uint8), np.full((n, 1), 10, dtype=np.uint8)], axis=1)
    return rec.reshape(-1)


def write_fastq(path: str, reads: np.ndarray, pos: np.ndarray, mate: int, err: float) -> None:
    fastq_bytes(reads, pos, mate, err).tofile(path)


def make_reads(genome_seqs, out_prefix: str, n: int, read_len: int, err: float, seed: int, paired: bool = True, **kw):
    """Writes <out_prefix>_1.fq (and _2.fq); returns the file names."""
    r1, r2, pos = simulate(genome_seqs, n, read_len, err, seed, paired=paired, **kw)
    f1 = out_prefix + "_1.fq"
    write_fastq(f1, r1, pos, 1, err)
    if not paired:
        return f1, None
    f2 = out_prefix + "_2.fq"
    write_fastq(f2, r2, pos, 2, err)
    return f1, f2


def ensure_dir(p: str) -> str:
    os.makedirs(p, exist_ok=True)
    return p
