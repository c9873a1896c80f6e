"""Shared helpers for the parity tests, __graft_entry__.smoke() and bench.py's checker leg.

The oracle (oracle/libkartoracle.so, and oracle/_ref/* when built) is loaded ONLY here, as the checker."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kart_b200 import KartIndex, Mapper, binding, synth   # noqa: E402

ORACLE_LIB = os.path.join(ROOT, "oracle", "libkartoracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libkartref.so")
REF_KART = os.path.join(ROOT, "oracle", "_ref", "kart")
REF_BWT_INDEX = os.path.join(ROOT, "oracle", "_ref", "bwt_index")
EMUL_LIB = os.path.join(ROOT, "tests", "emul", "libkartb200_emul.so")
ECOLI_PREFIX = os.path.join(ROOT, "data", "_gen", "ecoli", "EcoliIdx")
MINI_PREFIX = os.path.join(ROOT, "tests", "golden", "mini", "mini")
GEN_DIR = os.path.join(ROOT, "data", "_gen")


def have_ecoli() -> bool:
    return os.path.exists(ECOLI_PREFIX + ".bwt")


def default_prefix() -> str:
    return ECOLI_PREFIX if have_ecoli() else MINI_PREFIX


class Oracle:
    """ctypes view of oracle/libkartoracle.so (prefix kor_) or oracle/_ref/libkartref.so (prefix kref_)."""

    def __init__(self, prefix: str, pacbio=False, max_gaps=5, multihit=False, ref=False):
        self.ref = ref
        self.lib = C.CDLL(REF_LIB if ref else ORACLE_LIB)
        self.p = "kref_" if ref else "kor_"
        load = getattr(self.lib, self.p + "load")
        rc = load(prefix.encode(), int(pacbio), max_gaps, int(multihit), 1) if ref else load(prefix.encode(), int(pacbio), max_gaps, int(multihit))
        if rc != 0:
            raise RuntimeError("oracle load failed: %d" % rc)
        # explicit prototypes: arguments beyond the sixth travel on the stack, where an untyped Python int only fills 32 of 64 bits
        sig = {"map_pair": [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_long],
               "map_single": [C.c_char_p, C.c_int, C.c_char_p, C.c_long],
               "seeds": [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_long],
               "candidates": [C.c_char_p, C.c_int, C.c_char_p, C.c_long],
               "nw": [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p],
               "fragment_pairs": [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_long],
               "normal_pairs": [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_long],
               "process_pair": [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_char_p, C.c_long]}
        for f, a in sig.items():
            fn = getattr(self.lib, self.p + f)
            fn.restype = C.c_long
            fn.argtypes = a
        self.buf = C.create_string_buffer(1 << 22)

    def fn(self, name):
        return getattr(self.lib, self.p + name)

    def map_pair(self, s1: bytes, s2rc: bytes, est: int, stage=False) -> str:
        self.fn("map_pair")(s1, len(s1), s2rc, len(s2rc), int(est), int(stage), self.buf, len(self.buf))
        return self.buf.value.decode()

    def map_single(self, s: bytes) -> str:
        self.fn("map_single")(s, len(s), self.buf, len(self.buf))
        return self.buf.value.decode()

    def seeds(self, s: bytes, sensitive=False) -> str:
        self.fn("seeds")(s, len(s), int(sensitive), self.buf, len(self.buf))
        return self.buf.value.decode()

    def candidates(self, s: bytes) -> str:
        self.fn("candidates")(s, len(s), self.buf, len(self.buf))
        return self.buf.value.decode()

    def nw(self, a: bytes, b: bytes):
        o1 = C.create_string_buffer(len(a) + len(b) + 2)
        o2 = C.create_string_buffer(len(a) + len(b) + 2)
        self.fn("nw")(a, len(a), b, len(b), o1, o2)
        return o1.value.decode(), o2.value.decode()

    def counters(self, reset=True):
        a = (C.c_ulonglong * 7)()
        self.lib.kor_counters(a, int(reset))
        return dict(zip(["searches", "ext_steps", "occ_blocks64", "locates", "lf_steps", "nw_calls", "nw_cells"], [int(x) for x in a]))


def genome_of(idx: KartIndex):
    """Forward strand of every sequence of the index as upper-case uint8 arrays (decoded from the 2-bit .pac)."""
    pac = idx.pac
    pos = np.arange(idx.l_pac, dtype=np.int64)
    codes = (pac[pos >> 2] >> ((~pos & 3) << 1).astype(np.uint8)) & 3
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    out, s = [], 0
    for ln in idx.chr_len:
        out.append(bases[s:s + ln])
        s += ln
    return out


def pac_genome(idx: KartIndex):
    """The same genome for synth.simulate() without decoding it (multi-Gbp indexes)."""
    return synth.PacGenome(idx.pac, idx.l_pac, idx.chr_len)


def ensure_syn_index(mbp: int = 3100, contigs: int = 24, seed: int = 12345):
    """Prefix of the synthetic <mbp> Mbp index (BASELINE configs 3-5 use 3100), built on first use with this repo's `kart index -gpu`
    (seconds; the host builder when no device is usable: scripts/make_syn_index.py) and cached under data/_gen/syn/ -- the files are too large to travel with a snapshot. Returns None,
    with the reason printed, when the host cannot build it."""
    import subprocess
    prefix = os.path.join(GEN_DIR, "syn", "syn%d" % mbp)
    if all(os.path.exists(prefix + e) for e in (".bwt", ".sa", ".pac", ".ann")):
        return prefix
    avail_gb = 0
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            avail_gb = int(ln.split()[1]) >> 20
    need = mbp * 6 // 1000 + 4   # generator + .pac + what `kart index -gpu` brings back; the host builder (fallback) checks its own ~23 GB per Gbp
    if avail_gb < need:
        print("ensure_syn_index: NOT building the %d Mbp index (needs ~%d GB of host memory; this host: %d GB)" % (mbp, need, avail_gb))
        return None
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_syn_index.py"), str(mbp), str(contigs), str(seed)], capture_output=True, text=True)
    print("ensure_syn_index: " + " | ".join(r.stdout.strip().splitlines()))
    if r.returncode != 0 or not os.path.exists(prefix + ".sa"):
        print("ensure_syn_index: build failed: %s" % (r.stderr[-400:] or r.stdout[-400:]))
        return None
    return prefix


def interleave(r1: np.ndarray, r2: np.ndarray) -> np.ndarray:
    """Pairs in the layout the C ABI expects: read 2i = mate 1, read 2i+1 = reverse-complemented mate 2."""
    n, L = r1.shape
    out = np.empty((2 * n, L), dtype=np.uint8)
    out[0::2] = r1
    out[1::2] = synth.revcomp_bytes(r2)
    return out


def make_mapper(idx: KartIndex, emul=False, expand_sa=False, **params) -> Mapper:
    m = Mapper(lib_path=EMUL_LIB if emul else None)
    m.upload_index(idx, expand_sa=expand_sa)
    m.set_params(**params)
    return m


def compare_pairs(m: Mapper, orc: Oracle, reads: np.ndarray, est: int = 1500, show: int = 3):
    """Maps interleaved pairs through the C ABI and checks every pair against the oracle dump. Returns #mismatches."""
    flat, off = Mapper.pack_reads(reads)
    aln, pairs, cig = m.map_chunk(flat, off, est)
    st = m.dump_state()
    bad = 0
    for p in range(len(reads) // 2):
        exp = orc.map_pair(reads[2 * p].tobytes(), reads[2 * p + 1].tobytes(), est)
        got = binding.dump_read(st, 2 * p) + binding.dump_read(st, 2 * p + 1) + "P %d %d\n" % (pairs[p]["counted"], pairs[p]["absdist"])
        if exp != got:
            bad += 1
            if bad <= show:
                print("PAIR %d differs\n--- oracle\n%s--- kart_b200\n%s" % (p, exp, got))
    return bad


def compare_singles(m: Mapper, orc: Oracle, reads, show: int = 3):
    flat, off = Mapper.pack_reads(reads)
    m.map_chunk(flat, off)
    st = m.dump_state()
    bad = 0
    for r in range(len(off) - 1):
        s = flat[int(off[r]):int(off[r + 1])].tobytes()
        exp = orc.map_single(s)
        got = binding.dump_read(st, r)
        if exp != got:
            bad += 1
            if bad <= show:
                print("READ %d differs\n--- oracle\n%s--- kart_b200\n%s" % (r, exp, got))
    return bad



def big_gap_reads(g):
    """2.5 kbp reads whose middle 150..420 bases are random: fragment pairs without common 8-mers, i.e. nw_alignment problems
    larger than one thread takes (warp wavefront class) next to all the small ones."""
    rng = np.random.default_rng(9)
    reads = []
    for k in range(6):
        s = g[k % 3][2000 + 500 * k:2000 + 500 * k + 2500].copy()
        a, L = 800 + 100 * k, [150, 250, 350, 200, 420, 300][k]
        s[a:a + L] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)]
        reads.append(s.tobytes())
    return reads


def partition_stress_reads(g, n=24, seed=5, dirty=True):
    """Reads whose middle stretches carry a substitution every 9..12 bases plus a few small indels: no seed survives there, so the stretch
    becomes one fragment pair > 30 x 30 full of exact runs of 8..12 bases on neighbouring diagonals -- the 8-mer partition's food (runs that
    cross the 25-position cells of part_pairs_packed, runs at the fragment ends, both strands). With `dirty`, every third read also gets an N
    and a lower-case base inside the stretch: those fragments go through the literal id scan instead."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    reads = []
    for k in range(n):
        c = g[k % len(g)]
        L = int(rng.integers(500, 2600))
        at = int(rng.integers(0, len(c) - L - 1))
        s = c[at:at + L].copy()
        a, w = int(rng.integers(60, 150)), int(rng.integers(90, L - 200))
        out, p = [s[:a]], a
        while p < a + w:
            step = int(rng.integers(9, 13))   # below MinSeedLength (13 on the small genomes): no seed inside the stretch
            out.append(s[p:min(p + step, a + w)])
            p += step
            if p >= a + w:
                break
            u = rng.random()
            if u < 0.75:
                out.append(acgt[(np.searchsorted(acgt, s[p]) + 1 + rng.integers(0, 3)) % 4][None]); p += 1     # substitution
            elif u < 0.88:
                out.append(acgt[rng.integers(0, 4, size=int(rng.integers(1, 4)))])                              # insertion
            else:
                p += int(rng.integers(1, 4))                                                                     # deletion
        out.append(s[a + w:])
        r = np.concatenate(out)
        if dirty and k % 3 == 2:
            r[a + w // 2] = ord("N"); r[a + w // 3] = r[a + w // 3] + 32
        if k % 2:
            r = comp[r[::-1]] if not (dirty and k % 3 == 2) else r
        reads.append(r.tobytes())
    return reads


def repeat_rich_text(length=120000, seed=3):
    """A genome for the rescue stress tests: random bases with 150 micro-satellites / homopolymers (units of 1..6 bases, 20..200 bases
    long: mates whose 8-mers repeat, long chains in the mate index) and 60 copies of a 300-bp family at 4 % divergence."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, length)
    for _ in range(150):
        p = rng.integers(0, length - 400); unit = rng.integers(0, 4, rng.integers(1, 7)); n = rng.integers(20, 200)
        g[p:p + n] = np.resize(unit, n)
    fam = rng.integers(0, 4, 300)
    for _ in range(60):
        p = rng.integers(0, length - 400); c = fam.copy(); m = rng.random(300) < 0.04
        c[m] = rng.integers(0, 4, int(m.sum())); g[p:p + 300] = c
    return "".join("ACGT"[c] for c in g)


def bam_equal(path, golden):
    """Byte-identical when the zlib in this process matches the one the golden was deflated with; always identical in the BGZF
    block structure (ISIZE sequence) and in the inflated BAM payload."""
    import gzip, struct, zlib
    a, b = open(path, "rb").read(), open(golden, "rb").read()

    def isizes(d):
        out, p = [], 0
        while p < len(d):
            bs = struct.unpack("<H", d[p + 16:p + 18])[0] + 1
            out.append(struct.unpack("<I", d[p + bs - 4:p + bs])[0])
            p += bs
        return out
    assert isizes(a) == isizes(b)
    assert gzip.decompress(a) == gzip.decompress(b)
    if zlib.ZLIB_RUNTIME_VERSION == open(os.path.join(os.path.dirname(golden), "bam_zlib.txt")).read().strip():
        assert a == b
    return True


def smoke_check(n_pairs: int = 256):
    idx = KartIndex(default_prefix())
    genome = genome_of(idx)
    r1, r2, _ = synth.simulate(genome, n_pairs, 150, 0.02, seed=11, indel=0.002)
    reads = interleave(r1, r2)
    orc = Oracle(default_prefix())
    bad = compare_pairs(make_mapper(idx, paired=True), orc, reads)                    # sampled SA: k_fm_seed, k_sa_locate
    bad += compare_pairs(make_mapper(idx, paired=True, expand_sa=True), orc, reads)   # full SA: k_fm_seed_q (what bench.py runs), k_sa_locate_reads
    return bad, n_pairs
