"""Stage-level parity checks shared by the CPU suite (host-emulated kernels, tests/emul) and the -m gpu suite (the real kernels):
what BASELINE.json's north_star lists next to the final SAM -- seed SA intervals and positions, candidate lists, NW scores and
CIGARs -- each compared with the golden vectors generated from the unmodified reference (tests/golden/make_golden.py) or with the
pinned oracle. Reference hooks: src/bwt_search.cpp:171-181 (search results), src/AlignmentCandidates.cpp:77,129 (seed and candidate
lists), src/nw_alignment.cpp:18-72, src/tools.cpp:142-404 (fragment pairs)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import parity_util as pu
from kart_b200 import KartIndex, Mapper, binding

G = os.path.join(pu.ROOT, "tests", "golden")
KART = os.path.join(pu.ROOT, "kart_b200", "bin", "kart")
NT4 = {c: i for i, c in enumerate(b"ACGT")}
NT4.update({c: i for i, c in enumerate(b"acgt")})


def blocks(path):
    """(header fields, body text) of the '# ...' delimited golden files"""
    head, body = None, []
    for ln in open(path):
        if ln.startswith("#"):
            if head is not None:
                yield head, "".join(body)
            head, body = ln[2:].rstrip("\n").split(" "), []
        else:
            body.append(ln)
    if head is not None:
        yield head, "".join(body)


def seg_lines(st, r):
    """the read's seeds as the oracle prints them: S rPos rLen gLen gPos simple"""
    o, n = int(st["seed_off"][r]), int(st["n_seeds"][r])
    return ["S %d %d %d %d %d" % (s["rpos"], s["rlen"], s["glen"], s["gpos"], s["simple"]) for s in st["segs"][o:o + n]]


def prune(cands, pacbio):
    """RemoveRedundantCandidates (src/Mapping.cpp:317-346) on [score, ...] lists: scores below the threshold become 0"""
    if len(cands) <= 1:
        return
    s1 = s2 = 0
    for c in cands:
        if c[0] > s2:
            if c[0] >= s1:
                s2, s1 = s1, c[0]
            else:
                s2 = c[0]
    thr = s1 if (pacbio or s1 == s2 or s1 - s2 > 20) else s2
    for c in cands:
        if c[0] < thr:
            c[0] = 0


def parse_cands(text):
    out = []
    for ln in text.splitlines():
        f = ln.split()
        if f[0] == "C":
            out.append([int(f[1]), int(f[2]), int(f[3]), int(f[4]), []])
        else:
            out[-1][4].append(ln)
    return out


def fmt_cands(cands):
    out = []
    for c in cands:
        out.append("C %d %d %d %d" % tuple(c[:4]))
        out.extend(c[4])
    return "\n".join(out) + ("\n" if out else "")


def fast_mode_searches(orc, read: bytes, min_seed: int):
    """IdentifySeedPairs_FastMode's driver (src/AlignmentCandidates.cpp:49-75) over the oracle's BWT_Search: the searches that
    yield seeds as (rPos, len, freq, x0)"""
    codes = np.array([NT4.get(c, 4) for c in read], dtype=np.uint8)
    ln, fq, x0, x2 = C.c_int(), C.c_int(), C.c_ulonglong(), C.c_ulonglong()
    locs = (C.c_ulonglong * 64)()
    fn = orc.lib.kor_bwt_search
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    fn.restype = None
    out, pos, rlen = [], 0, len(read)
    while pos < rlen - min_seed:
        if codes[pos] > 3:
            pos += 1
            continue
        fn(codes.ctypes.data, pos, rlen, C.byref(ln), C.byref(fq), locs, C.byref(x0), C.byref(x2))
        if fq.value > 0:
            out.append((pos, ln.value, fq.value, x0.value))
        pos += ln.value + 1
    return out


def check_golden_seeds_and_candidates(emul: bool):
    """stage_seeds.txt: per read the seed list (fast and sensitive mode) and the candidate list of the reference itself."""
    idx = KartIndex(pu.MINI_PREFIX)
    gold = list(blocks(os.path.join(G, "stage_seeds.txt")))
    reads = [h[0].encode() for h, _ in gold]
    flat, off = Mapper.pack_reads(reads)
    orc = pu.Oracle(pu.MINI_PREFIX)
    n_seeds = n_cands = 0
    for full_sa in (False, True):   # k_fm_seed + k_sa_locate, then k_fm_seed_q + k_sa_locate_reads
        m = pu.make_mapper(idx, emul=emul, expand_sa=full_sa, paired=False)
        m.map_chunk(flat, off)
        st, hits = m.dump_state(), m.hits()
        for r, (_, body) in enumerate(gold):
            fast, rest = body[2:].split("Z\n")
            _, cand = rest.split("C\n")
            assert sorted(seg_lines(st, r)) == sorted(fast.splitlines()), "seeds of read %d" % r
            assert hits[r] == fast_mode_searches(orc, reads[r], m.min_seed_len), "SA intervals of read %d" % r
            exp = parse_cands(cand)
            prune(exp, False)
            assert binding.dump_cands(st, r) == fmt_cands(exp), "candidates of read %d" % r
            n_seeds += len(fast.splitlines()); n_cands += len(exp)
        mp = pu.make_mapper(idx, emul=emul, expand_sa=full_sa, pacbio=True)
        mp.map_chunk(flat, off)
        st = mp.dump_state()
        for r, (_, body) in enumerate(gold):
            sens = body[2:].split("Z\n")[1].split("C\n")[0]
            assert sorted(seg_lines(st, r)) == sorted(sens.splitlines()), "sensitive-mode seeds of read %d" % r
    assert n_seeds > 100 and n_cands >= 60
    return n_seeds, n_cands


def check_golden_pair_stage(emul: bool):
    """stage_pairs.txt: candidate lists of both mates after pairing / rescue / pruning, then reports and the insert-size
    contribution, for 120 pairs mapped by the reference with two EstDistance values."""
    idx = KartIndex(pu.MINI_PREFIX)
    gold = list(blocks(os.path.join(G, "stage_pairs.txt")))
    reads, est = [], []
    for h, _ in gold:
        reads += [h[0].encode(), h[1].encode()]
        est.append(int(h[2]))
    flat, off = Mapper.pack_reads(reads)
    m = pu.make_mapper(idx, emul=emul, paired=True)
    aln, pairs, cig = m.map_chunk(flat, off, np.array(est, dtype=np.int32))
    st = m.dump_state()
    for p, (_, body) in enumerate(gold):
        lines = body.split("\n")
        assert lines[0].startswith("X ") and lines[1] == "V1"
        got = "V1\n" + binding.dump_cands(st, 2 * p) + "V2\n" + binding.dump_cands(st, 2 * p + 1) + binding.dump_read(st, 2 * p) + binding.dump_read(st, 2 * p + 1) \
            + "P %d %d\n" % (pairs[p]["counted"], pairs[p]["absdist"])
        assert got == body[len(lines[0]) + 1:], "pair %d" % p
    return len(gold)


def ops_of_gapped(o1: str, o2: str):
    """AddNewCigarElements (src/tools.cpp:49-104) on the two gapped strings nw_alignment returns: (cigar string, identities)"""
    ops, ident = [], 0
    for a, b in zip(o1, o2):
        op = "D" if a == "-" else ("I" if b == "-" else "M")
        if op == "M" and a == b:
            ident += 1
        if ops and ops[-1][1] == op:
            ops[-1][0] += 1
        else:
            ops.append([1, op])
    return "".join("%d%s" % (n, o) for n, o in ops), ident


def text_index(tmp, texts):
    """an index whose forward strand is the concatenation of `texts` (pure ACGT), built by this repo's `kart index`; returns
    (prefix, start of every text)"""
    fa = os.path.join(tmp, "t.fa")
    starts, at = [], 0
    with open(fa, "w") as fh:
        fh.write(">t\n")
        for t in texts:
            starts.append(at)
            fh.write(t + "\n")
            at += len(t)
    prefix = os.path.join(tmp, "t")
    subprocess.run([KART, "index", fa, prefix], check=True, stdout=subprocess.DEVNULL)
    return prefix, starts


def check_nw_vectors(emul: bool, tmp: str, monkeypatch):
    """nw_vectors.txt + nw_vectors_large.txt: every vector whose reference side is pure ACGT goes to ONE nw_alignment call of the
    device code (mode 4), under the settings that route the 33..128 classes to the thread tiles, to the warp wavefront, and
    everything above 32 to the wavefront: run list and identity count must equal what the reference returned."""
    vec = []
    for f in ("nw_vectors.txt", "nw_vectors_large.txt"):
        for ln in open(os.path.join(G, f)):
            a, b, o1, o2 = ln.split()
            if set(b) <= set("ACGT"):
                vec.append((a, b) + ops_of_gapped(o1, o2))
    assert len(vec) > 250
    prefix, starts = text_index(tmp, [v[1] for v in vec])
    idx = KartIndex(prefix)
    flat, off = Mapper.pack_reads([v[0].encode() for v in vec])
    frags = [(i, 0, len(v[0]), starts[i], len(v[1]), 4) for i, v in enumerate(vec)]
    seen = np.zeros(7, dtype=np.int64)
    for env in ({}, {"KB_NW_WARP_BELOW": "0"}, {"KB_NW_TMAX": "32"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = pu.make_mapper(idx, emul=emul, paired=False)
        m.stage(flat, off)
        out, cig = m.debug_align(frags)
        cls = m.debug(9, np.uint32, 32)[16:23].astype(np.int64)
        for i, v in enumerate(vec):
            assert cig[i] == v[2] and int(out[i]["ident"]) == v[3] and int(out[i]["score"]) == v[3], (i, v[0], v[1], cig[i], v[2], int(out[i]["ident"]), v[3])
        # which solver took what: tiles own classes 0..5 by default only when a class is dense, KB_NW_WARP_BELOW=0 gives them 4 and 5
        assert cls.sum() == len(vec)
        if env.get("KB_NW_TMAX") == "32":
            assert cls[4] == 0 and cls[5] == 0 and cls[6] > 40
        else:
            assert (cls > 0).all(), cls
        seen += cls
        for k in env:
            monkeypatch.delenv(k)
    return len(vec), seen


def check_fragment_pairs(emul: bool):
    """frag_vectors.txt inputs (read fragment, substring of the mini genome) as middle / head / tail pairs and straight through
    GenerateNormalPairAlignment, in Illumina and -pacbio mode, against the oracle's Process{Normal,Head,Tail}SequencePair."""
    idx = KartIndex(pu.MINI_PREFIX)
    cat = b"".join(x.tobytes() for x in pu.genome_of(idx))
    vec = []
    for h, _ in blocks(os.path.join(G, "frag_vectors.txt")):
        a, b = h[1], h[2]
        g = cat.find(b.encode())
        assert g >= 0
        vec.append((a, g, len(b)))
    # plus short and odd shapes: equal-length near-identical (quick test), 1 x 1, pure gaps, the head/tail length limits
    rng = np.random.default_rng(7)
    for k in range(60):
        L = int(rng.integers(1, 130)); g = int(rng.integers(1000, len(cat) - 1000))
        a = bytearray(cat[g:g + L])
        for _ in range(int(rng.integers(0, 4))):
            a[int(rng.integers(0, L))] = b"ACGT"[int(rng.integers(0, 4))]
        if k % 5 == 0 and L > 6:
            del a[3:3 + int(rng.integers(1, 3))]
        vec.append((bytes(a).decode(), g, L))
    flat, off = Mapper.pack_reads([v[0].encode() for v in vec])
    n = 0
    for pacbio in (False, True):
        orc = pu.Oracle(pu.MINI_PREFIX, pacbio=pacbio)
        m = pu.make_mapper(idx, emul=emul, pacbio=pacbio, paired=False)
        m.stage(flat, off)
        for mode in (0, 1, 2):
            frags = [(i, 0, len(v[0]), v[1], v[2], mode) for i, v in enumerate(vec)]
            out, cig = m.debug_align(frags)
            for i, v in enumerate(vec):
                orc.fn("process_pair")(mode, v[0].encode(), 0, len(v[0]), v[1], v[2], orc.buf, len(orc.buf))
                exp = orc.buf.value.decode().splitlines()
                p = exp[0].split()
                score, gpos, glen = int(p[1]), int(p[4]), int(p[5])
                ecig = "".join("%s%s" % tuple(ln.split()[1:3]) for ln in exp[1:])
                assert cig[i] == ecig and int(out[i]["score"]) == score, (pacbio, mode, i, v, cig[i], ecig, int(out[i]["score"]), score)
                if mode == 1 and score > 0:
                    assert int(out[i]["g_first"]) == gpos, (pacbio, i, v)
                if mode == 2 and score > 0:
                    assert int(out[i]["g_end"]) == gpos + glen - 1, (pacbio, i, v)
                n += 1
    return n
