"""CPU: pins the oracle (our CPU restatement) against golden vectors generated from the unmodified reference
(tests/golden/make_golden.py) and, when oracle/_ref was built, against the live reference on fresh random inputs."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu
from kart_b200 import KartIndex, synth

G = os.path.join(pu.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def orc(built):
    return pu.Oracle(pu.MINI_PREFIX)


def _blocks(path):
    """yields (header fields, body text) of '# ...' delimited golden files"""
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


def test_index_files_match_reference_layout(built):
    idx = KartIndex(pu.MINI_PREFIX)
    assert idx.seq_len == 2 * idx.l_pac and idx.sa_intv == 32 and len(idx.chr_len) == 3
    assert idx.min_seed_len == 13


def test_nw_vectors(orc):
    n = 0
    for ln in open(os.path.join(G, "nw_vectors.txt")):
        a, b, o1, o2 = ln.split()
        assert orc.nw(a.encode(), b.encode()) == (o1, o2)
        n += 1
    assert n == 200


def test_seed_and_candidate_vectors(orc):
    for head, body in _blocks(os.path.join(G, "stage_seeds.txt")):
        s = head[0].encode()
        fast, rest = body[2:].split("Z\n")
        sens, cand = rest.split("C\n")
        assert orc.seeds(s, False) == fast
        assert orc.seeds(s, True) == sens
        assert orc.candidates(s) == cand


def test_fragment_partition_vectors(orc):
    orc.fn("fragment_pairs").restype = C.c_long
    for head, body in _blocks(os.path.join(G, "frag_vectors.txt")):
        shift, a, b = int(head[0]), head[1].encode(), head[2].encode()
        orc.fn("fragment_pairs")(shift, a, len(a), b, len(b), 1, orc.buf, len(orc.buf))
        assert orc.buf.value.decode() == body


def test_pair_stage_vectors(orc):
    n = 0
    for head, body in _blocks(os.path.join(G, "stage_pairs.txt")):
        assert orc.map_pair(head[0].encode(), head[1].encode(), int(head[2]), stage=True) == body
        n += 1
    assert n == 120


@pytest.mark.parametrize("tag,args,pacbio", [("pe150", ["pe150_1.fq", "pe150_2.fq"], False), ("se100", ["se100.fq", None], False), ("pb3k", ["pb3k.fq", None], True)])
def test_whole_file_sam_golden(built, tmp_path, tag, args, pacbio):
    o = pu.Oracle(pu.MINI_PREFIX, pacbio=pacbio)
    o.lib.kor_map_files.restype = C.c_long
    out = str(tmp_path / (tag + ".sam"))
    f2 = os.path.join(G, args[1]).encode() if args[1] else None
    n = o.lib.kor_map_files(os.path.join(G, args[0]).encode(), f2, 1 if f2 else 0, out.encode(), 2)
    assert n > 0
    assert open(out, "rb").read() == open(os.path.join(G, tag + ".sam"), "rb").read()


@pytest.mark.skipif(not (pu.have_ecoli() and os.path.exists("/root/reference/test/r1.fq")), reason="needs the reference fixture (build container only)")
def test_c1_run_test_fixture_md5(built, tmp_path):
    o = pu.Oracle(pu.ECOLI_PREFIX)
    o.lib.kor_map_files.restype = C.c_long
    out = str(tmp_path / "c1.sam")
    assert o.lib.kor_map_files(b"/root/reference/test/r1.fq", b"/root/reference/test/r2.fq", 1, out.encode(), 1) == 2000
    want = dict(ln.split() for ln in open(os.path.join(G, "ecoli_c1.md5")))
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == want["raw"] == "75cbcedeb1d5ebf50aeec84994b45654"


@pytest.mark.skipif(not os.path.exists(pu.REF_LIB), reason="oracle/_ref not built")
def test_oracle_vs_live_reference_random(built):
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    ref, o = pu.Oracle(pu.MINI_PREFIX, ref=True), pu.Oracle(pu.MINI_PREFIX)
    r1, r2, _ = synth.simulate(g, 300, 150, 0.04, seed=77, indel=0.006, n_rate=0.004)
    reads = pu.interleave(r1, r2)
    for p in range(300):
        a, b = reads[2 * p].tobytes(), reads[2 * p + 1].tobytes()
        assert o.map_pair(a, b, 600, stage=True) == ref.map_pair(a, b, 600, stage=True)
    s, _, _ = synth.simulate(g, 300, 100, 0.08, seed=78, paired=False)
    for r in s:
        assert o.map_single(r.tobytes()) == ref.map_single(r.tobytes())


@pytest.mark.skipif(not os.path.exists(pu.REF_LIB), reason="oracle/_ref not built")
def test_fm_primitives_vs_live_reference(built):
    ref, o = pu.Oracle(pu.MINI_PREFIX, ref=True), pu.Oracle(pu.MINI_PREFIX)
    idx = KartIndex(pu.MINI_PREFIX)
    rng = np.random.default_rng(5)
    a, b = (C.c_ulonglong * 4)(), (C.c_ulonglong * 4)()
    ref.lib.kref_sa.restype = C.c_ulonglong
    o.lib.kor_sa.restype = C.c_ulonglong
    for k in [0, 1, idx.primary - 1, idx.primary, idx.primary + 1, idx.seq_len - 1, idx.seq_len] + [int(x) for x in rng.integers(0, idx.seq_len, 300)]:
        ref.lib.kref_occ4(C.c_ulonglong(k), a)
        o.lib.kor_occ4(C.c_ulonglong(k), b)
        assert list(a) == list(b)
        assert ref.lib.kref_sa(C.c_ulonglong(k)) == o.lib.kor_sa(C.c_ulonglong(k))
