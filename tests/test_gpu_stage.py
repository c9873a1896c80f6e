"""-m gpu: stage-level parity of the CUDA path (tests/stage_util.py: seed SA intervals and positions, candidate lists, NW run
lists and identities, fragment-pair processing) against the reference's golden vectors and the pinned oracle; the C1 fixture
(run_test.sh) through the CUDA CLI; the pipelined kb_map_chunk against the oracle at full sub-batch sizes."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import parity_util as pu
import stage_util as su
from kart_b200 import KartIndex, Mapper, binding, synth

pytestmark = pytest.mark.gpu
G = os.path.join(pu.ROOT, "tests", "golden")
KART = os.path.join(pu.ROOT, "kart_b200", "bin", "kart")


def test_golden_seeds_and_candidates_gpu(built):
    su.check_golden_seeds_and_candidates(emul=False)


def test_golden_pair_stage_gpu(built):
    assert su.check_golden_pair_stage(emul=False) == 120


def test_golden_nw_vectors_gpu(built, tmp_path, monkeypatch):
    """k_nw_tile<0..5>, k_nw_warp fed with the reference's own nw_alignment vectors"""
    n, seen = su.check_nw_vectors(False, str(tmp_path), monkeypatch)
    assert n > 250 and (seen > 0).all()


def test_fragment_pairs_vs_oracle_gpu(built):
    """k_align_part, the solvers and k_align_gather on chosen fragment pairs, plus the head / tail rules"""
    assert su.check_fragment_pairs(emul=False) > 600


@pytest.mark.skipif(not pu.have_ecoli(), reason="E. coli index (data/_gen/ecoli, built by __graft_entry__.build() from the reference's test/ecoli.fa) is absent")
@pytest.mark.parametrize("threads", ["4", "1"])
def test_c1_run_test_fixture_md5_gpu(built, tmp_path, threads):
    """BASELINE config 1 = the reference's run_test.sh:28 (`kart -i EcoliIdx -f r1.fq -f2 r2.fq -t 4`): the CUDA CLI writes the
    reference's bytes (md5 75cbcedeb1d5ebf50aeec84994b45654; reads committed under tests/golden/c1)."""
    out = str(tmp_path / "alignment.sam")
    subprocess.run([KART, "-silent", "-t", threads, "-i", pu.ECOLI_PREFIX, "-f", os.path.join(G, "c1", "r1.fq"), "-f2", os.path.join(G, "c1", "r2.fq"), "-o", out],
                   check=True, stdout=subprocess.DEVNULL)
    want = dict(ln.split() for ln in open(os.path.join(G, "ecoli_c1.md5")))
    assert want["raw"] == "75cbcedeb1d5ebf50aeec84994b45654"
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == want["raw"]


def _oracle_lines(orc, reads, est):
    """per read what the SAM line needs, from the oracle's dump: (score, sub, mapq, mapped, flag, chr, pos, cigar)"""
    out = []
    for p in range(len(reads) // 2):
        txt = orc.map_pair(reads[2 * p].tobytes(), reads[2 * p + 1].tobytes(), est)
        cur = None
        for ln in txt.splitlines():
            f = ln.split()
            if f[0] == "R":
                cur = [int(f[1]), int(f[2]), int(f[3]), False, 0, 0, 0, ""]
                best = int(f[5]); out.append(cur)
            elif f[0] == "A" and len(f) > 4 and f[4].startswith("F") and (int(f[1]) == best or cur[0] == 0):
                cur[4] = int(f[4][1:])
                if int(f[2]) > 0 and cur[0] > 0:
                    cur[3] = True; cur[5] = int(f[6]); cur[6] = int(f[7]); cur[7] = f[8]
    return out


@pytest.mark.skipif(not pu.have_ecoli(), reason="E. coli index absent")
def test_pipelined_chunk_vs_oracle_full_size(built):
    """The slot pipeline of kb_map_chunk (what bench.py's e2e leg and the CLI's large batches run) against the ORACLE, not against
    itself: 280 000 reads (>= pipe_min_reads = 262 144), default plan, every read's line compared."""
    idx = KartIndex(pu.ECOLI_PREFIX)
    g = pu.genome_of(idx)
    r1, r2, _ = synth.simulate(g, 140000, 150, 0.02, seed=78, indel=0.001)
    reads = pu.interleave(r1, r2)
    flat, off = Mapper.pack_reads(reads)
    m = pu.make_mapper(idx, expand_sa=True, paired=True)
    aln, pairs, cig = m.map_chunk(flat, off, np.full(140000, 1500, dtype=np.int32), packed=True)   # the packed entry point: what bench.py's e2e leg calls
    assert m.work()["launches"] > 2 * 19, "the chunk did not go through the slot pipeline"
    exp = _oracle_lines(pu.Oracle(pu.ECOLI_PREFIX), reads, 1500)
    assert len(exp) == len(aln)
    bad = 0
    for r, e in enumerate(exp):
        a = aln[r]
        ok = (int(a["score"]), int(a["sub_score"]), int(a["mapq"])) == tuple(e[:3])
        if ok and e[3]:
            ok = int(a["kind"]) == 1 and (int(a["flag"]), int(a["chr"]), int(a["pos"])) == tuple(e[4:7]) and binding.cigar_string(cig, int(a["cig_off"]), int(a["cig_len"])) == e[7]
        elif ok and e[0] == 0:
            ok = int(a["kind"]) == 0 and int(a["flag"]) == e[4]
        if not ok:
            bad += 1
            if bad <= 3:
                print("read", r, "oracle", e, "gpu", a)
    assert bad == 0
