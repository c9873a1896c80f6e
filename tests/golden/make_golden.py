#!/usr/bin/env python
"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from
/root/reference). Run in the build container only:   python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  mini/mini.{fa,bwt,sa,pac,ann,amb}   synthetic 3-contig genome with repeats, indexed by the reference's bwt_index
  pe150_{1,2}.fq + pe150.sam          400 pairs 2x150 @2 % (+indels, N) mapped by `kart -t 1`
  se100.fq + se100.sam                300 single-end 100 bp reads @8 %
  pb3k.fq + pb3k.sam                  12 long reads (3 kbp @15 %) mapped with -pacbio
  stage_pairs.txt                     per-pair stage dumps (candidates before reports + final reports) of the first 120 pairs
  stage_seeds.txt                     seed lists (fast + sensitive mode) and candidate lists of 60 reads
  nw_vectors.txt                      200 nw_alignment input/output pairs
  frag_vectors.txt                    8-mer partition (+IdentifyNormalPairs) of 60 fragment pairs
  nw_vectors_large.txt                140 more nw_alignment pairs covering every size class of the CUDA solvers (1..8 up to
                                      129..400 on the longer side), reference side pure ACGT (`make_golden.py nwlarge`)
  c1/r1.fq, c1/r2.fq                  the reference's own run_test.sh reads (test/r1.fq, test/r2.fq), for the C1 md5 test on the GPU box
  ecoli_c1.md5                        md5 of the reference SAM for run_test.sh (C1), raw and `LC_ALL=C sort`ed
  dup/dup.* pe150m_{1,2}.fq pe150m.sam pe150m.bam se100m.sam pb3km.sam   -m (multiple alignments) runs, `make_golden.py multihit`
  pe150.bam se100.bam pb3k.bam        the same three runs with `-bo` (reference linked against its vendored htslib 1.5:
                                      oracle/_ref/kart_hts); bam_zlib.txt records the zlib version the bytes depend on
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu   # noqa: E402
from kart_b200 import KartIndex, synth   # noqa: E402


def kart(prefix, args, out):
    subprocess.run([pu.REF_KART, "-silent", "-t", "1", "-i", prefix] + args + ["-o", out], check=True, stdout=subprocess.DEVNULL)


def kart_bam(prefix, args, out):
    subprocess.run([os.path.join(os.path.dirname(pu.REF_KART), "kart_hts"), "-silent", "-t", "1", "-i", prefix] + args + ["-bo", out], check=True, stdout=subprocess.DEVNULL)


def bam_goldens():
    import zlib
    g = HERE + "/"
    kart_bam(pu.MINI_PREFIX, ["-f", g + "pe150_1.fq", "-f2", g + "pe150_2.fq"], g + "pe150.bam")
    kart_bam(pu.MINI_PREFIX, ["-f", g + "se100.fq"], g + "se100.bam")
    kart_bam(pu.MINI_PREFIX, ["-pacbio", "-f", g + "pb3k.fq"], g + "pb3k.bam")
    open(g + "bam_zlib.txt", "w").write(zlib.ZLIB_RUNTIME_VERSION + "\n")


def multihit_goldens():
    """-m: dup/ = 60 kbp, 2 contigs with EXACT repeats (so that whole pairs tie), 300 pairs @0.5 % -> pe150m.sam / .bam; the mini
    single-end and pacbio sets again with -m -> se100m.sam, pb3km.sam."""
    g = HERE + "/"
    dup = os.path.join(HERE, "dup"); os.makedirs(dup, exist_ok=True)
    names, seqs = synth.make_genome(60_000, 2, seed=777, repeats=((2500, 4, 0.0), (600, 8, 0.0), (300, 10, 0.01)))
    synth.write_fasta(os.path.join(dup, "dup.fa"), names, seqs)
    prefix = os.path.join(dup, "dup")
    subprocess.run([pu.REF_BWT_INDEX, prefix + ".fa", prefix], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    f1, f2 = synth.make_reads(seqs, g + "pe150m", 300, 150, 0.005, seed=201, indel=0.001)
    kart(prefix, ["-m", "-f", f1, "-f2", f2], g + "pe150m.sam")
    kart_bam(prefix, ["-m", "-f", f1, "-f2", f2], g + "pe150m.bam")
    kart(pu.MINI_PREFIX, ["-m", "-f", g + "se100.fq"], g + "se100m.sam")
    kart(pu.MINI_PREFIX, ["-m", "-pacbio", "-f", g + "pb3k.fq"], g + "pb3km.sam")


def nw_large_goldens():
    """nw_alignment on problems of every size class the CUDA path distinguishes; s2 (the reference side) is pure ACGT because the
    GPU test feeds it as an indexed text; s1 keeps the occasional N and lower-case character."""
    ref = pu.Oracle(pu.MINI_PREFIX, ref=True)
    rng = np.random.default_rng(106)
    acgt = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
    with open(os.path.join(HERE, "nw_vectors_large.txt"), "w") as fh:
        for cls, (lo, hi) in enumerate(((1, 8), (9, 16), (17, 24), (25, 32), (33, 64), (65, 128), (129, 400))):
            for k in range(20):
                mx = int(rng.integers(lo, hi + 1)); mn = int(rng.integers(max(1, mx // 3), mx + 1))
                m, n = (mx, mn) if k % 2 else (mn, mx)
                b = acgt[rng.integers(0, 4, size=n)]
                if k % 4 < 3:   # related sequences: substitutions plus a few indels
                    a = b.copy()
                    mask = rng.random(len(a)) < 0.12
                    a[mask] = acgt[rng.integers(0, 4, size=int(mask.sum()))]
                    a = a[rng.random(len(a)) > 0.04]
                    ins = rng.integers(0, len(a) + 1, size=int(rng.integers(0, 3)))
                    for p_ in sorted(ins.tolist(), reverse=True):
                        a = np.concatenate([a[:p_], acgt[rng.integers(0, 4, size=int(rng.integers(1, 6)))], a[p_:]])
                    a = a[:m] if len(a) >= m else np.concatenate([a, acgt[rng.integers(0, 4, size=m - len(a))]])
                else:
                    a = acgt[rng.integers(0, 4, size=m)]
                odd = rng.random(len(a)) < 0.02
                a = a.copy(); a[odd] = acgt[rng.integers(4, 9, size=int(odd.sum()))]
                o1, o2 = ref.nw(a.tobytes(), b.tobytes())
                fh.write("%s %s %s %s\n" % (a.tobytes().decode(), b.tobytes().decode(), o1, o2))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "nwlarge":
        return nw_large_goldens()
    if len(sys.argv) > 1 and sys.argv[1] == "bam":
        return bam_goldens()
    if len(sys.argv) > 1 and sys.argv[1] == "multihit":
        return multihit_goldens()
    assert os.path.exists(pu.REF_KART) and os.path.exists(pu.REF_LIB), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    mini = os.path.join(HERE, "mini")
    os.makedirs(mini, exist_ok=True)
    names, seqs = synth.make_genome(150_000, 3, seed=4242, repeats=((1500, 6, 0.02), (250, 40, 0.05)))
    synth.write_fasta(os.path.join(mini, "mini.fa"), names, seqs)
    subprocess.run([pu.REF_BWT_INDEX, os.path.join(mini, "mini.fa"), pu.MINI_PREFIX], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    g = HERE + "/"
    f1, f2 = synth.make_reads(seqs, g + "pe150", 400, 150, 0.02, seed=101, indel=0.004, n_rate=0.002)
    kart(pu.MINI_PREFIX, ["-f", f1, "-f2", f2], g + "pe150.sam")
    s1, _ = synth.make_reads(seqs, g + "se100", 300, 100, 0.08, seed=102, paired=False, indel=0.002)
    os.replace(s1, g + "se100.fq")
    kart(pu.MINI_PREFIX, ["-f", g + "se100.fq"], g + "se100.sam")
    p1, _ = synth.make_reads(seqs, g + "pb3k", 12, 3000, 0.15, seed=103, paired=False, indel=0.01)
    os.replace(p1, g + "pb3k.fq")
    kart(pu.MINI_PREFIX, ["-pacbio", "-f", g + "pb3k.fq"], g + "pb3k.sam")

    ref = pu.Oracle(pu.MINI_PREFIX, ref=True)
    r1, r2, _ = synth.simulate(seqs, 120, 150, 0.03, seed=104, indel=0.004, n_rate=0.002)
    reads = pu.interleave(r1, r2)
    with open(g + "stage_pairs.txt", "w") as fh:
        for p in range(120):
            a, b = reads[2 * p].tobytes(), reads[2 * p + 1].tobytes()
            fh.write("# %s %s %d\n" % (a.decode(), b.decode(), 1500 if p % 3 else 480))
            fh.write(ref.map_pair(a, b, 1500 if p % 3 else 480, stage=True))
    with open(g + "stage_seeds.txt", "w") as fh:
        for r in range(60):
            s = reads[r].tobytes()
            fh.write("# %s\n" % s.decode())
            fh.write("F\n" + ref.seeds(s, False) + "Z\n" + ref.seeds(s, True) + "C\n" + ref.candidates(s))
    rng = np.random.default_rng(105)
    acgt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    with open(g + "nw_vectors.txt", "w") as fh:
        for k in range(200):
            m, n = int(rng.integers(1, 60)), int(rng.integers(1, 60))
            a = acgt[rng.choice(5, size=m, p=[.245, .245, .245, .245, .02])]
            b = a.copy() if k % 2 else acgt[rng.integers(0, 4, size=n)]
            if k % 2:   # mutate a copy: substitutions and an indel
                b = b[rng.random(len(b)) > 0.05]
                mask = rng.random(len(b)) < 0.1
                b[mask] = acgt[rng.integers(0, 4, size=int(mask.sum()))]
                if len(b) == 0:
                    b = acgt[:1].copy()
            o1, o2 = ref.nw(a.tobytes(), b.tobytes())
            fh.write("%s %s %s %s\n" % (a.tobytes().decode(), b.tobytes().decode(), o1, o2))
    ref.fn("fragment_pairs").restype = __import__("ctypes").c_long
    with open(g + "frag_vectors.txt", "w") as fh:
        cat = np.concatenate(seqs)
        for k in range(60):
            L = int(rng.integers(31, 400))
            p = int(rng.integers(0, len(cat) - 500))
            b = cat[p:p + L].copy()
            a = b.copy()
            mask = rng.random(L) < 0.06
            a[mask] = acgt[rng.integers(0, 5, size=int(mask.sum()))]
            if k % 3 == 0:
                cut = int(rng.integers(5, L - 5))
                a = np.concatenate([a[:cut], a[cut + int(rng.integers(1, 4)):]])
            shift = 5 if k % 2 else 50
            ref.fn("fragment_pairs")(shift, a.tobytes(), len(a), b.tobytes(), len(b), 1, ref.buf, len(ref.buf))
            fh.write("# %d %s %s\n" % (shift, a.tobytes().decode(), b.tobytes().decode()))
            fh.write(ref.buf.value.decode())
    if pu.have_ecoli() and os.path.exists("/root/reference/test/r1.fq"):
        out = "/tmp/golden_c1.sam"
        subprocess.run([pu.REF_KART, "-silent", "-t", "1", "-i", pu.ECOLI_PREFIX, "-f", "/root/reference/test/r1.fq", "-f2", "/root/reference/test/r2.fq", "-o", out],
                       check=True, stdout=subprocess.DEVNULL)
        raw = open(out, "rb").read()
        srt = b"".join(sorted(raw.splitlines(keepends=True)))
        open(g + "ecoli_c1.md5", "w").write("raw %s\nsorted %s\n" % (hashlib.md5(raw).hexdigest(), hashlib.md5(srt).hexdigest()))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
