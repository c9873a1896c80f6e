"""-m gpu: `kart index -gpu` (csrc/kb_index_build.cu: suffix array by MSD radix sorting in HBM, BWT / Occ / SA samples on the device)
writes the reference builder's bytes: against the committed reference-built indexes (tests/golden/mini, dup; E. coli and the 100 Mbp
genome when present) and against this repo's host builder on awkward texts (tiny, odd lengths, all-A, tandem repeats, exact
duplicates), with one and with many super-buckets."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
G = os.path.join(pu.ROOT, "tests", "golden")
KART = os.path.join(pu.ROOT, "kart_b200", "bin", "kart")


def _same(a, b, exts):
    for e in exts:
        x, y = open(a + e, "rb").read(), open(b + e, "rb").read()
        assert x == y, "%s differs (%d vs %d bytes)" % (e, len(x), len(y))


@pytest.mark.parametrize("bucket", [None, "5000"])
@pytest.mark.parametrize("name", ["mini", "dup"])
def test_gpu_index_from_fasta_equals_reference_builder(built, tmp_path, monkeypatch, name, bucket):
    if bucket:
        monkeypatch.setenv("KB_INDEX_BUCKET", bucket)   # many super-buckets on a small genome
    ref = os.path.join(G, name, name)
    out = str(tmp_path / name)
    subprocess.run([KART, "index", "-gpu", ref + ".fa", out], check=True, stdout=subprocess.DEVNULL)
    _same(out, ref, [".bwt", ".sa", ".pac", ".ann", ".amb"])


@pytest.mark.parametrize("prefix", [pu.ECOLI_PREFIX, os.path.join(pu.GEN_DIR, "syn", "syn100")])
def test_gpu_index_from_pac_equals_reference_builder(built, tmp_path, prefix):
    if not os.path.exists(prefix + ".bwt"):
        pytest.skip("%s is not on this box" % prefix)
    out = str(tmp_path / "x")
    for e in (".pac", ".ann", ".amb"):
        shutil.copy(prefix + e, out + e)
    subprocess.run([KART, "index", "-gpu", "-pac", out], check=True, stdout=subprocess.DEVNULL)
    _same(out, prefix, [".bwt", ".sa"])


def _write_fa(path, seqs):
    with open(path, "w") as fh:
        for i, s in enumerate(seqs):
            fh.write(">c%d\n" % i)
            for k in range(0, len(s), 60):
                fh.write(s[k:k + 60] + "\n")


@pytest.mark.parametrize("bucket", [None, "1024"])
def test_gpu_index_awkward_texts_equal_host_builder(built, tmp_path, monkeypatch, bucket):
    if bucket:
        monkeypatch.setenv("KB_INDEX_BUCKET", bucket)
    rng = np.random.default_rng(3)
    rnd = lambda n: "".join("ACGT"[c] for c in rng.integers(0, 4, size=n))
    unit = rnd(37)
    dup = rnd(700)
    cases = {"one": ["A"], "five": ["ACGTT"], "l33": [rnd(33)], "l127": [rnd(127)], "l128": [rnd(64), rnd(64)], "l129": [rnd(129)], "l4099": [rnd(4099)],
             "allA": ["A" * 3001], "allT": ["T" * 2000], "AT": ["AT" * 1500], "tandem": [unit * 90 + rnd(100)], "dups": [rnd(300) + dup + rnd(200) + dup + rnd(111), dup],
             "withN": [rnd(500) + "NNNNNNNNNN" + rnd(300) + "RY" + rnd(77)], "palin": ["ACGT" * 300 + "GAATTC" * 100]}
    bad = []
    for name, seqs in cases.items():
        fa = str(tmp_path / (name + ".fa"))
        _write_fa(fa, seqs)
        host, dev = str(tmp_path / (name + "_h")), str(tmp_path / (name + "_d"))
        subprocess.run([KART, "index", fa, host], check=True, stdout=subprocess.DEVNULL)
        r = subprocess.run([KART, "index", "-gpu", fa, dev], capture_output=True, text=True, env=dict(os.environ, KB_INDEX_TRACE="1"))
        if r.returncode != 0:
            bad.append((name, "exit %d: %s %s" % (r.returncode, r.stdout[-300:], r.stderr[-600:])))
            continue
        try:
            _same(dev, host, [".bwt", ".sa", ".pac", ".ann", ".amb"])
        except AssertionError as e:
            bad.append((name, str(e)))
    for name, why in bad:   # in full: pytest abbreviates a long assertion message
        print("== %s: %s" % (name, why))
    assert not bad, "%d of %d texts differ or failed: %s" % (len(bad), len(cases), [b[0] for b in bad])
