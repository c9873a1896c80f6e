"""CPU: `kart index` (kart_b200/host/index_build.cpp) writes the same five files, byte for byte, as the reference's bwt_index:
against the committed indexes of the golden genomes (made with the reference's tool, tests/golden/make_golden.py) and, where
the compiled reference is present, on a FASTA with CRLF line ends, header comments, lower case and runs of ambiguity codes."""
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu

G = os.path.join(pu.ROOT, "tests", "golden")
EXT = ("bwt", "sa", "pac", "ann", "amb")
BWT_INDEX = os.path.join(pu.ROOT, "oracle", "_ref", "bwt_index")


@pytest.fixture(scope="module")
def kart_cli(built):
    """The host code linked against the host-compiled device code (no GPU needed for `index`)."""
    exe = os.path.join(pu.ROOT, "tests", "emul", "kart_emul")
    host = os.path.join(pu.ROOT, "kart_b200", "host")
    src = [os.path.join(host, f) for f in os.listdir(host) if f.endswith(".cpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in src + [os.path.join(host, "kart_host.h")]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe] + src + ["-L" + os.path.dirname(exe), "-lkartb200_emul", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("genome", ["mini", "dup"])
def test_index_files_equal_the_reference_builders(kart_cli, tmp_path, genome, wide):
    """Both instantiations of the builder: 32-bit suffix positions, and the 64-bit ones a genome above 2.1 Gbp gets."""
    out = str(tmp_path / genome)
    subprocess.run([kart_cli, "index", os.path.join(G, genome, genome + ".fa"), out], check=True, stdout=subprocess.DEVNULL, env=dict(os.environ, KART_INDEX_64=wide))
    for e in EXT:
        assert open(out + "." + e, "rb").read() == open(os.path.join(G, genome, genome + "." + e), "rb").read(), e


@pytest.mark.skipif(not os.path.exists(BWT_INDEX), reason="needs oracle/_ref/bwt_index (the compiled reference)")
def test_index_odd_fasta_vs_reference(kart_cli, tmp_path):
    rng = np.random.default_rng(7)

    def seq(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    fa = str(tmp_path / "odd.fa")
    with open(fa, "w", newline="") as f:
        s = seq(5000)
        s = s[:100] + "N" * 37 + s[137:900] + "nnnNNRY" + s[907:2000].lower() + s[2000:]
        f.write(">c1 first contig  with comment\r\n" + "".join(s[i:i + 60] + "\r\n" for i in range(0, len(s), 60)))
        s = seq(3001)
        s = "N" + s[1:1500] + "N" * 5 + "K" + s[1506:] + "N"
        f.write(">c2\n" + "".join(s[i:i + 70] + "\n" for i in range(0, len(s), 70)))
        f.write("\n>c3\tdesc tab\n" + seq(777) + "\n")
        f.write(">c4 poly\n" + "A" * 400 + "ACGT" * 50 + "T" * 300 + "\n")   # long runs: deep suffix comparisons, suffixes running into the end of the text
    subprocess.run([BWT_INDEX, fa, str(tmp_path / "ref")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([kart_cli, "index", fa, str(tmp_path / "ours")], check=True, stdout=subprocess.DEVNULL)
    for e in EXT:
        assert open(str(tmp_path / "ours") + "." + e, "rb").read() == open(str(tmp_path / "ref") + "." + e, "rb").read(), e


def test_index_usage(kart_cli):
    r = subprocess.run([kart_cli, "index", "only_one_argument"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage:" in r.stderr and "index ref.fa prefix" in r.stderr
