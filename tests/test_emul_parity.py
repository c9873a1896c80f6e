"""CPU: the per-thread bodies of the CUDA kernels, compiled for the host (tests/emul, test infrastructure only), and the
whole C-ABI / host flow around them, against the oracle. The real kernels are checked by the -m gpu tests."""
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu
from kart_b200 import KartIndex, Mapper, synth

G = os.path.join(pu.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def mini(built):
    idx = KartIndex(pu.MINI_PREFIX)
    return idx, pu.genome_of(idx)


@pytest.mark.parametrize("row64", [False, True])
def test_paired_multi_contig(mini, monkeypatch, row64):
    """Seeding runs with 32-bit BWT row numbers when the text allows it and with 64-bit ones otherwise (KB_ROW64 forces them)."""
    idx, g = mini
    if row64:
        monkeypatch.setenv("KB_ROW64", "1")
    r1, r2, _ = synth.simulate(g, 1500, 150, 0.07, seed=31, indel=0.005, n_rate=0.003)
    m = pu.make_mapper(idx, emul=True, paired=True)
    assert pu.compare_pairs(m, pu.Oracle(pu.MINI_PREFIX), pu.interleave(r1, r2)) == 0
    assert m.work()["rescues"] > 0


@pytest.mark.parametrize("row64", [False, True])
def test_paired_small_est_and_full_sa(mini, monkeypatch, row64):
    idx, g = mini
    if row64:
        monkeypatch.setenv("KB_ROW64", "1")
    r1, r2, _ = synth.simulate(g, 600, 150, 0.02, seed=32)
    m = pu.make_mapper(idx, emul=True, expand_sa=True, paired=True)
    assert pu.compare_pairs(m, pu.Oracle(pu.MINI_PREFIX), pu.interleave(r1, r2), est=470) == 0
    # with the full SA a search that is down to one row finishes by comparing against the text (kb_unique_tail): same step count
    m2 = pu.make_mapper(idx, emul=True, expand_sa=False, paired=True)
    assert pu.compare_pairs(m2, pu.Oracle(pu.MINI_PREFIX), pu.interleave(r1, r2), est=470) == 0
    assert m.work()["ext_steps"] == m2.work()["ext_steps"] and m.work()["seeds"] == m2.work()["seeds"]
    assert m.work()["occ_blocks"] < m2.work()["occ_blocks"]


def test_single_end_high_error(mini):
    idx, g = mini
    r, _, _ = synth.simulate(g, 1500, 100, 0.08, seed=33, paired=False, indel=0.003)
    m = pu.make_mapper(idx, emul=True, paired=False)
    assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX), r) == 0


def test_pacbio_long_reads(mini):
    idx, g = mini
    r, _, _ = synth.simulate(g, 16, 3000, 0.15, seed=34, paired=False, indel=0.01)
    m = pu.make_mapper(idx, emul=True, pacbio=True)
    assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX, pacbio=True), r) == 0


@pytest.mark.parametrize("pacbio", [False, True])
def test_partition_on_packed_words(mini, pacbio):
    """The 8-mer partition of k_align_part on fragments full of short exact runs (part_pairs_packed: kb_match8 cells, run starts, runs across
    cell borders, both strands) and on fragments with an N / a lower-case base (literal id scan), against the oracle."""
    idx, g = mini
    reads = pu.partition_stress_reads(g, n=30 if pacbio else 24, seed=7 if pacbio else 5)
    m = pu.make_mapper(idx, emul=True, pacbio=pacbio)
    assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX, pacbio=pacbio), reads) == 0
    c = m.debug(9, np.uint32, 32)
    assert c[23] >= len(reads) // 3, c[23]   # the partition did run


@pytest.mark.parametrize("stack,raw", [("1", "1"), ("3", "4")])
def test_partition_buffers_spill(mini, monkeypatch, stack, raw):
    """k_align_part keeps the first entries of a job's work stack and a short exact-match run list in the warp's pool; with tiny limits every
    partition goes through the HBM part of the stack and through part_grow()'s second pass over the full-size list. Same reads, same answers."""
    idx, g = mini
    monkeypatch.setenv("KB_PART_STACK", stack); monkeypatch.setenv("KB_PART_RAW", raw)
    r, _, _ = synth.simulate(g, 12, 3000, 0.15, seed=36, paired=False, indel=0.01)
    assert pu.compare_singles(pu.make_mapper(idx, emul=True, pacbio=True), pu.Oracle(pu.MINI_PREFIX, pacbio=True), r) == 0
    reads = pu.big_gap_reads(g)
    assert pu.compare_singles(pu.make_mapper(idx, emul=True), pu.Oracle(pu.MINI_PREFIX), reads) == 0


def test_all_nw_size_classes(mini, monkeypatch):
    """Every nw_alignment size class (register tiles <= 8/16/24/32, column tiles <= 64/128, warp wavefront) against the oracle."""
    idx, g = mini
    reads = pu.big_gap_reads(g)
    seen = np.zeros(7, dtype=np.int64)
    # sparse column-tile classes (33..128) are solved by the warp kernel (KB_NW_WARP_BELOW), KB_NW_TMAX moves the thread limit:
    # cover every kernel on the same problems
    for pac, env in ((True, {}), (False, {}), (False, {"KB_NW_WARP_BELOW": "0"}), (True, {"KB_NW_WARP_BELOW": "0"}), (True, {"KB_NW_TMAX": "32"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = pu.make_mapper(idx, emul=True, pacbio=pac)
        assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX, pacbio=pac), reads) == 0
        seen += m.debug(9, np.uint32, 32)[16:23]
        for k in env:
            monkeypatch.delenv(k)
    r, _, _ = synth.simulate(g, 300, 100, 0.08, seed=35, paired=False, indel=0.003)
    for below in (None, "0"):
        if below:
            monkeypatch.setenv("KB_NW_WARP_BELOW", below)
        m = pu.make_mapper(idx, emul=True, paired=False)
        assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX), r) == 0
        seen += m.debug(9, np.uint32, 32)[16:23]
    assert (seen > 0).all(), seen


@pytest.mark.parametrize("full_sa", [False, True])
def test_edge_reads(mini, full_sa):
    """Degenerate reads; with the full SA they go through the lane-queue seeding and its text comparison (kb_unique_tail)."""
    idx, g = mini
    m = pu.make_mapper(idx, emul=True, paired=False, expand_sa=full_sa)
    reads = [b"A", b"ACGTACGTACGTAC", b"N" * 60, g[0][:150].tobytes(), g[2][-150:].tobytes(), g[1][100:113].tobytes(),
             (g[0][3000:3075].tobytes() + g[1][500:575].tobytes()), g[0][200:350].tobytes().lower(), b"ACGTRYKM" * 15,
             g[0][-80:].tobytes() + g[1][:70].tobytes(), synth.revcomp_bytes(g[2][-100:]).tobytes() + b"ACGT" * 5,
             g[1][1000:1100].tobytes() + b"N" + g[1][1101:1200].tobytes()]   # contig junction, end of the text on the reverse strand, N inside a unique match
    assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX), reads) == 0
    aln, pairs, cig = m.map_chunk(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(aln) == 0


@pytest.fixture(scope="module")
def kart_emul(built):
    exe = os.path.join(pu.ROOT, "tests", "emul", "kart_emul")
    src = [os.path.join(pu.ROOT, "kart_b200", "host", f) for f in os.listdir(os.path.join(pu.ROOT, "kart_b200", "host")) if f.endswith(".cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe] + src + ["-L" + os.path.dirname(exe), "-lkartb200_emul", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("tag,args", [("pe150", ["-f", "pe150_1.fq", "-f2", "pe150_2.fq"]), ("se100", ["-f", "se100.fq"]), ("pb3k", ["-pacbio", "-f", "pb3k.fq"])])
def test_cli_sam_is_byte_identical_to_reference(kart_emul, tmp_path, tag, args):
    """The host C++ (CLI, input parsing, chunk recurrence, SAM text) over the emulated device code reproduces the reference's SAM bytes."""
    out = str(tmp_path / (tag + ".sam"))
    a = [os.path.join(G, x) if x.endswith(".fq") else x for x in args]
    subprocess.run([kart_emul, "-silent", "-t", "2", "-i", pu.MINI_PREFIX] + a + ["-o", out, "--batch", "400"], check=True, stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(G, tag + ".sam"), "rb").read()


@pytest.mark.parametrize("tag,args", [("pe150", ["-f", "pe150_1.fq", "-f2", "pe150_2.fq"]), ("se100", ["-f", "se100.fq"]), ("pb3k", ["-pacbio", "-f", "pb3k.fq"])])
def test_cli_bam_is_byte_identical_to_reference(kart_emul, tmp_path, tag, args):
    """-bo: the freshly written BAM encoder + BGZF blocker (kart_b200/host/bam_writer.cpp, parallel deflate) reproduces the
    bytes of the reference linked against its vendored htslib 1.5 (goldens: tests/golden/*.bam, made by make_golden.py bam)."""
    out = str(tmp_path / (tag + ".bam"))
    a = [os.path.join(G, x) if x.endswith(".fq") else x for x in args]
    subprocess.run([kart_emul, "-silent", "-t", "3", "-i", pu.MINI_PREFIX] + a + ["-bo", out, "--batch", "400"], check=True, stdout=subprocess.DEVNULL)
    assert pu.bam_equal(out, os.path.join(G, tag + ".bam"))


MH_CASES = [("pe150m", "dup", ["-f", "pe150m_1.fq", "-f2", "pe150m_2.fq"]), ("se100m", "mini", ["-f", "se100.fq"]), ("pb3km", "mini", ["-pacbio", "-f", "pb3k.fq"])]


@pytest.mark.parametrize("tag,genome,args", MH_CASES)
def test_cli_multihit_sam_is_byte_identical_to_reference(kart_emul, tmp_path, tag, genome, args):
    """-m (bMultiHit): one line per report from iBestAlnCanIdx on (Mapping.cpp:194-223,242-263,289-308). dup/ has exact repeats, so
    whole pairs tie (198 extra lines in pe150m.sam); goldens from `kart -t 1 -m` (make_golden.py multihit)."""
    out = str(tmp_path / (tag + ".sam"))
    a = [os.path.join(G, x) if x.endswith(".fq") else x for x in args]
    subprocess.run([kart_emul, "-silent", "-t", "2", "-m", "-i", os.path.join(G, genome, genome)] + a + ["-o", out, "--batch", "400"], check=True, stdout=subprocess.DEVNULL)
    gold = open(os.path.join(G, tag + ".sam"), "rb").read()
    assert open(out, "rb").read() == gold
    if tag == "pe150m":
        assert len(gold.splitlines()) > 2 * 300 + 3 + 100   # the case does exercise extra lines
        bam = str(tmp_path / "pe150m.bam")
        subprocess.run([kart_emul, "-silent", "-t", "3", "-m", "-i", os.path.join(G, genome, genome)] + a + ["-bo", bam, "--batch", "400"], check=True, stdout=subprocess.DEVNULL)
        assert pu.bam_equal(bam, os.path.join(G, "pe150m.bam"))


@pytest.mark.skipif(not os.path.exists(pu.REF_KART), reason="needs oracle/_ref/kart")
def test_cli_multihit_with_est_recurrence_vs_reference(kart_emul, tmp_path):
    """-m past 1000 counted pairs (settle_est replaces re-mapped pairs together with their extra lines): identical to `kart -t 1 -m`
    except for the flag of lines whose report the reference never flags (uninitialised SamFlag, AlignmentCandidates.cpp:636; we print 0)."""
    from kart_b200 import KartIndex, synth
    prefix = os.path.join(G, "dup", "dup")
    g = pu.genome_of(KartIndex(prefix))
    f1, f2 = synth.make_reads(g, str(tmp_path / "mh"), 3000, 150, 0.01, seed=5, indel=0.002)
    ours, ref = str(tmp_path / "ours.sam"), str(tmp_path / "ref.sam")
    subprocess.run([kart_emul, "-silent", "-t", "2", "-m", "-i", prefix, "-f", f1, "-f2", f2, "-o", ours], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([pu.REF_KART, "-silent", "-t", "1", "-m", "-i", prefix, "-f", f1, "-f2", f2, "-o", ref], check=True, stdout=subprocess.DEVNULL)
    a, b = open(ours, "rb").read().splitlines(), open(ref, "rb").read().splitlines()
    assert len(a) == len(b) and len(a) > 6000 + 500
    for x, y in zip(a, b):
        if x != y:
            fx, fy = x.split(b"\t"), y.split(b"\t")
            assert fx[:1] + fx[2:] == fy[:1] + fy[2:] and fx[1] == b"0"


KART_HTS = os.path.join(pu.ROOT, "oracle", "_ref", "kart_hts")


@pytest.mark.skipif(not os.path.exists(KART_HTS), reason="oracle/_ref/kart_hts (reference + its htslib) not built")
def test_cli_bam_fasta_unmapped_and_odd_characters(kart_emul, tmp_path):
    """FASTA input (QUAL 0xff), unmappable reads (bin 4680, flag |= 4), lower-case and IUPAC characters (4-bit codes), long names."""
    rng = np.random.default_rng(5)
    lines = open(os.path.join(G, "se100.fq")).read().split("\n")
    fa, fq = [], []
    for i in range(0, len(lines) - 1, 4):
        s = lines[i + 1]
        if i % 40 == 0:
            s = "".join("ACGT"[k] for k in rng.integers(0, 4, size=100))
        if i % 28 == 0:
            s = s[:10] + "nRYK" + s[14:].lower()
        name = lines[i][1:] + ("x" * 260 if i == 8 else "")
        fa += [">" + name, s]
        fq += ["@" + name + " extra words", s, "+", "".join(chr(33 + int(q)) for q in rng.integers(0, 42, size=len(s)))]
    for ext, rec in (("fa", fa), ("fq", fq)):
        f = str(tmp_path / ("in." + ext))
        open(f, "w").write("\n".join(rec) + "\n")
        ours, ref = str(tmp_path / ("o_" + ext + ".bam")), str(tmp_path / ("r_" + ext + ".bam"))
        subprocess.run([kart_emul, "-silent", "-t", "2", "-i", pu.MINI_PREFIX, "-f", f, "-bo", ours], check=True, stdout=subprocess.DEVNULL)
        subprocess.run([KART_HTS, "-silent", "-t", "1", "-i", pu.MINI_PREFIX, "-f", f, "-bo", ref], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert open(ours, "rb").read() == open(ref, "rb").read()


def test_cli_interleaved_and_gz_inputs(kart_emul, tmp_path):
    import gzip
    a = open(os.path.join(G, "pe150_1.fq"), "rb").read().split(b"\n")
    b = open(os.path.join(G, "pe150_2.fq"), "rb").read().split(b"\n")
    inter = []
    for i in range(0, len(a) - 1, 4):
        inter += a[i:i + 4] + b[i:i + 4]
    p = tmp_path / "inter.fq.gz"
    with gzip.open(p, "wb") as fh:
        fh.write(b"\n".join(inter) + b"\n")
    out = str(tmp_path / "i.sam")
    subprocess.run([kart_emul, "-silent", "-i", pu.MINI_PREFIX, "-p", "-f", str(p), "-o", out], check=True, stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(G, "pe150.sam"), "rb").read()


@pytest.mark.parametrize("flag", ["--full-sa", "--sampled-sa"])
@pytest.mark.parametrize("tag,args", [("pe150", ["-f", "pe150_1.fq", "-f2", "pe150_2.fq"]), ("se100", ["-f", "se100.fq"]), ("pb3k", ["-pacbio", "-f", "pb3k.fq"])])
def test_cli_sa_modes_give_the_same_sam(kart_emul, tmp_path, tag, args, flag):
    """--full-sa (lane-queue seeding that finishes one-row searches against the text, one-load locates) and --sampled-sa
    (every base walked, LF-walk locates) both reproduce the reference's SAM."""
    a = [os.path.join(G, x) if x.endswith(".fq") else x for x in args]
    out = str(tmp_path / (tag + ".sam"))
    subprocess.run([kart_emul, "-silent", "-t", "2", "-i", pu.MINI_PREFIX] + a + ["-o", out, flag], check=True, stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(G, tag + ".sam"), "rb").read()


def test_cli_argument_errors(kart_emul):
    r = subprocess.run([kart_emul, "-i", pu.MINI_PREFIX, "-zzz"], capture_output=True, text=True)
    assert r.returncode == 1 and "Error! Unknown parameter: -zzz" in r.stdout
    r = subprocess.run([kart_emul, "-i", pu.MINI_PREFIX], capture_output=True, text=True)
    assert r.returncode == 1 and "Please specify a valid read input" in r.stdout
    r = subprocess.run([kart_emul, "-v"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("kart v2.5.6")


@pytest.mark.parametrize("plan", [{}, {"KB_PIPE_FIRST": "60", "KB_PIPE_GROW": "150", "KB_PIPE_TAIL": "80"}])
def test_pipelined_chunk_equals_single_batch(mini, monkeypatch, plan):
    """kb_map_chunk streams large chunks through the slots (uniform or ramped sub-batches, chunk-wide cigar arena copied back
    range by range): same records as one batch."""
    idx, g = mini
    r1, r2, _ = synth.simulate(g, 700, 150, 0.04, seed=41, indel=0.004, n_rate=0.002)
    reads = pu.interleave(r1, r2)
    flat, off = Mapper.pack_reads(reads)
    est = np.full(len(reads) // 2, 1500, dtype=np.int32)
    est[::7] = 480
    m0 = pu.make_mapper(idx, emul=True, paired=True)
    a0, p0, c0 = m0.map_chunk(flat, off, est)
    monkeypatch.setenv("KB_PIPE_MIN_READS", "64")
    monkeypatch.setenv("KB_PIPE_SUB_READS", "250")
    for k, v in plan.items():
        monkeypatch.setenv(k, v)
    m1 = pu.make_mapper(idx, emul=True, paired=True)
    a1, p1, c1 = m1.map_chunk(flat, off, est)
    assert m1.work()["launches"] > 3 * m0.work()["launches"]          # really went through several sub-batches
    for f in ("pos", "mate_pos", "kind", "flag", "chr", "mapq", "score", "sub_score", "tlen", "fwd", "cig_len"):
        assert np.array_equal(a0[f], a1[f]), f
    assert np.array_equal(p0, p1)
    for i in range(len(a0)):
        assert np.array_equal(c0[a0["cig_off"][i]:a0["cig_off"][i] + a0["cig_len"][i]], c1[a1["cig_off"][i]:a1["cig_off"][i] + a1["cig_len"][i]])
    assert m0.work()["ext_steps"] == m1.work()["ext_steps"] and m0.work()["nw_cells"] == m1.work()["nw_cells"]


def test_rescue_fast_path_and_fallback(mini, monkeypatch):
    """Rescue windows go through the warp-per-window fast path (kb_rf_*) unless the window leaves the text or the mate holds a
    character that is no base; either route gives the oracle's pairs."""
    idx, g = mini
    r1, r2, _ = synth.simulate(g, 1200, 150, 0.07, seed=41, indel=0.005)
    reads = pu.interleave(r1, r2)
    reads[5::40, 70] = ord("N")   # a few dirty mates: slow list
    orc = pu.Oracle(pu.MINI_PREFIX)
    m = pu.make_mapper(idx, emul=True, paired=True)
    assert pu.compare_pairs(m, orc, reads) == 0
    c = m.debug(9, np.uint32, 32)
    assert c[27] > 20 and c[29] < c[27] // 2, (c[27], c[29])
    for cap in ("0", "3"):   # the probe list of kb_rf_scan too short: positions beyond it are probed on the spot
        monkeypatch.setenv("KB_RF_CAND", cap)
        assert pu.compare_pairs(pu.make_mapper(idx, emul=True, paired=True), orc, reads) == 0
    monkeypatch.delenv("KB_RF_CAND")
    monkeypatch.setenv("KB_RF_STRIDE", "1")   # every window position scanned instead of every third
    assert pu.compare_pairs(pu.make_mapper(idx, emul=True, paired=True), orc, reads) == 0
    monkeypatch.delenv("KB_RF_STRIDE")
    monkeypatch.setenv("KB_RESCUE_FAST", "0")
    m = pu.make_mapper(idx, emul=True, paired=True)
    assert pu.compare_pairs(m, orc, reads) == 0
    c = m.debug(9, np.uint32, 32)
    assert c[29] == c[27]


def test_rescue_sampled_scan_repeat_rich(mini, tmp_path, monkeypatch):
    """k_rescue_fast scans every third window position (a run of >= 10 bases holds three consecutive 8-mers) and walks back to the
    run's start: same pairs as the oracle on a genome of micro-satellites and a diverged repeat family at 10 % read error, where
    thousands of windows are rescued and the mate's 8-mer chains are long; the every-position scan next to it."""
    import stage_util as su
    prefix, _ = su.text_index(str(tmp_path), [pu.repeat_rich_text()])
    idx = KartIndex(prefix)
    r1, r2, _ = synth.simulate(pu.genome_of(idx), 3000, 150, 0.10, seed=11, indel=0.01)
    reads = pu.interleave(r1, r2)
    orc = pu.Oracle(prefix)
    for stride, reuse in (("3", "1"), ("1", "1"), ("3", "0")):   # reuse: the mate's 8-mer index kept across consecutive windows of one mate
        monkeypatch.setenv("KB_RF_STRIDE", stride)
        monkeypatch.setenv("KB_RF_REUSE", reuse)
        m = pu.make_mapper(idx, emul=True, paired=True)
        assert pu.compare_pairs(m, orc, reads) == 0
        c = m.debug(9, np.uint32, 32)
        assert m.work()["rescues"] > 500 and c[27] - c[29] > 200, (m.work()["rescues"], c[27], c[29])


def test_seeding_load_paths(mini, monkeypatch):
    """The lane-queue seeding kernel with a read's packed words staged in shared memory or walked in HBM, and with or without the
    L1 no-allocate loads of Occ blocks / table / SA entries: the oracle's pairs every time (reads of 150 and of 250 bases: the
    latter are too long for the stage)."""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    orc = pu.Oracle(pu.MINI_PREFIX)
    sets = []
    for L, seed in ((150, 5), (250, 6)):
        r1, r2, _ = synth.simulate(g, 300, L, 0.03, seed=seed, indel=0.002)
        reads = pu.interleave(r1, r2)
        reads[3::50, 40] = ord("N")
        sets.append(reads)
    counts = {}
    for stage, hint, fast in (("1", "1", "1"), ("0", "1", "1"), ("1", "0", "0"), ("0", "0", "1"), ("1", "1", "0")):   # fast: kb_unique_tail's one-strand path
        monkeypatch.setenv("KB_SEED_STAGE", stage)
        monkeypatch.setenv("KB_SEED_LD_HINT", hint)
        monkeypatch.setenv("KB_SEED_TAIL_FAST", fast)
        m = pu.make_mapper(idx, emul=True, expand_sa=True, paired=True)
        for k, reads in enumerate(sets):
            assert pu.compare_pairs(m, orc, reads) == 0
            w = m.work()
            assert counts.setdefault(k, (w["ext_steps"], w["occ_blocks"])) == (w["ext_steps"], w["occ_blocks"])   # the work counters do not depend on the path


def _same_results(a, b, paired=True):
    (a0, p0, c0), (a1, p1, c1) = a, b
    for f in ("pos", "mate_pos", "kind", "flag", "chr", "mapq", "score", "sub_score", "tlen", "fwd", "cig_len"):
        assert np.array_equal(a0[f], a1[f]), f
    assert not paired or np.array_equal(p0, p1)
    for x, y in zip(a0, a1):
        assert np.array_equal(c0[int(x["cig_off"]):int(x["cig_off"]) + int(x["cig_len"])], c1[int(y["cig_off"]):int(y["cig_off"]) + int(y["cig_len"])])


def test_packed_entry_point(mini, monkeypatch):
    """kb_map_chunk_packed (2-bit words + exception list over PCIe, characters rebuilt by k_unpack) == kb_map_chunk on the text,
    including N, lower case and IUPAC characters and ragged lengths; single batch and through the slot pipeline; vs the oracle."""
    idx, g = mini
    r1, r2, _ = synth.simulate(g, 1500, 150, 0.05, seed=51, indel=0.004, n_rate=0.004)
    reads = pu.interleave(r1, r2)
    reads[3::17, 20] = ord("a"); reads[4::29, 100] = ord("R"); reads[7::31, 0] = ord("n"); reads[9::37, 149] = ord("t")
    flat, off = Mapper.pack_reads(reads)
    est = np.full(1500, 1500, dtype=np.int32)
    m = pu.make_mapper(idx, emul=True, paired=True)
    base = m.map_chunk(flat, off, est)
    pk = m.pack(flat, off, threads=3)
    assert pk[0].n_exc > 100
    _same_results(base, m.map_chunk(flat, off, est, packed=pk))
    monkeypatch.setenv("KART_TEST_PACKED", "1")
    assert pu.compare_pairs(m, pu.Oracle(pu.MINI_PREFIX), reads) == 0
    monkeypatch.delenv("KART_TEST_PACKED")
    monkeypatch.setenv("KB_PIPE_MIN_READS", "500"); monkeypatch.setenv("KB_PIPE_SUB_READS", "700")
    m2 = pu.make_mapper(idx, emul=True, paired=True)
    _same_results(base, m2.map_chunk(flat, off, est, packed=pk))
    assert m2.work()["launches"] > 3 * 19
    # ragged single-end reads
    rag = [b"ACGT", b"N" * 40, g[0][1000:1013].tobytes(), g[0][5000:5250].tobytes(), b"acgtn" * 20, g[0][70000:70031].tobytes().lower(), g[1][300:364].tobytes()]
    m3 = pu.make_mapper(idx, emul=True, paired=False)
    f3, o3 = Mapper.pack_reads(rag)
    _same_results(m3.map_chunk(f3, o3), m3.map_chunk(f3, o3, packed=True), paired=False)


def test_cli_multi_device_worker_pool(kart_emul, mini, tmp_path):
    """The host's device worker pool (one thread + context per device, batches settled in input order, index cloned from device 0)
    over three emulated devices: 6 batches of 4000 reads, byte-identical to the oracle's `kart -t 1` replay whatever the interleaving,
    and all three devices took part."""
    import ctypes as C
    idx, g = mini
    f1, f2 = synth.make_reads(g, str(tmp_path / "md"), 11000, 150, 0.02, seed=71, indel=0.002)
    out, exp = str(tmp_path / "md.sam"), str(tmp_path / "exp.sam")
    o = pu.Oracle(pu.MINI_PREFIX)
    o.lib.kor_map_files.restype = C.c_long
    assert o.lib.kor_map_files(f1.encode(), f2.encode(), 1, exp.encode(), 4) == 22000
    env = dict(os.environ, KB_EMUL_DEVICES="3", KART_B200_TRACE="1")
    r = subprocess.run([kart_emul, "-silent", "-t", "2", "--gpus", "3", "-i", pu.MINI_PREFIX, "-f", f1, "-f2", f2, "-o", out, "--batch", "4000"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
    assert open(out, "rb").read() == open(exp, "rb").read()
    assert "device 1 up" in r.stderr and "device 2 up" in r.stderr
    assert len({ln.split()[2] for ln in r.stderr.splitlines() if ln.startswith("[kart trace] gpu")}) >= 2


def test_clone_index(mini):
    """kb_clone_index: a second context mapped from the first one's device arrays gives the same records"""
    idx, g = mini
    r1, r2, _ = synth.simulate(g, 400, 150, 0.03, seed=61)
    m1 = pu.make_mapper(idx, emul=True, expand_sa=True, paired=True)
    m2 = Mapper(lib_path=pu.EMUL_LIB)
    m2.clone_index_from(m1)
    m2.set_params(paired=True)
    assert pu.compare_pairs(m2, pu.Oracle(pu.MINI_PREFIX), pu.interleave(r1, r2)) == 0


def test_chunks_in_flight(mini):
    """kb_map_chunk_begin / _end: three chunks in flight (text and packed), ended out of order, equal the synchronous call"""
    idx, g = mini
    m = pu.make_mapper(idx, emul=True, paired=True)
    chunks = []
    for k in range(3):
        r1, r2, _ = synth.simulate(g, 300 + 50 * k, 150, 0.04, seed=80 + k, indel=0.003)
        flat, off = Mapper.pack_reads(pu.interleave(r1, r2))
        chunks.append((flat, off, np.full(300 + 50 * k, 1500 - 400 * k, dtype=np.int32)))
    want = [m.map_chunk(*c) for c in chunks]
    hs = [m.map_chunk_begin(c[0], c[1], c[2], packed=(True if k == 1 else None)) for k, c in enumerate(chunks)]
    with pytest.raises(Exception):
        m.map_chunk_begin(*chunks[0])          # a fourth chunk: every slot is taken
    with pytest.raises(Exception):
        m.map_chunk(*chunks[0])                # the synchronous call is refused while chunks are in flight
    got = {k: m.map_chunk_end(hs[k]) for k in (2, 0, 1)}
    for k in range(3):
        _same_results(want[k], got[k])
    _same_results(want[0], m.map_chunk(*chunks[0]))


def test_heavy_candidate_items(mini, monkeypatch):
    """Reads with more than KB_CAND_HEAVY seeds (repeat copies) are sorted by a warp in k_cand_heavy; with KB_CAND_HEAVY=0 every item stays
    in k_cand_pair. Both against the oracle, paired and single-end."""
    idx, g = mini
    r1, r2, _ = synth.simulate(g, 1500, 150, 0.02, seed=31, indel=0.002)
    r, _, _ = synth.simulate(g, 1500, 100, 0.05, seed=32, paired=False)
    orc = pu.Oracle(pu.MINI_PREFIX)
    for split in ("1", "0"):
        monkeypatch.setenv("KB_CAND_HEAVY", split)
        m = pu.make_mapper(idx, emul=True, paired=True)
        assert pu.compare_pairs(m, orc, pu.interleave(r1, r2)) == 0
        assert (m.debug(9, np.uint32, 32)[15] > 10) == (split == "1")
        m = pu.make_mapper(idx, emul=True, paired=False)
        assert pu.compare_singles(m, orc, r) == 0
        assert (m.debug(9, np.uint32, 32)[15] > 10) == (split == "1")
