"""-m gpu: the CUDA path (through the C ABI) against the oracle, read by read, bit-exact."""
import numpy as np
import pytest

import parity_util as pu
from kart_b200 import KartIndex, Mapper, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eco(built):
    idx = KartIndex(pu.default_prefix())
    return idx, pu.genome_of(idx), pu.default_prefix()


def test_pipeline_work_counters_c1_like(eco):
    """Seeding work on a small paired set equals the oracle's counted work (extension steps, seeds, LF steps)."""
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 2000, 150, 0.02, seed=21)
    reads = pu.interleave(r1, r2)
    orc = pu.Oracle(prefix)
    orc.counters(reset=True)
    for r in reads:
        orc.seeds(r.tobytes())
    c = orc.counters()
    m = pu.make_mapper(idx, paired=True)
    flat, off = Mapper.pack_reads(reads)
    m.map_chunk(flat, off, 1500)
    w = m.work()
    assert w["ext_steps"] == c["ext_steps"]
    assert w["seeds"] == c["locates"]
    assert w["lf_steps"] == c["lf_steps"]


@pytest.mark.parametrize("seed,err,kw", [(1, 0.02, {}), (2, 0.02, dict(indel=0.005, n_rate=0.003)), (3, 0.05, dict(indel=0.002))])
def test_paired_vs_oracle(eco, seed, err, kw):
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 6000, 150, err, seed=seed, **kw)
    m = pu.make_mapper(idx, paired=True)
    assert pu.compare_pairs(m, pu.Oracle(prefix), pu.interleave(r1, r2)) == 0


@pytest.mark.parametrize("full_sa", [False, True])
def test_paired_64bit_row_kernel(eco, monkeypatch, full_sa):
    """k_fm_seed<.., u64> and k_fm_seed_q<.., u64> (what a multi-Gbp index runs; E. coli defaults to the 32-bit row kernels) against the oracle."""
    idx, g, prefix = eco
    monkeypatch.setenv("KB_ROW64", "1")
    r1, r2, _ = synth.simulate(g, 6000, 150, 0.03, seed=12, indel=0.002, n_rate=0.002)
    m = pu.make_mapper(idx, paired=True, expand_sa=full_sa)
    assert pu.compare_pairs(m, pu.Oracle(prefix), pu.interleave(r1, r2)) == 0


def test_paired_full_sa_identical(eco):
    """Expanding the sampled SA on the device changes traffic, not results."""
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 4000, 150, 0.02, seed=9)
    m = pu.make_mapper(idx, expand_sa=True, paired=True)
    assert pu.compare_pairs(m, pu.Oracle(prefix), pu.interleave(r1, r2)) == 0
    assert m.work()["lf_steps"] == 0
    # one-row searches finish against the text (kb_unique_tail): same searches, same step count, fewer sectors
    m2 = pu.make_mapper(idx, expand_sa=False, paired=True)
    assert pu.compare_pairs(m2, pu.Oracle(prefix), pu.interleave(r1, r2)) == 0
    assert m.work()["ext_steps"] == m2.work()["ext_steps"] and m.work()["seeds"] == m2.work()["seeds"]
    assert m.work()["occ_blocks"] < m2.work()["occ_blocks"]


def test_single_end_high_error_vs_oracle(eco):
    idx, g, prefix = eco
    r1, _, _ = synth.simulate(g, 8000, 100, 0.08, seed=4, paired=False)
    m = pu.make_mapper(idx, paired=False)
    assert pu.compare_singles(m, pu.Oracle(prefix), r1) == 0


def test_est_distance_values(eco):
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 3000, 150, 0.02, seed=5)
    m = pu.make_mapper(idx, paired=True)
    orc = pu.Oracle(prefix)
    for est in (400, 520, 1500):
        assert pu.compare_pairs(m, orc, pu.interleave(r1, r2), est=est) == 0


def test_empty_and_ragged_chunks(eco):
    idx, g, prefix = eco
    m = pu.make_mapper(idx, paired=False)
    aln, pairs, cig = m.map_chunk(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(aln) == 0
    reads = [b"ACGT", b"N" * 40, g[0][1000:1013].tobytes(), g[0][5000:5250].tobytes(), b"acgtn" * 20, g[0][70000:70031].tobytes().lower()]
    assert pu.compare_singles(m, pu.Oracle(prefix), reads) == 0


def test_pacbio_vs_oracle(built):
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    r, _, _ = synth.simulate(g, 64, 3000, 0.15, seed=34, paired=False, indel=0.01)
    m = pu.make_mapper(idx, pacbio=True)
    assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX, pacbio=True), r) == 0


@pytest.mark.parametrize("pacbio", [False, True])
def test_partition_on_packed_words_gpu(built, pacbio):
    """The 8-mer partition of k_align_part on fragments full of short exact runs (part_pairs_packed: kb_match8 cells, run starts, runs across
    cell borders, both strands) and on fragments with an N / a lower-case base (literal id scan), against the oracle."""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    reads = pu.partition_stress_reads(g, n=30 if pacbio else 24, seed=7 if pacbio else 5)
    m = pu.make_mapper(idx, pacbio=pacbio)
    assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX, pacbio=pacbio), reads) == 0
    c = m.debug(9, np.uint32, 32)
    assert c[23] >= len(reads) // 3, c[23]   # the partition did run


@pytest.mark.parametrize("stack,raw", [("1", "1"), ("3", "4")])
def test_partition_buffers_spill_gpu(built, monkeypatch, stack, raw):
    """The CUDA k_align_part with tiny shared-memory limits for the work stack and the run list (HBM part of the stack, part_grow's second pass)."""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    monkeypatch.setenv("KB_PART_STACK", stack); monkeypatch.setenv("KB_PART_RAW", raw)
    r, _, _ = synth.simulate(g, 48, 3000, 0.15, seed=36, paired=False, indel=0.01)
    assert pu.compare_singles(pu.make_mapper(idx, pacbio=True), pu.Oracle(pu.MINI_PREFIX, pacbio=True), r) == 0
    assert pu.compare_singles(pu.make_mapper(idx), pu.Oracle(pu.MINI_PREFIX), pu.big_gap_reads(g)) == 0


def test_all_nw_size_classes(built, monkeypatch):
    """Every nw_alignment size class of the CUDA path (thread-per-problem register tiles, column tiles, warp wavefront) against the oracle."""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    reads = pu.big_gap_reads(g)
    seen = np.zeros(7, dtype=np.int64)
    # sparse column-tile classes (33..128) are solved by the warp kernel (KB_NW_WARP_BELOW), KB_NW_TMAX moves the thread limit:
    # cover every kernel on the same problems
    for pac, env in ((True, {}), (False, {}), (False, {"KB_NW_WARP_BELOW": "0"}), (True, {"KB_NW_WARP_BELOW": "0"}), (True, {"KB_NW_TMAX": "32"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = pu.make_mapper(idx, pacbio=pac)
        assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX, pacbio=pac), reads) == 0
        seen += m.debug(9, np.uint32, 32)[16:23]
        for k in env:
            monkeypatch.delenv(k)
    r, _, _ = synth.simulate(g, 3000, 100, 0.08, seed=35, paired=False, indel=0.003)
    for below in (None, "0"):
        if below:
            monkeypatch.setenv("KB_NW_WARP_BELOW", below)
        m = pu.make_mapper(idx, paired=False)
        assert pu.compare_singles(m, pu.Oracle(pu.MINI_PREFIX), r) == 0
        seen += m.debug(9, np.uint32, 32)[16:23]
    assert (seen > 0).all(), seen


def test_multi_contig_paired_with_rescue(built):
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    r1, r2, _ = synth.simulate(g, 4000, 150, 0.07, seed=31, indel=0.005, n_rate=0.003)
    m = pu.make_mapper(idx, paired=True)
    assert pu.compare_pairs(m, pu.Oracle(pu.MINI_PREFIX), pu.interleave(r1, r2)) == 0
    assert m.work()["rescues"] > 0


def test_packed_entry_point_gpu(eco, monkeypatch):
    """kb_map_chunk_packed (2-bit words + exceptions in, k_unpack on the device) against kb_map_chunk on the text and against the
    oracle: N / lower case / IUPAC characters, single batch and slot pipeline."""
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 20000, 150, 0.03, seed=51, indel=0.004, n_rate=0.004)
    reads = pu.interleave(r1, r2)
    reads[3::17, 20] = ord("a"); reads[4::29, 100] = ord("R"); reads[7::31, 0] = ord("n"); reads[9::37, 149] = ord("t")
    flat, off = Mapper.pack_reads(reads)
    est = np.full(20000, 1500, dtype=np.int32)
    m = pu.make_mapper(idx, expand_sa=True, paired=True)
    a0, p0, c0 = m.map_chunk(flat, off, est)
    pk = m.pack(flat, off)
    assert pk[0].n_exc > 1000

    def same(res):
        a1, p1, c1 = res
        for f in ("pos", "mate_pos", "kind", "flag", "chr", "mapq", "score", "sub_score", "tlen", "fwd", "cig_len"):
            assert np.array_equal(a0[f], a1[f]), f
        assert np.array_equal(p0, p1)
        i0 = np.repeat(a0["cig_off"].astype(np.int64), a0["cig_len"]) + (np.arange(int(a0["cig_len"].sum())) - np.repeat(np.cumsum(a0["cig_len"]) - a0["cig_len"], a0["cig_len"]))
        i1 = np.repeat(a1["cig_off"].astype(np.int64), a1["cig_len"]) + (np.arange(int(a1["cig_len"].sum())) - np.repeat(np.cumsum(a1["cig_len"]) - a1["cig_len"], a1["cig_len"]))
        assert np.array_equal(c0[i0], c1[i1])
    same(m.map_chunk(flat, off, est, packed=pk))
    monkeypatch.setenv("KB_PIPE_MIN_READS", "1000"); monkeypatch.setenv("KB_PIPE_SUB_READS", "9000")
    m2 = pu.make_mapper(idx, expand_sa=True, paired=True)
    same(m2.map_chunk(flat, off, est, packed=pk))
    assert m2.work()["launches"] > 3 * 19
    monkeypatch.delenv("KB_PIPE_MIN_READS"); monkeypatch.delenv("KB_PIPE_SUB_READS")
    monkeypatch.setenv("KART_TEST_PACKED", "1")
    assert pu.compare_pairs(pu.make_mapper(idx, paired=True), pu.Oracle(prefix), reads[:8000]) == 0


def test_heavy_candidate_items_gpu(built, monkeypatch):
    """k_cand_heavy (warp per item with a long seed list, cooperative sort) and the single-kernel path (KB_CAND_HEAVY=0) against the oracle"""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    r1, r2, _ = synth.simulate(g, 6000, 150, 0.02, seed=31, indel=0.002)
    r, _, _ = synth.simulate(g, 6000, 100, 0.05, seed=32, paired=False)
    orc = pu.Oracle(pu.MINI_PREFIX)
    for split in ("1", "0"):
        monkeypatch.setenv("KB_CAND_HEAVY", split)
        m = pu.make_mapper(idx, paired=True)
        assert pu.compare_pairs(m, orc, pu.interleave(r1, r2)) == 0
        assert (m.debug(9, np.uint32, 32)[15] > 50) == (split == "1")
        m = pu.make_mapper(idx, paired=False)
        assert pu.compare_singles(m, orc, r) == 0
        assert (m.debug(9, np.uint32, 32)[15] > 50) == (split == "1")


def test_chunks_in_flight_gpu(eco):
    """kb_map_chunk_begin / _end on the device: three chunks in flight (text and packed), ended out of order, equal the synchronous call"""
    idx, g, prefix = eco
    m = pu.make_mapper(idx, expand_sa=True, paired=True)
    chunks = []
    for k in range(3):
        r1, r2, _ = synth.simulate(g, 30000 + 5000 * k, 150, 0.03, seed=80 + k, indel=0.003)
        flat, off = Mapper.pack_reads(pu.interleave(r1, r2))
        chunks.append((flat, off, np.full(30000 + 5000 * k, 1500 - 400 * k, dtype=np.int32)))
    want = [m.map_chunk(*c) for c in chunks]
    hs = [m.map_chunk_begin(c[0], c[1], c[2], packed=(True if k == 1 else None)) for k, c in enumerate(chunks)]
    with pytest.raises(Exception):
        m.map_chunk_begin(*chunks[0])
    got = {k: m.map_chunk_end(hs[k]) for k in (2, 0, 1)}
    for k in range(3):
        a0, p0, c0 = want[k]; a1, p1, c1 = got[k]
        for f in ("pos", "mate_pos", "kind", "flag", "chr", "mapq", "score", "sub_score", "tlen", "fwd", "cig_len"):
            assert np.array_equal(a0[f], a1[f]), (k, f)
        assert np.array_equal(p0, p1) and len(c0) == len(c1)
        i0 = np.repeat(a0["cig_off"].astype(np.int64), a0["cig_len"]) + (np.arange(int(a0["cig_len"].sum())) - np.repeat(np.cumsum(a0["cig_len"]) - a0["cig_len"], a0["cig_len"]))
        i1 = np.repeat(a1["cig_off"].astype(np.int64), a1["cig_len"]) + (np.arange(int(a1["cig_len"].sum())) - np.repeat(np.cumsum(a1["cig_len"]) - a1["cig_len"], a1["cig_len"]))
        assert np.array_equal(c0[i0], c1[i1])


def test_rescue_fast_path_and_fallback(built, monkeypatch):
    """k_rescue_fast (warp per window, shared memory) takes the clean windows, k_rescue_win the rest; both against the oracle."""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    r1, r2, _ = synth.simulate(g, 6000, 150, 0.07, seed=41, indel=0.005)
    reads = pu.interleave(r1, r2)
    reads[5::40, 70] = ord("N")
    orc = pu.Oracle(pu.MINI_PREFIX)
    m = pu.make_mapper(idx, paired=True)
    assert pu.compare_pairs(m, orc, reads) == 0
    c = m.debug(9, np.uint32, 32)
    assert c[27] > 100 and c[29] < c[27] // 2, (c[27], c[29])
    for cap in ("0", "3"):   # the probe list of kb_rf_scan too short: positions beyond it are probed on the spot
        monkeypatch.setenv("KB_RF_CAND", cap)
        assert pu.compare_pairs(pu.make_mapper(idx, paired=True), orc, reads) == 0
    monkeypatch.delenv("KB_RF_CAND")
    monkeypatch.setenv("KB_RF_STRIDE", "1")   # every window position scanned instead of every third
    assert pu.compare_pairs(pu.make_mapper(idx, paired=True), orc, reads) == 0
    monkeypatch.delenv("KB_RF_STRIDE")
    monkeypatch.setenv("KB_RESCUE_FAST", "0")
    m = pu.make_mapper(idx, paired=True)
    assert pu.compare_pairs(m, orc, reads) == 0
    assert m.debug(9, np.uint32, 32)[29] == m.debug(9, np.uint32, 32)[27]


def test_rescue_sampled_scan_repeat_rich(built, tmp_path, monkeypatch):
    """k_rescue_fast scans every third window position (a run of >= 10 bases holds three consecutive 8-mers) and walks back to the
    run's start: same pairs as the oracle on a genome of micro-satellites and a diverged repeat family at 10 % read error, where
    thousands of windows are rescued and the mate's 8-mer chains are long; the every-position scan next to it."""
    import stage_util as su
    prefix, _ = su.text_index(str(tmp_path), [pu.repeat_rich_text()])
    idx = KartIndex(prefix)
    r1, r2, _ = synth.simulate(pu.genome_of(idx), 20000, 150, 0.10, seed=11, indel=0.01)
    reads = pu.interleave(r1, r2)
    orc = pu.Oracle(prefix)
    for stride, reuse in (("3", "1"), ("1", "1"), ("3", "0")):   # reuse: the mate's 8-mer index kept across consecutive windows of one mate
        monkeypatch.setenv("KB_RF_STRIDE", stride)
        monkeypatch.setenv("KB_RF_REUSE", reuse)
        m = pu.make_mapper(idx, paired=True)
        assert pu.compare_pairs(m, orc, reads) == 0
        c = m.debug(9, np.uint32, 32)
        assert m.work()["rescues"] > 4000 and c[27] - c[29] > 1500, (m.work()["rescues"], c[27], c[29])


def test_seeding_load_paths(built, monkeypatch):
    """The lane-queue seeding kernel with a read's packed words staged in shared memory or walked in HBM, and with or without the
    L1 no-allocate loads of Occ blocks / table / SA entries: the oracle's pairs every time (reads of 150 and of 250 bases: the
    latter are too long for the stage)."""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    orc = pu.Oracle(pu.MINI_PREFIX)
    sets = []
    for L, seed in ((150, 5), (250, 6)):
        r1, r2, _ = synth.simulate(g, 4000, L, 0.03, seed=seed, indel=0.002)
        reads = pu.interleave(r1, r2)
        reads[3::50, 40] = ord("N")
        sets.append(reads)
    counts = {}
    for stage, hint, fast in (("1", "1", "1"), ("0", "1", "1"), ("1", "0", "0"), ("0", "0", "1"), ("1", "1", "0")):   # fast: kb_unique_tail's one-strand path
        monkeypatch.setenv("KB_SEED_STAGE", stage)
        monkeypatch.setenv("KB_SEED_LD_HINT", hint)
        monkeypatch.setenv("KB_SEED_TAIL_FAST", fast)
        m = pu.make_mapper(idx, expand_sa=True, paired=True)
        for k, reads in enumerate(sets):
            assert pu.compare_pairs(m, orc, reads) == 0
            w = m.work()
            assert counts.setdefault(k, (w["ext_steps"], w["occ_blocks"])) == (w["ext_steps"], w["occ_blocks"])   # the work counters do not depend on the path


def test_segx_slab_sizes(built, monkeypatch):
    """k_segments hands out segment slots from a per-warp slab (KB_SEG_SLAB, 256 by default), falling back to the arena cursor when a warp
    runs out: no slab, a slab that is always too small and the default give the oracle's pairs."""
    idx = KartIndex(pu.MINI_PREFIX)
    g = pu.genome_of(idx)
    r1, r2, _ = synth.simulate(g, 6000, 150, 0.05, seed=23, indel=0.004)
    reads = pu.interleave(r1, r2)
    orc = pu.Oracle(pu.MINI_PREFIX)
    for slab in ("0", "8", "256"):
        monkeypatch.setenv("KB_SEG_SLAB", slab)
        assert pu.compare_pairs(pu.make_mapper(idx, paired=True), orc, reads) == 0


@pytest.mark.parametrize("plan", [{"KB_PIPE_SUB_READS": "65536"}, {"KB_PIPE_SUB_READS": "100000", "KB_PIPE_FIRST": "8192", "KB_PIPE_GROW": "150", "KB_PIPE_TAIL": "8192"}])
def test_pipelined_chunk_full_size_equals_single_batch(eco, monkeypatch, plan):
    """C2 at size: 400k reads through kb_map_chunk's slot pipeline (uniform sub-batches of 65536, or a ramped plan; cigar ranges
    copied back as sub-batches retire) and as one resident batch give the same records (the single-batch path is the one the
    oracle checks read by read above)."""
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 200000, 150, 0.02, seed=77, indel=0.001)
    flat, off = Mapper.pack_reads(pu.interleave(r1, r2))
    est = np.full(200000, 1500, dtype=np.int32)
    monkeypatch.setenv("KB_PIPE_MIN_READS", "2000000000")
    m0 = pu.make_mapper(idx, expand_sa=True, paired=True)
    a0, p0, c0 = m0.map_chunk(flat, off, est)
    monkeypatch.setenv("KB_PIPE_MIN_READS", "1000")
    for k, v in plan.items():
        monkeypatch.setenv(k, v)
    m1 = pu.make_mapper(idx, expand_sa=True, paired=True)
    a1, p1, c1 = m1.map_chunk(flat, off, est)
    assert m1.work()["launches"] >= 5 * 19
    for f in ("pos", "mate_pos", "kind", "flag", "chr", "mapq", "score", "sub_score", "tlen", "fwd", "cig_len"):
        assert np.array_equal(a0[f], a1[f]), f
    assert np.array_equal(p0, p1)
    assert len(c0) == len(c1)
    idx0 = np.repeat(a0["cig_off"].astype(np.int64), a0["cig_len"]) + (np.arange(int(a0["cig_len"].sum())) - np.repeat(np.cumsum(a0["cig_len"]) - a0["cig_len"], a0["cig_len"]))
    idx1 = np.repeat(a1["cig_off"].astype(np.int64), a1["cig_len"]) + (np.arange(int(a1["cig_len"].sum())) - np.repeat(np.cumsum(a1["cig_len"]) - a1["cig_len"], a1["cig_len"]))
    assert np.array_equal(c0[idx0], c1[idx1])
    assert (a0["score"] > 0).mean() > 0.99


import os            # noqa: E402
import subprocess    # noqa: E402

G = os.path.join(pu.ROOT, "tests", "golden")
KART = os.path.join(pu.ROOT, "kart_b200", "bin", "kart")


@pytest.mark.parametrize("tag,args", [("pe150", ["-f", "pe150_1.fq", "-f2", "pe150_2.fq"]), ("se100", ["-f", "se100.fq"]), ("pb3k", ["-pacbio", "-f", "pb3k.fq"])])
def test_cli_golden_sam(built, tmp_path, tag, args):
    """kart_b200/bin/kart (CUDA) writes the reference's SAM bytes for the committed golden inputs."""
    out = str(tmp_path / (tag + ".sam"))
    a = [os.path.join(G, x) if x.endswith(".fq") else x for x in args]
    subprocess.run([KART, "-silent", "-i", pu.MINI_PREFIX] + a + ["-o", out], check=True, stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(G, tag + ".sam"), "rb").read()


@pytest.mark.parametrize("tag,args", [("pe150", ["-f", "pe150_1.fq", "-f2", "pe150_2.fq"]), ("pb3k", ["-pacbio", "-f", "pb3k.fq"])])
def test_cli_golden_bam(built, tmp_path, tag, args):
    """-bo through the CUDA CLI: the reference's BAM bytes (htslib 1.5 record encoding and BGZF block policy)."""
    out = str(tmp_path / (tag + ".bam"))
    a = [os.path.join(G, x) if x.endswith(".fq") else x for x in args]
    subprocess.run([KART, "-silent", "-i", pu.MINI_PREFIX] + a + ["-bo", out], check=True, stdout=subprocess.DEVNULL)
    assert pu.bam_equal(out, os.path.join(G, tag + ".bam"))


@pytest.mark.parametrize("tag,genome,args", [("pe150m", "dup", ["-f", "pe150m_1.fq", "-f2", "pe150m_2.fq"]), ("se100m", "mini", ["-f", "se100.fq"]), ("pb3km", "mini", ["-pacbio", "-f", "pb3k.fq"])])
def test_cli_multihit_golden_sam(built, tmp_path, tag, genome, args):
    """-m through the CUDA CLI (k_finalize emits the further lines, kb_fetch_extra returns them): the reference's bytes."""
    out = str(tmp_path / (tag + ".sam"))
    a = [os.path.join(G, x) if x.endswith(".fq") else x for x in args]
    subprocess.run([KART, "-silent", "-m", "-i", os.path.join(G, genome, genome)] + a + ["-o", out], check=True, stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(G, tag + ".sam"), "rb").read()


@pytest.mark.skipif(not os.path.exists(pu.REF_KART), reason="needs oracle/_ref/kart (built in the build container, shipped with the snapshot)")
def test_cli_multihit_6000_pairs_vs_reference(built, tmp_path):
    """-m with the EstDistance recurrence active (>= 1000 counted pairs): identical to `kart -t 1 -m` except for the SAM flag of the
    lines whose report the reference never flags (uninitialised AlignmentReport_t::SamFlag, AlignmentCandidates.cpp:636; we print 0)."""
    prefix = os.path.join(G, "dup", "dup")
    g = pu.genome_of(KartIndex(prefix))
    f1, f2 = synth.make_reads(g, str(tmp_path / "mh"), 6000, 150, 0.01, seed=5, indel=0.002)
    ours, ref = str(tmp_path / "ours.sam"), str(tmp_path / "ref.sam")
    subprocess.run([KART, "-silent", "-m", "-i", prefix, "-f", f1, "-f2", f2, "-o", ours], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([pu.REF_KART, "-silent", "-t", "1", "-m", "-i", prefix, "-f", f1, "-f2", f2, "-o", ref], check=True, stdout=subprocess.DEVNULL)
    a, b = open(ours, "rb").read().splitlines(), open(ref, "rb").read().splitlines()
    assert len(a) == len(b) and len(a) > 12000 + 1000
    n_flag = 0
    for x, y in zip(a, b):
        if x != y:
            fx, fy = x.split(b"\t"), y.split(b"\t")
            assert fx[:1] + fx[2:] == fy[:1] + fy[2:] and fx[1] == b"0"
            n_flag += 1
    assert n_flag < 100


@pytest.mark.skipif(not (os.path.exists(pu.REF_KART) and pu.have_ecoli()), reason="needs oracle/_ref/kart and the E. coli index (built in the build container, shipped with the snapshot)")
def test_cli_100k_pairs_identical_to_reference_t1(built, tmp_path):
    """C2-shaped input through the real CLI: byte-identical to `kart -t 1` including the per-chunk EstDistance recurrence."""
    idx = KartIndex(pu.ECOLI_PREFIX)
    g = pu.genome_of(idx)
    f1, f2 = synth.make_reads(g, str(tmp_path / "r"), 100000, 150, 0.02, seed=1)
    ours, ref = str(tmp_path / "ours.sam"), str(tmp_path / "ref.sam")
    subprocess.run([KART, "-silent", "-i", pu.ECOLI_PREFIX, "-f", f1, "-f2", f2, "-o", ours, "--batch", "60000"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([pu.REF_KART, "-silent", "-t", "1", "-i", pu.ECOLI_PREFIX, "-f", f1, "-f2", f2, "-o", ref], check=True, stdout=subprocess.DEVNULL)
    assert open(ours, "rb").read() == open(ref, "rb").read()


SYN = os.path.join(pu.ROOT, "data", "_gen", "syn", "syn100")


@pytest.mark.skipif(not os.path.exists(SYN + ".bwt"), reason="100 Mbp synthetic index (scripts/make_syn_index.py 100 4) not built")
def test_c3_c4_shaped_on_100mbp_vs_oracle(built):
    """C3/C4-shaped reads (PE 2x150 @ 1 %, SE 100 @ 8 %) on a 100 Mbp, 4-contig genome with injected repeat families:
    33-bit-free but HBM-sized index, many rescues, multi-contig coordinates."""
    idx = KartIndex(SYN)
    g = pu.genome_of(idx)
    r1, r2, _ = synth.simulate(g, 6000, 150, 0.01, seed=2)
    m = pu.make_mapper(idx, paired=True)
    orc = pu.Oracle(SYN)
    assert pu.compare_pairs(m, orc, pu.interleave(r1, r2)) == 0
    assert m.work()["rescues"] > 50
    r, _, _ = synth.simulate(g, 6000, 100, 0.08, seed=3, paired=False)
    m = pu.make_mapper(idx, paired=False)
    assert pu.compare_singles(m, orc, r) == 0


def _n_gpus():
    try:
        import ctypes
        from kart_b200.binding import load_library
        return int(load_library().kb_device_count())
    except Exception:
        return 0


@pytest.mark.skipif(not (os.path.exists(pu.REF_KART) and pu.have_ecoli()), reason="needs oracle/_ref/kart and the E. coli index")
def test_cli_multi_gpu_identical_to_reference_t1(built, tmp_path):
    """The CLI's device worker pool on every visible GPU (batches of 40 000 reads, so that a box with 2+ GPUs brings them all up):
    byte-identical to `kart -t 1` including the EstDistance recurrence across devices. With one GPU this is the single-device path."""
    idx = KartIndex(pu.ECOLI_PREFIX)
    g = pu.genome_of(idx)
    f1, f2 = synth.make_reads(g, str(tmp_path / "r"), 160000, 150, 0.02, seed=3)
    ours, ref = str(tmp_path / "ours.sam"), str(tmp_path / "ref.sam")
    r = subprocess.run([KART, "-silent", "-t", "8", "-i", pu.ECOLI_PREFIX, "-f", f1, "-f2", f2, "-o", ours, "--batch", "40000"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, KART_B200_TRACE="1"))
    subprocess.run([pu.REF_KART, "-silent", "-t", "1", "-i", pu.ECOLI_PREFIX, "-f", f1, "-f2", f2, "-o", ref], check=True, stdout=subprocess.DEVNULL)
    assert open(ours, "rb").read() == open(ref, "rb").read()
    if _n_gpus() >= 2:
        assert "device 1 up" in r.stderr
    else:
        print("one visible GPU: the multi-device pool ran with a single worker")


@pytest.mark.skipif(_n_gpus() < 2, reason="kb_clone_index between devices needs two GPUs (run with gpurun --gpus 2)")
def test_clone_index_between_devices(eco):
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 20000, 150, 0.02, seed=62)
    m1 = pu.make_mapper(idx, expand_sa=True, paired=True)
    m2 = Mapper(device=1)
    m2.clone_index_from(m1)
    m2.set_params(paired=True)
    assert pu.compare_pairs(m2, pu.Oracle(prefix), pu.interleave(r1[:6000], r2[:6000])) == 0
