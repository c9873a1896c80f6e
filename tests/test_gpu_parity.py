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


def test_paired_full_sa_identical(eco):
    """Expanding the sampled SA on the device changes traffic, not results."""
    idx, g, prefix = eco
    r1, r2, _ = synth.simulate(g, 4000, 150, 0.02, seed=9)
    m = pu.make_mapper(idx, expand_sa=True, paired=True)
    assert pu.compare_pairs(m, pu.Oracle(prefix), pu.interleave(r1, r2)) == 0
    assert m.work()["lf_steps"] == 0


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
