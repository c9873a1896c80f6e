"""CPU: the stage-level parity checks of tests/stage_util.py over the host-emulated kernels (tests/emul, test infrastructure only).
The same checks run over the real kernels in tests/test_gpu_stage.py."""
import stage_util as su


def test_golden_seeds_and_candidates(built):
    su.check_golden_seeds_and_candidates(emul=True)


def test_golden_pair_stage(built):
    assert su.check_golden_pair_stage(emul=True) == 120


def test_golden_nw_vectors(built, tmp_path, monkeypatch):
    su.check_nw_vectors(True, str(tmp_path), monkeypatch)


def test_fragment_pairs_vs_oracle(built):
    assert su.check_fragment_pairs(emul=True) > 600
