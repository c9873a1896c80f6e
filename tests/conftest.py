import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Everything native is built in-tree once per session (idempotent; seconds when up to date)."""
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "kart_b200", "libkartb200.so"), os.path.join(ROOT, "oracle", "libkartoracle.so"),
            os.path.join(ROOT, "tests", "emul", "libkartb200_emul.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
    return True
