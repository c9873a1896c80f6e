"""CPU: the product library loads, exports every symbol include/kart_b200.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import parity_util as pu
from kart_b200 import binding


def test_exports_match_header(built):
    hdr = open(os.path.join(pu.ROOT, "include", "kart_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(kb_[a-z_0-9]+)\s*\(", hdr)))
    lib = C.CDLL(binding.DEFAULT_LIB)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(binding.EXPORTS) == declared


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        return
    lib = binding.load_library()
    h = C.c_void_p()
    assert lib.kb_init(0, C.byref(h)) == -1          # KB_ENODEV
    assert b"no CUDA device" in lib.kb_strerror(-1)


def test_product_does_not_link_oracle_or_emulation(built):
    import subprocess
    out = subprocess.run(["ldd", binding.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emul" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", binding.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "kor_" not in syms and "kb_emul" not in syms
