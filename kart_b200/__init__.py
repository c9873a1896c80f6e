"""kart_b200: B200-native (sm_100a) implementation of Kart's per-read hot path behind a C ABI.

Layout: csrc/ CUDA kernels + C ABI (libkartb200.so), host/ the C++ `kart`-compatible command line,
binding.py / index.py / synth.py the Python host-side mirror used by tests and bench.py."""
from .index import KartIndex            # noqa: F401
from .binding import Mapper, KartB200Error, load_library, DEFAULT_LIB   # noqa: F401
