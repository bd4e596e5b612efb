"""dextractor_b200 -- B200-native DEXTRACTOR compression hot path.

The product is libdexb200.so (hand-written sm_100a CUDA kernels behind the C ABI declared in
include/dexb200.h) plus the C command-line tools in tools/.  This package is only the thin Python
view of that ABI used by tests/ and bench.py; it contains no codec logic and no CPU fallback:
without the built library or without a CUDA device every call fails loudly.
"""
from .lib import (ARROW, FASTA, LIB_PATH, Carry, Coding, Context, DexError, Stats,  # noqa: F401
                  build_library, load_library)
