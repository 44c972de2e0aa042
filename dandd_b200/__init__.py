"""dandd_b200 -- B200-native (sm_100a) implementation of DandD's sketch-and-count hot path.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/dandd_b200.h)
  _lib.py          ctypes binding of that ABI (fails loudly when the library or a GPU is missing)
  engine.py        torch-owned buffers + the batched operations the host layer calls
  lib/             drop-in mirror of the reference's Python modules (sketch_classes, huffman_dandd,
                   species_specifics, dandd_cmd, `dandd` launcher) with the subprocess calls to
                   dashing / kmc / GNU parallel replaced by Engine calls

There is no CPU fallback anywhere in this package: the oracle under oracle/ is test
infrastructure and is never imported from here.
"""
import os
import sys

__version__ = "0.1.0"

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


def enable_compat():
    """Make `import sketch_classes`, `huffman_dandd`, `species_specifics`, `dandd_cmd` resolve to
    this package's GPU-backed mirrors -- the reference's modules are top-level (its lib/ directory
    is the script directory), and its pickles name classes by those module names
    (SURVEY.md Appendix D), so the mirrors must be importable under the same names."""
    if LIB_DIR not in sys.path:
        sys.path.insert(0, LIB_DIR)
    return LIB_DIR
