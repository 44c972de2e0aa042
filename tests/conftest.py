import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def host_emul():
    """tests/host_emul.cpp compiled with g++: the shipped host+device arithmetic run on the CPU."""
    import ctypes as C
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhost_emul.so")
    srcs = [os.path.join(here, "host_emul.cpp"), os.path.join(ROOT, "dandd_b200", "csrc", "common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", srcs[0], "-o", so])
    L = C.CDLL(so)
    u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    L.emul_pack.restype = C.c_size_t
    L.emul_pack.argtypes = [u8p, C.c_size_t, u32p, u32p, C.c_size_t, u32p, u32p]
    L.emul_sketch.restype = None
    L.emul_sketch.argtypes = [u32p, u32p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, u8p]
    for name in ("emul_rank_split", "emul_index_split", "emul_rank", "emul_index"):
        getattr(L, name).restype = C.c_uint32
        getattr(L, name).argtypes = [C.c_uint64, C.c_int]
    L.emul_wang.restype = C.c_uint64
    L.emul_wang.argtypes = [C.c_uint64]
    L.emul_kmer_long.restype = C.c_int
    L.emul_kmer_long.argtypes = [u32p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.emul_valid_run.restype = C.c_int
    L.emul_valid_run.argtypes = [u32p, C.c_uint64, C.c_int]
    L.emul_mle.restype = C.c_double
    L.emul_mle.argtypes = [u32p, C.c_int]
    return L
