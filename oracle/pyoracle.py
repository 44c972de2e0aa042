"""ctypes binding of oracle/dandd_oracle.c (TEST INFRASTRUCTURE ONLY; parity unpinned -- see the
C header for what that means and why).  Every function here cites the reference call site whose
external program it stands in for."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
SHIM = os.path.join(_HERE, "_build", "orc_shim")


def build(force=False):
    """Compile liboracle.so and the CLI shims with gcc (idempotent)."""
    src = os.path.join(_HERE, "dandd_oracle.c")
    shim_src = os.path.join(_HERE, "shims", "orc_shim.c")
    stale = (not os.path.exists(_SO) or not os.path.exists(SHIM)
             or os.path.getmtime(_SO) < os.path.getmtime(src)
             or os.path.getmtime(SHIM) < max(os.path.getmtime(src), os.path.getmtime(shim_src)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.orc_fasta_symbols.restype = C.c_size_t
        L.orc_fasta_symbols.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_polyt_sentinel.restype = None
        L.orc_polyt_sentinel.argtypes = [u8p, C.c_size_t]
        L.orc_wang.restype = C.c_uint64
        L.orc_wang.argtypes = [C.c_uint64]
        L.orc_revcomp.restype = C.c_uint64
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_hll_sketch.restype = None
        L.orc_hll_sketch.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_hist.restype = None
        L.orc_hist.argtypes = [u8p, C.c_int, u32p]
        L.orc_ertl_mle.restype = C.c_double
        L.orc_ertl_mle.argtypes = [u32p, C.c_int]
        L.orc_card.restype = C.c_double
        L.orc_card.argtypes = [u8p, C.c_int]
        L.orc_exact_count.restype = C.c_uint64
        L.orc_exact_count.argtypes = [C.POINTER(u8p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int]
        L.orc_kmers.restype = C.c_size_t
        L.orc_kmers.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, u64p, C.c_size_t]
        L.orc_sketch_fasta.restype = C.c_double
        L.orc_sketch_fasta.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, u8p]
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def fasta_symbols(text: bytes) -> np.ndarray:
    """FASTA bytes -> symbol stream (0..3 bases, 4 = break).  Stands in for kseq inside
    `dashing sketch` (lib/sketch_classes.py:358-365) and `kmc -fm` (:444-447)."""
    buf, bp = _u8(np.frombuffer(text, dtype=np.uint8))
    out = np.empty(max(1, buf.size), dtype=np.uint8)
    n = lib().orc_fasta_symbols(bp, buf.size, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out[:n].copy()


def polyt_sentinel(sym: np.ndarray) -> np.ndarray:
    """SURVEY.md A.6 switched ON: a copy of the symbol stream with every 32nd T of a T run broken."""
    out, op = _u8(np.array(sym, dtype=np.uint8, copy=True))
    lib().orc_polyt_sentinel(op, out.size)
    return out


def wang(x: int) -> int:
    return int(lib().orc_wang(C.c_uint64(x & 0xFFFFFFFFFFFFFFFF)))


def revcomp(v: int, k: int) -> int:
    return int(lib().orc_revcomp(C.c_uint64(v), k))


def hll_sketch(sym: np.ndarray, k: int, p: int = 20, canon: bool = True, regs=None) -> np.ndarray:
    """`dashing sketch -k K -S p` registers of one symbol stream (accumulates into regs)."""
    sym, sp = _u8(sym)
    if regs is None:
        regs = np.zeros(1 << p, dtype=np.uint8)
    lib().orc_hll_sketch(sp, sym.size, k, p, int(bool(canon)), regs.ctypes.data_as(C.POINTER(C.c_uint8)))
    return regs


def union_max(sketches) -> np.ndarray:
    """`dashing union` (lib/sketch_classes.py:368-373): register-wise max."""
    return np.maximum.reduce([np.asarray(s, dtype=np.uint8) for s in sketches])


def hist(regs: np.ndarray, p: int) -> np.ndarray:
    regs, rp = _u8(regs)
    c = np.zeros(66, dtype=np.uint32)
    lib().orc_hist(rp, p, c.ctypes.data_as(C.POINTER(C.c_uint32)))
    return c


def ertl_mle(counts: np.ndarray, p: int) -> float:
    c = np.zeros(66, dtype=np.uint32)
    c[:len(counts)] = counts
    return float(lib().orc_ertl_mle(c.ctypes.data_as(C.POINTER(C.c_uint32)), p))


def card(regs: np.ndarray, p: int) -> float:
    """`dashing card --presketched` (lib/sketch_classes.py:306-321)."""
    regs, rp = _u8(regs)
    assert regs.size == 1 << p
    return float(lib().orc_card(rp, p))


def exact_count(syms, k: int, canon: bool = True) -> int:
    """`kmc` + `kmc_tools info` / `complex` union (lib/sketch_classes.py:389-399,434-465):
    number of distinct (canonical) k-mers in the union of the given symbol streams."""
    arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in syms]
    u8p = C.POINTER(C.c_uint8)
    ptrs = (u8p * len(arrs))(*[a.ctypes.data_as(u8p) for a in arrs])
    lens = (C.c_size_t * len(arrs))(*[a.size for a in arrs])
    return int(lib().orc_exact_count(ptrs, lens, len(arrs), k, int(bool(canon))))


def kmers(sym: np.ndarray, k: int, canon: bool = True) -> np.ndarray:
    sym, sp = _u8(sym)
    n = lib().orc_kmers(sp, sym.size, k, int(bool(canon)), None, 0)
    out = np.empty(max(1, n), dtype=np.uint64)
    lib().orc_kmers(sp, sym.size, k, int(bool(canon)), out.ctypes.data_as(C.POINTER(C.c_uint64)), n)
    return out[:n]


def sketch_fasta(text: bytes, k: int, p: int = 20, canon: bool = True):
    """One whole `dashing sketch` + `dashing card` job: parse, sketch one k, estimate.  This is the
    unit of work the reference fans out with GNU parallel (lib/huffman_dandd.py:217); ctypes drops
    the GIL, so a thread pool over (file, k) reproduces that topology for the CPU baseline."""
    buf, bp = _u8(np.frombuffer(text, dtype=np.uint8))
    regs = np.empty(1 << p, dtype=np.uint8)
    c = lib().orc_sketch_fasta(bp, buf.size, k, p, int(bool(canon)), regs.ctypes.data_as(C.POINTER(C.c_uint8)))
    return regs, float(c)


def install_shims(bindir: str) -> str:
    """Create dashing / kmc / kmc_tools / parallel in `bindir` (symlinks to the oracle-backed
    shims) so the unmodified reference Python can run without the real binaries."""
    build()
    os.makedirs(bindir, exist_ok=True)
    for name in ("dashing", "kmc", "kmc_tools"):
        dst = os.path.join(bindir, name)
        if os.path.lexists(dst):
            os.remove(dst)
        os.symlink(SHIM, dst)
    dst = os.path.join(bindir, "parallel")
    if os.path.lexists(dst):
        os.remove(dst)
    os.symlink(os.path.join(_HERE, "shims", "parallel"), dst)
    return bindir
