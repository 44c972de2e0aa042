"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly the symbols
include/dandd_b200.h declares, refuses to run without a device (no CPU fallback), and the
product package never reaches into oracle/.  No compute call is made here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dandd_b200 import build, _lib
    build.build()
    return _lib.load()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dandd_b200.h")).read()
    return sorted(set(re.findall(r"DD_API\s+[\w\s\*]+?\b(dd_\w+)\s*\(", txt)))


def test_header_and_binding_agree(lib):
    from dandd_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 26
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "dandd_b200", "libdandd_b200.so")], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(header_symbols()) <= exported
    assert all(s.startswith("dd_") for s in exported), exported  # nothing else leaks out


def test_integration_doc_maps_every_entry():
    """INTEGRATION.md names, for every exported entry, the reference interface it stands in for."""
    import re
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    named = set(re.findall(r"`(dd_[a-z0-9_]+)`", doc))
    missing = sorted(set(header_symbols()) - named)
    assert not missing, missing


def test_abi_version_and_size_queries(lib):
    assert lib.dd_abi_version() == 2
    assert lib.dd_pack_codes_bytes(1 << 20) >= (1 << 20) // 4
    assert lib.dd_pack_invalid_bytes(1 << 20) >= (1 << 20) // 8
    assert lib.dd_sketch_workspace_bytes(23, 20) >= 23 * 2 * (1 << 20)
    assert lib.dd_exact_workspace_bytes(12, 0) >= (4 ** 12) // 8
    assert lib.dd_exact_workspace_bytes(20, 1 << 20) >= 8 << 20


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    assert lib.dd_init(0) < 0
    assert b"no CPU fallback" in lib.dd_last_error()
    from dandd_b200.engine import Engine
    from dandd_b200._lib import DandDError
    with pytest.raises(DandDError):
        Engine(0)


def test_argument_errors_are_reported(lib):
    assert lib.dd_sketch_begin(None, 0, 3, 20, None) == -1
    assert b"dd_sketch_begin" in lib.dd_last_error()
    assert lib.dd_exact_begin(None, 0, 40, 0, None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dandd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "dandd":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
                assert "liboracle" not in src and "pyoracle" not in src, os.path.join(dirpath, f)


def test_cli_startup_trim_is_safe():
    """The launchers drop torch's queued Triton-operator registration before CUDA initialises; the
    helper must never raise and must leave torch importable/usable whatever torch's internals are."""
    import torch
    from dandd_b200._startup import trim_torch_cuda_init
    before = len(getattr(torch.cuda, "_queued_calls", []))
    dropped = trim_torch_cuda_init()
    after = len(getattr(torch.cuda, "_queued_calls", []))
    assert isinstance(dropped, bool) and after <= before
    assert trim_torch_cuda_init() is False or after == len(torch.cuda._queued_calls)   # idempotent
    assert torch.zeros(2).sum().item() == 0.0


def test_skip_preamble_host_variants():
    """Engine.skip_preamble on host inputs (bytes: memchr; numpy / memoryview: blockwise scan)."""
    import numpy as np
    from dandd_b200.engine import Engine
    f = Engine.skip_preamble
    assert f(b">abc") == 0 and f(b"xx\n>abc") == 3 and f(b"none") == 4 and f(b"") == 0 and f(bytearray(b"a>")) == 1
    a = np.frombuffer(b"xx\n>abc", dtype=np.uint8)
    assert f(a) == 3 and f(memoryview(b"xx\n>abc")) == 3 and f(np.zeros(5, dtype=np.uint8)) == 5
    big = np.zeros((1 << 24) + 100, dtype=np.uint8)
    big[(1 << 24) + 7] = 62
    assert f(big) == (1 << 24) + 7


# ---- host-side text helpers of the library (no GPU work, so they run everywhere) ------------------
def _normalise(lib, text: bytes) -> bytes:
    import ctypes as C
    src = (C.c_uint8 * max(1, len(text))).from_buffer_copy(text or b"\0")
    dst = (C.c_uint8 * (len(text) + 16))()
    n = lib.dd_fastq_to_fasta_host(C.addressof(src), len(text), C.addressof(dst))
    assert n <= len(text) + 16
    return bytes(dst[:n])


def test_first_record_marker(lib):
    import ctypes as C
    for text, want in [(b">a\nAC", 0), (b"xx\n@r\nAC", 3), (b"junk>a", 4), (b"a@b>c", 1), (b"none", 4), (b"", 0)]:
        buf = (C.c_uint8 * max(1, len(text))).from_buffer_copy(text or b"\0")
        assert lib.dd_fasta_first_record_host(C.addressof(buf), len(text)) == want, text


def test_fastq_normaliser_matches_the_oracle_walk(lib):
    """dd_fastq_to_fasta_host is the product's restatement of kseq_read(); the oracle is another.
    Feeding the normalised text back through the oracle must give the symbols of the raw text, and
    the normalised text must be plain FASTA (no line begins with '+' or '@')."""
    import numpy as np
    from oracle import pyoracle as orc
    from tests.test_oracle import KSEQ_CASES
    texts = [t for t, _ in KSEQ_CASES]
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACGTacgtN>@+\n\n\n\r I", dtype=np.uint8)
    texts += [alpha[rng.integers(0, alpha.size, int(rng.integers(0, 80)))].tobytes() for _ in range(4000)]
    for text in texts:
        fasta = _normalise(lib, text)
        assert orc.fasta_symbols(fasta).tolist() == orc.fasta_symbols(text).tolist(), (text, fasta)
        assert all(not ln.startswith((b"+", b"@")) for ln in fasta.split(b"\n")), (text, fasta)
