"""The shipped host+device arithmetic (dandd_b200/csrc/common.cuh) executed on the CPU through
tests/host_emul.cpp and checked against the oracle: packer classification + transition-function
algebra, window extraction, canonical k-mers, Wang hash, register rule, Ertl MLE.  These are the
bit-deciding pieces of the CUDA kernels; the -m gpu tests check the kernels themselves."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.util import adversarial_fasta, decode_packed, kseq_fasta, random_bases, to_fasta

U8P, U32P = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)


def emul_pack(L, text: bytes, chunk=None):
    """Pack `text` (already starting at its first '>') in one or several chunks."""
    n = len(text)
    codes = np.zeros(n // 16 + 4, dtype=np.uint32)
    invalid = np.zeros(n // 32 + 4, dtype=np.uint32)
    last, hdr = C.c_uint32(ord("\n")), C.c_uint32(0)
    buf = np.frombuffer(text, dtype=np.uint8)
    nsym, off = 0, 0
    chunk = chunk or max(n, 1)
    while off < n:
        part = np.ascontiguousarray(buf[off:off + chunk])
        got = L.emul_pack(part.ctypes.data_as(U8P), part.size, codes.ctypes.data_as(U32P), invalid.ctypes.data_as(U32P),
                          nsym, C.byref(last), C.byref(hdr))
        assert got != 2 ** 64 - 1, "transition-function algebra disagrees with the sequential walk"
        nsym += got
        off += chunk
    return codes, invalid, nsym


def emul_sketch(L, codes, invalid, nsym, k, p, canon=True, ranges=None):
    regs = np.zeros(1 << p, dtype=np.uint8)
    for b, e in (ranges or [(0, nsym)]):
        L.emul_sketch(codes.ctypes.data_as(U32P), invalid.ctypes.data_as(U32P), b, e, k, p, int(canon),
                      regs.ctypes.data_as(U8P))
    return regs


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("chunk", [None, 16, 48, 1000, 4096])
def test_pack_matches_oracle_symbols(host_emul, seed, chunk):
    rng = np.random.default_rng(seed)
    txt = adversarial_fasta(rng, n=3000 + 517 * seed)
    if seed == 1:
        txt = txt.replace(b"\n", b"\r\n")         # CRLF
    if seed == 2:
        txt = txt.rstrip(b"\n")                   # no trailing newline
    if seed == 3:
        txt = to_fasta([(b"one-line", random_bases(rng, 5000))], width=100000)  # single long line
    want = orc.fasta_symbols(txt)
    codes, invalid, nsym = emul_pack(host_emul, txt, chunk)
    assert nsym == want.size
    assert np.array_equal(decode_packed(codes, invalid, nsym), want)


@pytest.mark.parametrize("chunk", [None, 16, 80, 4096])
def test_pack_kseq_markers(host_emul, chunk):
    """'@' at the start of a line opens a record like '>' does; '@', '+', '>' inside a line are
    plain (invalid) sequence characters."""
    rng = np.random.default_rng(7)
    txt = kseq_fasta(rng, n=4000)
    want = orc.fasta_symbols(txt)
    start = min(i for i in (txt.find(b">"), txt.find(b"@")) if i >= 0)
    codes, invalid, nsym = emul_pack(host_emul, txt[start:], chunk)
    assert nsym == want.size
    assert np.array_equal(decode_packed(codes, invalid, nsym), want)


def test_pack_header_spanning_chunks(host_emul):
    txt = b">" + b"x" * 100 + b"\nACGT\n>" + b"y" * 37 + b">>>\nGGGG\n>z"
    want = orc.fasta_symbols(txt)
    for chunk in (16, 32, 64, 7 * 16):
        codes, invalid, nsym = emul_pack(host_emul, txt, chunk)
        assert np.array_equal(decode_packed(codes, invalid, nsym), want), chunk


@pytest.mark.parametrize("k", [1, 2, 3, 8, 15, 16, 17, 23, 31, 32])
@pytest.mark.parametrize("canon", [True, False])
def test_sketch_arithmetic_matches_oracle(host_emul, k, canon):
    rng = np.random.default_rng(100 + k)
    txt = adversarial_fasta(rng, n=5000)
    sym = orc.fasta_symbols(txt)
    codes, invalid, nsym = emul_pack(host_emul, txt)
    for p in (8, 12):
        want = orc.hll_sketch(sym, k, p, canon)
        got = emul_sketch(host_emul, codes, invalid, nsym, k, p, canon)
        assert np.array_equal(got, want), (k, p, canon)
    # chunked updates over arbitrary (unaligned) symbol ranges are exact
    cuts = [0, 5, 16, 37, 1000, 1001, 4097, nsym]
    got = emul_sketch(host_emul, codes, invalid, nsym, k, 12, canon, ranges=list(zip(cuts[:-1], cuts[1:])))
    assert np.array_equal(got, orc.hll_sketch(sym, k, 12, canon))


def test_hash_rank_index_formulations(host_emul):
    rng = np.random.default_rng(5)
    xs = rng.integers(0, 1 << 63, 2000, dtype=np.uint64).tolist() + [0, 1, (1 << 64) - 1, 1 << 43, (1 << 44) - 1, 1 << 44]
    for x in xs:
        assert host_emul.emul_wang(x) == orc.wang(x)
        for p in (4, 10, 20, 26):
            q = 64 - p
            rest = x & ((1 << q) - 1)
            rank = q - rest.bit_length() + 1
            assert host_emul.emul_rank(x, p) == rank
            assert host_emul.emul_rank_split(x, p) == rank
            assert host_emul.emul_index(x, p) == x >> q
            assert host_emul.emul_index_split(x, p) == x >> q


def test_device_mle_matches_oracle(host_emul):
    rng = np.random.default_rng(9)
    for p, n in [(8, 10), (8, 3000), (12, 100), (12, 50000), (16, 10 ** 6), (20, 5 * 10 ** 6)]:
        h = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
        q = 64 - p
        idx = (h >> np.uint64(q)).astype(np.int64)
        rest = (h & np.uint64((1 << q) - 1))
        lz = q - np.floor(np.log2(np.maximum(rest.astype(np.float64), 1.0))).astype(np.int64)  # approximate, fine here
        regs = np.zeros(1 << p, dtype=np.uint8)
        np.maximum.at(regs, idx, np.clip(lz, 1, q + 1).astype(np.uint8))
        c = orc.hist(regs, p)
        hist64 = np.ascontiguousarray(c[:64])
        got = host_emul.emul_mle(hist64.ctypes.data_as(U32P), p)
        assert got == pytest.approx(orc.ertl_mle(c, p), rel=1e-12)


@pytest.mark.parametrize("k", [33, 34, 40, 48, 49, 63, 64])
@pytest.mark.parametrize("canon", [True, False])
def test_wide_kmer_values_match_big_integers(host_emul, k, canon):
    """k = 33..64 (exact mode): the 128-bit window extraction / reverse complement / canonical
    choice of exact_insert_wide_kernel against Python integers on the oracle's symbol stream."""
    import ctypes as C
    rng = np.random.default_rng(300 + k)
    txt = adversarial_fasta(rng, n=3000)
    sym = orc.fasta_symbols(txt)
    codes, invalid, nsym = emul_pack(host_emul, txt)
    out = np.zeros(2 * max(1, nsym), dtype=np.uint64)
    fn = host_emul.emul_kmers_wide
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    cnt = fn(codes.ctypes.data, invalid.ctypes.data, 0, nsym, k, int(canon), out.ctypes.data)
    got = [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(cnt)]
    want = []
    run, fwd, rc = 0, 0, 0
    mask = (1 << (2 * k)) - 1
    for c in sym.tolist():
        if c > 3:
            run, fwd, rc = 0, 0, 0
            continue
        fwd = ((fwd << 2) | c) & mask
        rc = (rc >> 2) | ((3 - c) << (2 * (k - 1)))
        run += 1
        if run >= k:
            want.append(min(fwd, rc) if canon else fwd)
    assert got == want
    assert len(set(got)) == orc.exact_count([sym], k, canon)
    # arbitrary ranges partition the k-mers by end position
    cuts = [0, 7, 16, 100, 1001, nsym]
    tot = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        tot += fn(codes.ctypes.data, invalid.ctypes.data, a, b, k, int(canon), out.ctypes.data)
    assert tot == cnt


def test_pack_fuzz_against_oracle(host_emul):
    """Property test (hypothesis): any byte string over a FASTA-ish alphabet, cut into chunks of any
    16-byte multiple, packs to exactly the oracle's symbol stream -- the transition-function algebra
    (dd::chunk_xfer / xfer_compose) and the per-chunk emission (dd::chunk_symbols) shipped in the kernels."""
    from hypothesis import given, settings, strategies as st
    alphabet = b">ACGTNacgtn\n\n\r x;-"

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.integers(0, len(alphabet) - 1), min_size=0, max_size=400), st.integers(1, 12),
           st.booleans())
    def run(idx, chunk16, lead_header):
        txt = bytes(alphabet[i] for i in idx)
        if lead_header:
            txt = b">h\n" + txt
        want = orc.fasta_symbols(txt)
        # the kernels are handed the text from its first '>' on, wherever that is (Engine.skip_preamble;
        # kseq skips to the first record marker and treats it as opening a header line)
        start = txt.find(b">")
        if start < 0:
            assert want.size == 0
            return
        codes, invalid, nsym = emul_pack(host_emul, txt[start:], 16 * chunk16)
        assert nsym == want.size
        assert np.array_equal(decode_packed(codes, invalid, nsym), want)

    run()


def test_sketch_fuzz_against_oracle(host_emul):
    """Property test: random short FASTA texts (breaks, lower case, odd line widths), random k, p,
    strand mode and update ranges -- the window extraction / canonical choice / hash / rank of the
    kernels (host emulation) against the oracle's rolling formulation."""
    from hypothesis import given, settings, strategies as st
    alphabet = b"ACGTACGTACGTacgtN\n"

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.integers(0, len(alphabet) - 1), min_size=1, max_size=300), st.integers(1, 32), st.integers(4, 12),
           st.booleans(), st.lists(st.integers(0, 300), min_size=0, max_size=4))
    def run(idx, k, p, canon, cuts):
        txt = b">r\n" + bytes(alphabet[i] for i in idx) + b"\n"
        sym = orc.fasta_symbols(txt)
        codes, invalid, nsym = emul_pack(host_emul, txt)
        assert nsym == sym.size
        want = orc.hll_sketch(sym, k, p, canon)
        assert np.array_equal(emul_sketch(host_emul, codes, invalid, nsym, k, p, canon), want)
        pts = sorted({0, nsym} | {min(c, nsym) for c in cuts})
        ranges = list(zip(pts[:-1], pts[1:]))
        assert np.array_equal(emul_sketch(host_emul, codes, invalid, nsym, k, p, canon, ranges=ranges), want)

    run()


@pytest.mark.parametrize("k", [65, 96, 97, 128, 129, 200, 255, 256])
@pytest.mark.parametrize("canon", [True, False])
def test_long_kmers_match_string_formulation(host_emul, k, canon):
    """kmer_long_at (common.cuh; exact mode for 64 < k <= 256): the multi-word value of every valid window
    equals the base-4 number of the canonical symbol STRING, and valid_run_upto agrees with a walk."""
    rng = np.random.default_rng(k)
    txt = to_fasta([(b"a", random_bases(rng, 1500)), (b"b", random_bases(rng, 700))], width=61)
    txt = txt.replace(b"\n", b"N\n", 3)          # a few breaks
    sym = orc.fasta_symbols(txt)
    codes, invalid, nsym = emul_pack(host_emul, txt)
    out = (C.c_uint64 * 8)()
    run = 0
    checked = 0
    for s in range(nsym):
        run = 0 if sym[s] > 3 else run + 1
        got_run = host_emul.emul_valid_run(invalid.ctypes.data_as(U32P), s, k)
        assert min(run, k) == min(got_run, k), (s, run, got_run)
        if run < k or (s % 7 and s % 64 > 3):          # every window near word boundaries, a sample elsewhere
            continue
        w = host_emul.emul_kmer_long(codes.ctypes.data_as(U32P), s, k, int(canon), out)
        assert w == (2 * k + 63) // 64
        value = sum(int(out[t]) << (64 * t) for t in range(w))
        fwd = sym[s - k + 1:s + 1].tolist()
        want_syms = fwd
        if canon:
            rc = [3 - c for c in reversed(fwd)]
            want_syms = min(fwd, rc)
        want = 0
        for c in want_syms:
            want = (want << 2) | int(c)
        assert value == want, (s, k)
        checked += 1
    assert checked > 100
