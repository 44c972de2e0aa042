"""Parity of the CUDA path with the oracle, through the C ABI, on a real B200 (-m gpu).

Bar (north-star): HLL registers and exact counts bit-identical; cardinalities within 1e-6
relative (we hold 1e-9); sizes the oracle finishes in seconds, plus size-independent properties
at the full config-2 genome size."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.util import adversarial_fasta, decode_packed, mutate, random_bases, to_fasta

pytestmark = pytest.mark.gpu
CARD_RTOL = 1e-9   # target is 1e-6 (BASELINE.json north_star); same f64 algorithm => far tighter


@pytest.fixture(scope="module")
def eng():
    from dandd_b200 import build
    build.build()
    from dandd_b200.engine import Engine
    return Engine(0)


def packed_to_numpy(seq):
    n = seq.nsym
    return decode_packed(seq.codes.cpu().numpy().view(np.uint32), seq.invalid.cpu().numpy().view(np.uint32), n)


# ---------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("variant", ["plain", "crlf", "no_final_newline", "one_long_line", "preamble", "big"])
@pytest.mark.parametrize("chunk", [None, 4096 + 16, 100000])
def test_pack_bit_exact(eng, variant, chunk):
    rng = np.random.default_rng(hash(variant) % 1000)
    txt = adversarial_fasta(rng, n=70000)
    if variant == "crlf":
        txt = txt.replace(b"\n", b"\r\n")
    elif variant == "no_final_newline":
        txt = txt.rstrip(b"\n")
    elif variant == "one_long_line":
        txt = to_fasta([(b"x", random_bases(rng, 300001))], width=10 ** 9)
    elif variant == "preamble":
        txt = b"; comment line\nACGT\n" + txt
    elif variant == "big":
        txt = to_fasta([(b"c%d" % i, mutate(rng, random_bases(rng, 400000))) for i in range(5)], width=80)
    want = orc.fasta_symbols(txt)
    seq = eng.pack(txt, chunk_bytes=chunk)
    assert seq.nsym == want.size
    assert np.array_equal(packed_to_numpy(seq), want)


def test_pack_empty_and_headers_only(eng):
    for txt in (b"", b"no records at all\n", b">only a header", b">h1\n>h2\n\n>h3\n"):
        want = orc.fasta_symbols(txt)
        seq = eng.pack(txt)
        assert seq.nsym == want.size
        assert np.array_equal(packed_to_numpy(seq), want)


# ---------------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("p", [8, 12, 20])
@pytest.mark.parametrize("canon", [True, False])
def test_sketch_all_k_bit_exact(eng, p, canon):
    rng = np.random.default_rng(p)
    txt = adversarial_fasta(rng, n=50000)
    sym = orc.fasta_symbols(txt)
    ks = list(range(1, 33))
    regs, cards = eng.sketch(eng.pack(txt), ks, p=p, canon=canon)
    regs, cards = regs.cpu().numpy(), cards.cpu().numpy()
    for i, k in enumerate(ks):
        want = orc.hll_sketch(sym, k, p, canon)
        assert np.array_equal(regs[i], want), (k, p, canon)
        assert cards[i] == pytest.approx(orc.card(want, p), rel=CARD_RTOL)


def test_sketch_k_subsets_and_slots(eng):
    rng = np.random.default_rng(3)
    txt = to_fasta([(b"g", random_bases(rng, 30000))])
    sym = orc.fasta_symbols(txt)
    seq = eng.pack(txt)
    for ks in ([14], [10, 32], [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31], list(range(10, 33))):
        regs, _ = eng.sketch(seq, ks, p=10)
        regs = regs.cpu().numpy()
        for i, k in enumerate(sorted(ks)):
            assert np.array_equal(regs[i], orc.hll_sketch(sym, k, 10)), (ks, k)


def test_sketch_chunked_ranges_and_floor_filter_exact(eng):
    """Streaming updates over unaligned ranges, with the min-register floor refreshed between
    chunks (n/m = 800, so the floor is well above zero), change nothing."""
    rng = np.random.default_rng(4)
    txt = to_fasta([(b"a", random_bases(rng, 120000)), (b"b", random_bases(rng, 85000))], width=70)
    sym = orc.fasta_symbols(txt)
    seq = eng.pack(txt)
    ks = [4, 9, 16, 17, 25, 32]
    n = seq.nsym
    cuts = [0, 7, 16, 1000, 33333, 33334, 150001, n]
    a, _ = eng.sketch(seq, ks, p=8, ranges=list(zip(cuts[:-1], cuts[1:])))
    b, cb = eng.sketch(seq, ks, p=8, floor_every=20000)
    for i, k in enumerate(ks):
        want = orc.hll_sketch(sym, k, 8)
        assert np.array_equal(a[i].cpu().numpy(), want), k
        assert np.array_equal(b[i].cpu().numpy(), want), k
        assert float(cb[i]) == pytest.approx(orc.card(want, 8), rel=CARD_RTOL)


def test_sketch_degenerate_inputs(eng):
    for txt in (b"", b">h\n", b">h\nACG\n", b">h\nNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN\n"):
        regs, cards = eng.sketch(eng.pack(txt), [4, 20], p=8)
        assert int(regs.sum()) == 0 and float(cards.abs().sum()) == 0.0
    # poly-T >= 32 is a valid k-mer run (SURVEY.md A.6 default) and canonicalises to poly-A
    txt = b">t\n" + b"T" * 40 + b"\n"
    regs, _ = eng.sketch(eng.pack(txt), [32], p=8)
    assert np.array_equal(regs[0].cpu().numpy(), orc.hll_sketch(orc.fasta_symbols(txt), 32, 8))
    assert int((regs[0] > 0).sum()) == 1


def test_host_buffer_path_matches_oracle(eng):
    rng = np.random.default_rng(5)
    txt = b"leading junk\n" + adversarial_fasta(rng, n=40000)
    sym = orc.fasta_symbols(txt)
    ks = [7, 16, 21, 32]
    regs, cards = eng.sketch_fasta_host(txt, ks, p=12)
    for i, k in enumerate(ks):
        want = orc.hll_sketch(sym, k, 12)
        assert np.array_equal(regs[i], want)
        assert cards[i] == pytest.approx(orc.card(want, 12), rel=CARD_RTOL)


def test_host_buffer_path_multi_chunk_with_floor(eng):
    """> 32 MiB of text: exercises the double-buffered copy stream, chunked packing through the
    carried state and the floor refresh of the host path."""
    rng = np.random.default_rng(6)
    txt = to_fasta([(b"chr%d" % i, random_bases(rng, 9_000_000)) for i in range(4)], width=80)
    assert len(txt) > (32 << 20)
    sym = orc.fasta_symbols(txt)
    ks = [12, 31]
    regs, cards = eng.sketch_fasta_host(txt, ks, p=12)
    for i, k in enumerate(ks):
        want = orc.hll_sketch(sym, k, 12)
        assert np.array_equal(regs[i], want)
        assert cards[i] == pytest.approx(orc.card(want, 12), rel=CARD_RTOL)


def test_full_size_genome_properties(eng):
    """Config-2 size (5 Mbp, p=20): bit-exact against the oracle for three k, plus properties
    that need no oracle: idempotence (sketching twice == once) and union == sketch of both."""
    rng = np.random.default_rng(2)
    anc = random_bases(rng, 5_000_000)
    ta = to_fasta([(b"anc", anc)], width=80)
    tb = to_fasta([(b"mut", mutate(rng, anc))], width=80)
    ks = list(range(10, 33))
    sa, sb = eng.pack(ta), eng.pack(tb)
    ra, ca = eng.sketch(sa, ks, p=20)
    rb, _ = eng.sketch(sb, ks, p=20)
    syma = orc.fasta_symbols(ta)
    for k in (10, 21, 32):
        want = orc.hll_sketch(syma, k, 20)
        assert np.array_equal(ra[k - 10].cpu().numpy(), want), k
        assert float(ca[k - 10]) == pytest.approx(orc.card(want, 20), rel=CARD_RTOL)
    ra2, _ = eng.sketch(sa, ks, p=20)
    assert bool((ra2 == ra).all())
    both = eng.pack(ta + tb)
    rab, _ = eng.sketch(both, ks, p=20)
    assert bool((rab == eng.union([ra, rb])).all())


# ------------------------------------------------------------------------------------------ K3/K4/K6
def make_sketches(eng, rng, n_genomes, ks, p, length=20000):
    anc = random_bases(rng, length)
    regs, syms = [], []
    for g in range(n_genomes):
        txt = to_fasta([(b"g%d" % g, mutate(rng, anc, sub=0.05))])
        syms.append(orc.fasta_symbols(txt))
        regs.append(eng.sketch(eng.pack(txt), ks, p=p)[0])
    import torch
    return torch.stack(regs).contiguous(), syms


def test_cards_and_union(eng):
    import torch
    rng = np.random.default_rng(7)
    ks, p = [8, 14, 27], 12
    regs, _ = make_sketches(eng, rng, 5, ks, p)
    cards = eng.cards(regs, p).cpu().numpy()
    h = regs.cpu().numpy()
    for g in range(5):
        for i in range(len(ks)):
            assert cards[g, i] == pytest.approx(orc.card(h[g, i], p), rel=CARD_RTOL)
    u = eng.union([regs[g] for g in range(5)]).cpu().numpy()
    assert np.array_equal(u, h.max(axis=0))
    # empty and saturated sketches
    z = torch.zeros((1, 1 << p), dtype=torch.uint8, device=regs.device)
    assert float(eng.cards(z, p)) == 0.0
    s = torch.full((1, 1 << p), 64 - p + 1, dtype=torch.uint8, device=regs.device)
    assert np.isinf(float(eng.cards(s, p)))


@pytest.mark.parametrize("p", [8, 14, 16])
def test_prefix_union_cards(eng, p):
    rng = np.random.default_rng(8 + p)
    ks = [10, 15, 20, 31]
    n = 6
    regs, _ = make_sketches(eng, rng, n, ks, p)
    orders = [list(range(n)), [5, 3, 1, 0, 2, 4], [2, 2, 2, 1, 1, 0], [4, -1, 0, -1, 3, 1]]
    cards, unions = eng.prefix_union_cards(regs, orders, p, materialize=True)
    cards, unions, h = cards.cpu().numpy(), unions.cpu().numpy(), regs.cpu().numpy()
    for o, order in enumerate(orders):
        run = np.zeros_like(h[0])
        for s, g in enumerate(order):
            if g >= 0:
                run = np.maximum(run, h[g])
            assert np.array_equal(unions[o, s], run)
            for i in range(len(ks)):
                assert cards[o, s, i] == pytest.approx(orc.card(run[i], p), rel=CARD_RTOL)
    fin = eng.prefix_union_cards(regs, orders, p, final_only=True).cpu().numpy()
    assert fin.shape == (len(orders), 1, len(ks))
    assert np.allclose(fin[:, 0], cards[:, -1], rtol=0, atol=0)


def test_pairwise_union_cards(eng):
    rng = np.random.default_rng(9)
    ks, p, n = [12, 21], 12, 7
    regs, _ = make_sketches(eng, rng, n, ks, p)
    pairs = [(a, b) for a in range(n) for b in range(a, n)]
    got = eng.pairwise_cards(regs, pairs, p).cpu().numpy()
    h = regs.cpu().numpy()
    for j, (a, b) in enumerate(pairs):
        for i in range(len(ks)):
            assert got[j, i] == pytest.approx(orc.card(np.maximum(h[a, i], h[b, i]), p), rel=CARD_RTOL)


# ---------------------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("k", [1, 4, 11, 16, 17, 24, 32])
@pytest.mark.parametrize("canon", [True, False])
def test_exact_counts_progressive(eng, k, canon):
    rng = np.random.default_rng(10 + k)
    anc = random_bases(rng, 30000)
    txts = [adversarial_fasta(rng, n=20000)] + [to_fasta([(b"m%d" % i, mutate(rng, anc, sub=0.02))]) for i in range(3)]
    txts.append(b">polyT\n" + b"T" * 100 + b"\n")
    seqs = [eng.pack(t) for t in txts]
    syms = [orc.fasta_symbols(t) for t in txts]
    got = eng.exact_counts(seqs, k, canon)
    want = [orc.exact_count(syms[:i + 1], k, canon) for i in range(len(syms))]
    assert got == want


def test_exact_table_overflow_is_loud(eng):
    from dandd_b200._lib import DandDError
    rng = np.random.default_rng(11)
    seq = eng.pack(to_fasta([(b"x", random_bases(rng, 50000))]))
    with pytest.raises(DandDError):
        eng.exact_counts([seq], 25, capacity=1024)


def test_small_k_persistent_bitmaps_bit_exact(eng):
    """k <= 9 goes through the persistent kernel with per-CTA presence bitmaps; 6.5 Mbp gives every
    CTA several tiles, low-complexity stretches give heavily repeated k-mers, and the chunked
    variant re-enters the kernel with fresh bitmaps mid-genome."""
    rng = np.random.default_rng(12)
    a = random_bases(rng, 6_500_000)
    a[1_000_000:1_200_000] = ord("A")                       # homopolymer
    a[2_000_000:2_300_000] = np.tile(np.frombuffer(b"ACGTTGCA", dtype=np.uint8), 37500)   # tandem repeat
    a[3_000_000:3_000_050] = ord("N")
    txt = to_fasta([(b"lowcomplex", a)], width=80)
    sym = orc.fasta_symbols(txt)
    seq = eng.pack(txt)
    ks = list(range(1, 14)) + [21, 32]
    regs, cards = eng.sketch(seq, ks, p=14)
    chunked, _ = eng.sketch(seq, ks, p=14, floor_every=1_000_003)
    for i, k in enumerate(ks):
        want = orc.hll_sketch(sym, k, 14)
        assert np.array_equal(regs[i].cpu().numpy(), want), k
        assert np.array_equal(chunked[i].cpu().numpy(), want), k
        assert float(cards[i]) == pytest.approx(orc.card(want, 14), rel=CARD_RTOL)
    nc, _ = eng.sketch(seq, [3, 7, 9], p=14, canon=False)
    for i, k in enumerate([3, 7, 9]):
        assert np.array_equal(nc[i].cpu().numpy(), orc.hll_sketch(sym, k, 14, canon=False)), k


@pytest.mark.parametrize("p", [12, 15, 20])
def test_bit_plane_prefix_kernel_equals_byte_kernel(eng, p):
    """dd_prefix_union_card has two kernels (byte histograms / bit-sliced planes); on sketches whose
    values span every bucket -- empty registers, the usual 1..15 bulk, sparse and dense 16..31,
    32..47 and the 64-p+1 cap -- both must give the same histograms, hence identical cards."""
    import torch
    from dandd_b200._lib import check
    rng = np.random.default_rng(20 + p)
    n, nk, m = 5, 3, 1 << p
    cap = 64 - p + 1
    regs = np.minimum(rng.geometric(0.5, size=(n, nk, m)) + rng.integers(0, 3, size=(n, nk, 1)), cap).astype(np.uint8)
    regs[0, 0, ::7] = 0
    regs[1, :, : m // 3] = np.minimum(rng.integers(14, 40, size=(nk, m // 3)), cap)      # dense high buckets
    regs[2, 1, 5] = cap
    regs[3, 2, :] = 0                                                                   # an empty sketch
    d = torch.from_numpy(regs).to(eng.device)
    orders = [[0, 1, 2, 3, 4], [3, 3, 2, 0, -1], [4, 1, -1, -1, 0]]
    check(eng.lib.dd_set_option(b"prefix_planes", 0))
    a = eng.prefix_union_cards(d, orders, p).cpu().numpy()
    af = eng.prefix_union_cards(d, orders, p, final_only=True).cpu().numpy()
    check(eng.lib.dd_set_option(b"prefix_planes", 1))
    b = eng.prefix_union_cards(d, orders, p).cpu().numpy()
    bf = eng.prefix_union_cards(d, orders, p, final_only=True).cpu().numpy()
    assert np.array_equal(a, b) and np.array_equal(af, bf)
    run = np.maximum(regs[0], regs[1])
    for i in range(nk):
        assert b[0, 1, i] == pytest.approx(orc.card(run[i], p), rel=CARD_RTOL)


def test_pack_unaligned_text_pointer(eng):
    """The bulk-copy (TMA) staging needs 16-byte aligned text; any other pointer takes the
    cooperative-copy fallback inside the same kernels and must give the same stream."""
    import torch
    from dandd_b200._lib import check
    rng = np.random.default_rng(31)
    txt = adversarial_fasta(rng, n=90000)
    want = orc.fasta_symbols(txt)
    n = len(txt)
    for shift in (1, 7, 13):
        base = torch.zeros(n + 64, dtype=torch.uint8, device=eng.device)
        base[shift:shift + n] = torch.from_numpy(np.frombuffer(txt, dtype=np.uint8).copy()).to(eng.device)
        assert (base.data_ptr() + shift) % 16 != 0
        cb, ib = eng.lib.dd_pack_codes_bytes(n), eng.lib.dd_pack_invalid_bytes(n)
        codes = torch.empty(cb, dtype=torch.uint8, device=eng.device)
        inval = torch.empty(ib, dtype=torch.uint8, device=eng.device)
        state = torch.empty(32, dtype=torch.uint8, device=eng.device)
        ws = torch.empty(eng.lib.dd_pack_workspace_bytes(n), dtype=torch.uint8, device=eng.device)
        check(eng.lib.dd_pack_reset(codes.data_ptr(), cb, inval.data_ptr(), ib, state.data_ptr(), eng.stream))
        check(eng.lib.dd_pack_fasta(base.data_ptr() + shift, n, codes.data_ptr(), inval.data_ptr(), n, state.data_ptr(),
                                    ws.data_ptr(), ws.numel(), eng.stream))
        nsym = int(state.cpu().numpy().view(np.uint64)[0])
        assert nsym == want.size
        got = decode_packed(codes.cpu().numpy().view(np.uint32), inval.cpu().numpy().view(np.uint32), nsym)
        assert np.array_equal(got, want), shift


def _simple_tile_cases():
    rng = np.random.default_rng(77)
    letters = np.frombuffer(b"ACGTacgtNnRYKMswbdhvUuXx", dtype=np.uint8)

    def noisy(n, frac=0.01):
        s = np.frombuffer(random_bases(rng, n), dtype=np.uint8).copy()
        idx = rng.choice(n, size=max(1, int(n * frac)), replace=False)
        s[idx] = letters[rng.integers(0, letters.size, idx.size)]
        return s.tobytes()

    def wrap(seq, width):
        return b"\n".join(seq[i:i + width] for i in range(0, len(seq), width)) + b"\n"

    cases = {}
    cases["w60"] = b">r1 plain record\n" + wrap(noisy(200_000), 60)
    cases["w7_multi_newline_per_chunk"] = b">r\n" + wrap(noisy(120_000, 0.05), 7)
    cases["w1"] = b">r\n" + wrap(noisy(40_000, 0.05), 1)
    cases["one_line"] = b">r\n" + noisy(150_000) + b"\n"
    cases["letters_only_header_spanning_tiles"] = b">" + noisy(45_000) + b"\n" + wrap(noisy(100_000), 80)
    cases["blank_line_runs"] = b">r\n" + wrap(noisy(50_000), 61).replace(b"\n", b"\n\n\n", 300) + wrap(noisy(50_000), 61)
    hi = np.frombuffer(noisy(100_000), dtype=np.uint8).copy()
    hi[rng.choice(hi.size, 500, replace=False)] = rng.integers(0xC0, 0x100, 500).astype(np.uint8)
    cases["high_bytes"] = b">r\n" + wrap(hi.tobytes(), 70)
    # general tiles (headers, CR, digits, blanks) interleaved with simple ones, at tile-sized distances
    mixed = b""
    for i in range(12):
        mixed += b">rec%d some text\r\n" % i + wrap(noisy(int(rng.integers(10_000, 50_000))), 60)
        mixed += b"ACGT 12 acgt\n" if i % 3 == 0 else b""
    cases["mixed"] = mixed
    cases["no_final_newline"] = (b">r\n" + wrap(noisy(70_000), 60)).rstrip(b"\n")
    cases["exact_tiles"] = (b">r\n" + wrap(noisy(70_000), 60))[:4 * 16384]
    return cases


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [None, 50_001, 16384, 65536 + 16])
def test_pack_simple_tiles(eng, chunk):
    """K1's fast path (tiles holding only letters and newlines) against the oracle, including its
    seams with general tiles, header state carried into a simple tile, and chunked/unaligned calls."""
    for name, txt in _simple_tile_cases().items():
        want = orc.fasta_symbols(txt)
        seq = eng.pack(txt, chunk_bytes=chunk)
        assert seq.nsym == want.size, name
        got = packed_to_numpy(seq)
        if not np.array_equal(got, want):
            bad = np.flatnonzero(got != want)
            raise AssertionError((name, chunk, bad[:10].tolist(), bad.size))


@pytest.mark.gpu
@pytest.mark.parametrize("n,p", [(5, 12), (70, 12)])
def test_prefix_union_identical_prefix_sets(eng, n, p):
    """Orderings that share a prefix SET share the histogram row (deduplicated on the device for
    n <= 64 genomes, counted independently above that); either way every (ordering, step) must carry
    the cardinality of its own running union."""
    rng = np.random.default_rng(100 + n)
    ks = [11, 21]
    regs, _ = make_sketches(eng, rng, n, ks, p, length=3000)
    base = [list(rng.permutation(n)) for _ in range(6)]
    orders = base + [list(reversed(base[0])), base[1][:2][::-1] + base[1][2:], [0] * n, base[2][:-1] + [-1]]
    orders = [[int(g) for g in o] for o in orders]
    cards = eng.prefix_union_cards(regs, orders, p).cpu().numpy()
    fin = eng.prefix_union_cards(regs, orders, p, final_only=True).cpu().numpy()
    h = regs.cpu().numpy()
    memo = {}
    for o, order in enumerate(orders):
        run = np.zeros_like(h[0])
        for s, g in enumerate(order):
            if g >= 0:
                run = np.maximum(run, h[g])
            for i in range(len(ks)):
                key = run[i].tobytes()
                if key not in memo:
                    memo[key] = orc.card(run[i], p)
                assert cards[o, s, i] == pytest.approx(memo[key], rel=CARD_RTOL), (o, s, i)
        assert np.array_equal(fin[o, 0], cards[o, -1])


@pytest.mark.gpu
@pytest.mark.parametrize("k", [33, 40, 63, 64])
@pytest.mark.parametrize("canon", [True, False])
def test_exact_counts_wide_k(eng, k, canon):
    """--exact above k = 32 (README.md:82: "to sweep higher ks you must use KMC via --exact"): 128-bit
    keys, `atom.cas.b128`.  The last text is the reverse complement of part of the first, so the
    canonical and plain unions differ; poly-T at k = 64 is the all-ones key (the empty marker)."""
    rng = np.random.default_rng(500 + k)
    anc = random_bases(rng, 40000)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    rc = bytes(comp[b] for b in anc[10000:25000].tolist()[::-1])
    txts = [to_fasta([(b"a", anc)]), adversarial_fasta(rng, n=20000),
            to_fasta([(b"m", mutate(rng, anc, sub=0.02))], width=61), b">polyT\n" + b"T" * 200 + b"\n",
            to_fasta([(b"rc", np.frombuffer(rc, dtype=np.uint8))])]
    seqs = [eng.pack(t) for t in txts]
    syms = [orc.fasta_symbols(t) for t in txts]
    got = eng.exact_counts(seqs, k, canon)
    want = [orc.exact_count(syms[:i + 1], k, canon) for i in range(len(syms))]
    assert got == want
    if canon:
        assert got[-1] == got[-2]           # the reverse complement adds nothing to a canonical set
    else:
        assert got[-1] > got[-2]


@pytest.mark.gpu
@pytest.mark.parametrize("k", [65, 99, 128, 129, 200, 256])
@pytest.mark.parametrize("canon", [True, False])
def test_exact_counts_long_k(eng, k, canon):
    """--exact for 64 < k <= 256 (KMC's own limit; helpers/allpairs.py:295 sweeps to 99): entries hold a
    fingerprint + a reference to one occurrence, matches are verified against the packed stream.  The
    oracle side is the string formulation (sort + unique of k-byte strings).  Progressive counts over
    several streams exercise references into earlier streams; the reverse-complement text exercises
    canonicalisation across word boundaries; a low-complexity text forces many verified duplicates."""
    rng = np.random.default_rng(900 + k)
    anc = random_bases(rng, 30000)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    rc = bytes(comp[b] for b in anc[5000:20000].tolist()[::-1])
    repeat = np.tile(random_bases(rng, 700), 12)
    txts = [to_fasta([(b"a", anc)]), adversarial_fasta(rng, n=15000),
            to_fasta([(b"m", mutate(rng, anc, sub=0.01))], width=61), b">polyT\n" + b"T" * 600 + b"\n",
            to_fasta([(b"rc", np.frombuffer(rc, dtype=np.uint8))]), to_fasta([(b"rep", repeat)], width=73)]
    seqs = [eng.pack(t) for t in txts]
    syms = [orc.fasta_symbols(t) for t in txts]
    got = eng.exact_counts(seqs, k, canon)
    want = [orc.exact_count(syms[:i + 1], k, canon) for i in range(len(syms))]
    assert got == want
    if canon:
        assert got[4] == got[3]             # the reverse complement adds nothing to a canonical set
    # key-range shards add up to the whole
    parts = [eng.exact_counts(seqs, k, canon, shard=(r, 3)) for r in range(3)]
    assert [sum(p[i] for p in parts) for i in range(len(seqs))] == want


@pytest.mark.gpu
def test_exact_k_above_256_is_refused(eng):
    from dandd_b200._lib import DandDError
    seq = eng.pack(b">x\n" + b"ACGT" * 100 + b"\n")
    with pytest.raises(DandDError):
        eng.exact_counts([seq], 257)


@pytest.mark.gpu
def test_pairwise_union_cards_large_batch(eng):
    """A batch of distinct pairs large enough that the quadratic identical-prefix search of the prefix
    kernel is skipped (rep == nullptr path): 60 sketches, all 1830 pairs incl. (a, a)."""
    rng = np.random.default_rng(19)
    ks, p, n = [13, 24], 12, 60
    regs, _ = make_sketches(eng, rng, n, ks, p, length=2500)
    pairs = [(a, b) for a in range(n) for b in range(a, n)]
    got = eng.pairwise_cards(regs, pairs, p).cpu().numpy()
    h = regs.cpu().numpy()
    memo = {}
    for j in rng.choice(len(pairs), 150, replace=False).tolist() + [0, len(pairs) - 1]:
        a, b = pairs[j]
        for i in range(len(ks)):
            u = np.maximum(h[a, i], h[b, i])
            key = u.tobytes()
            if key not in memo:
                memo[key] = orc.card(u, p)
            assert got[j, i] == pytest.approx(memo[key], rel=CARD_RTOL), (a, b, i)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [9, 16, 21, 32, 40, 64])
def test_exact_counts_key_range_shards_add_up(eng, k):
    """dd_exact_insert_shard: the distinct counts of the world's key ranges add up to the unsharded
    progressive counts (what sum_counts() all-reduces over ranks), for the bitmap, 64-bit and 128-bit sets."""
    rng = np.random.default_rng(700 + k)
    anc = random_bases(rng, 30000)
    txts = [to_fasta([(b"a", anc)]), to_fasta([(b"m", mutate(rng, anc, sub=0.03))], width=70),
            adversarial_fasta(rng, n=15000), b">polyT\n" + b"T" * 150 + b"\n"]
    seqs = [eng.pack(t) for t in txts]
    want = eng.exact_counts(seqs, k)
    assert want == [orc.exact_count([orc.fasta_symbols(t) for t in txts[:i + 1]], k) for i in range(len(txts))]
    for world in (2, 3, 8):
        parts = [eng.exact_counts(seqs, k, shard=(r, world)) for r in range(world)]
        assert [sum(col) for col in zip(*parts)] == want, world
        if k >= 16:
            assert all(p[-1] > 0 for p in parts)      # every shard holds a share of the keys
    from dandd_b200._lib import DandDError
    with pytest.raises(DandDError):
        eng.exact_counts(seqs, k, shard=(2, 2))


@pytest.mark.gpu
def test_prefix_union_many_orderings(eng):
    """More orderings than the identical-prefix search window (256): rows beyond it either find
    their class among the first 256 orderings or are counted themselves -- all must be right."""
    rng = np.random.default_rng(321)
    n, p, ks = 6, 12, [15]
    regs, _ = make_sketches(eng, rng, n, ks, p, length=2500)
    orders = [[int(g) for g in rng.permutation(n)] for _ in range(400)]
    orders[300] = [5, 5, 5, 4, 4, 4]           # a class that first appears beyond the window
    orders[350] = [5, 5, 5, 4, 4, 4]
    cards = eng.prefix_union_cards(regs, orders, p).cpu().numpy()
    h = regs.cpu().numpy()
    memo = {}
    for o, order in enumerate(orders):
        run = np.zeros_like(h[0, 0])
        for s, g in enumerate(order):
            run = np.maximum(run, h[g, 0])
            key = run.tobytes()
            if key not in memo:
                memo[key] = orc.card(run, p)
            assert cards[o, s, 0] == pytest.approx(memo[key], rel=CARD_RTOL), (o, s)
