"""Round-2 GPU parity tests (-m gpu): kseq record rules on the device path, the FASTQ detour, the
poly-T switch, the bit-plane entry points with caller-owned scratch, and parity at
the sizes BASELINE.json's configs name (config 2: every k at 5 Mbp / p=20; config 4: exact counts
at 100 Mbp; config 5: every pair x every k of 40 genomes at p=18)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.util import adversarial_fasta, decode_packed, kseq_fasta, mutate, random_bases, to_fasta

pytestmark = pytest.mark.gpu
CARD_RTOL = 1e-9
THREADS = max(1, min(32, (os.cpu_count() or 2) - 1))


@pytest.fixture(scope="module")
def eng():
    from dandd_b200 import build
    build.build()
    from dandd_b200.engine import Engine
    return Engine(0)


def packed_to_numpy(seq):
    n = seq.nsym
    return decode_packed(seq.codes.cpu().numpy().view(np.uint32), seq.invalid.cpu().numpy().view(np.uint32), n)


# ------------------------------------------------------------------------------------- kseq rules
@pytest.mark.parametrize("n", [4000, 300000])
@pytest.mark.parametrize("chunk", [None, 16384 + 16])
def test_pack_at_markers(eng, n, chunk):
    """'@' lines are headers, '@' '+' '>' inside lines are invalid bases; at n=300000 most tiles take
    the fast path and the ones holding an '@' must not."""
    rng = np.random.default_rng(n)
    txt = kseq_fasta(rng, n=n)
    want = orc.fasta_symbols(txt)
    seq = eng.pack(txt, chunk_bytes=chunk)
    assert seq.nsym == want.size
    assert np.array_equal(packed_to_numpy(seq), want)


def test_at_header_inside_fast_tiles(eng):
    """A FASTA whose later records use '@' markers, large enough for whole tiles of pure sequence
    around them; registers must match the oracle's."""
    rng = np.random.default_rng(3)
    recs = [(b"c%d" % i, random_bases(rng, 120000)) for i in range(4)]
    txt = to_fasta(recs, width=80).replace(b">c1", b"@c1").replace(b">c3", b"@c3 comment")
    sym = orc.fasta_symbols(txt)
    seq = eng.pack(txt)
    assert np.array_equal(packed_to_numpy(seq), sym)
    regs, cards = eng.sketch(seq, [15, 31], p=14)
    for i, k in enumerate((15, 31)):
        want = orc.hll_sketch(sym, k, 14)
        assert np.array_equal(regs[i].cpu().numpy(), want)
        assert float(cards[i]) == pytest.approx(orc.card(want, 14), rel=CARD_RTOL)


def test_fastq_is_flagged_and_detoured(eng):
    from dandd_b200.engine import FastqInput
    from dandd_b200.store import GpuSketchStore
    rng = np.random.default_rng(4)
    txt = kseq_fasta(rng, n=50000, fastq=True)
    sym = orc.fasta_symbols(txt)
    with pytest.raises(FastqInput):
        eng.pack(txt).check()
    fasta = eng.fastq_to_fasta(txt)
    assert np.array_equal(packed_to_numpy(eng.pack(fasta)), sym)
    store = GpuSketchStore(engine=eng)
    assert np.array_equal(packed_to_numpy(store._pack_text(txt)), sym)
    # host-buffer C ABI: the synchronous entry normalises and retries, the asynchronous one poisons
    ks = [12, 20, 32]
    regs, cards = eng.sketch_fasta_host(txt, ks, p=12)
    for i, k in enumerate(ks):
        want = orc.hll_sketch(sym, k, 12)
        assert np.array_equal(regs[i], want)
        assert cards[i] == pytest.approx(orc.card(want, 12), rel=CARD_RTOL)
    import torch
    pinned = torch.empty(len(ks), dtype=torch.float64).pin_memory()
    eng.sketch_fasta_host(txt, ks, p=12, want_regs=False, cards_out=pinned, sync=False)
    torch.cuda.synchronize()
    assert bool(torch.isnan(pinned).all())
    # a plain FASTA is not poisoned
    eng.sketch_fasta_host(fasta, ks, p=12, want_regs=False, cards_out=pinned, sync=False)
    torch.cuda.synchronize()
    assert not bool(torch.isnan(pinned).any())


def test_kseq_fuzz_device_vs_oracle(eng):
    """Random soups of the bytes that matter to the record rules: packer (general path) vs oracle;
    texts with a line-initial '+' must be flagged, and their normalised form must match."""
    from dandd_b200.engine import FastqInput
    rng = np.random.default_rng(12)
    alpha = np.frombuffer(b"ACGTacgtN>@+\n\n\n\r I", dtype=np.uint8)
    for _ in range(150):
        txt = alpha[rng.integers(0, alpha.size, int(rng.integers(1, 400)))].tobytes()
        want = orc.fasta_symbols(txt)
        try:
            got = packed_to_numpy(eng.pack(txt))
        except FastqInput:
            got = packed_to_numpy(eng.pack(eng.fastq_to_fasta(txt)))
        assert np.array_equal(got, want), txt


# ------------------------------------------------------------------------------------- A.6 switch
def test_polyt_sentinel_switch(eng):
    from dandd_b200._lib import check
    rng = np.random.default_rng(6)
    body = random_bases(rng, 40000)
    body[100:170] = ord("T")            # 70 T: breaks at the 32nd and 64th
    body[5000:5031] = ord("t")          # 31: untouched
    body[9000:9032] = ord("T")          # exactly 32
    body[20000:20500] = ord("T")        # long run, crosses chunk boundaries below
    txt = to_fasta([(b"a", body), (b"b", np.full(100, ord("T"), dtype=np.uint8))], width=70)
    base = orc.fasta_symbols(txt)
    want = orc.polyt_sentinel(base)
    assert (want != base).sum() >= 2 + 1 + 15 + 3
    assert not eng.polyt_sentinel
    assert np.array_equal(packed_to_numpy(eng.pack(txt)), base)          # default: poly-T is valid sequence
    eng.polyt_sentinel = True
    try:
        for chunk in (None, 4096 + 16, 20480):
            assert np.array_equal(packed_to_numpy(eng.pack(txt, chunk_bytes=chunk)), want), chunk
        regs, _ = eng.sketch(eng.pack(txt), [31, 32], p=10)
        for i, k in enumerate((31, 32)):
            assert np.array_equal(regs[i].cpu().numpy(), orc.hll_sketch(want, k, 10))
        check(eng.lib.dd_set_option(b"polyt_sentinel", 1))
        hregs, _ = eng.sketch_fasta_host(txt, [32], p=10)
        assert np.array_equal(hregs[0], orc.hll_sketch(want, 32, 10))
    finally:
        eng.polyt_sentinel = False
        check(eng.lib.dd_set_option(b"polyt_sentinel", 0))
    hregs, _ = eng.sketch_fasta_host(txt, [32], p=10)
    assert np.array_equal(hregs[0], orc.hll_sketch(base, 32, 10))


# ------------------------------------------------------------------------------------- bit planes
def _sketch_set(eng, rng, n, ks, p, length=30000, sub=0.05):
    import torch
    anc = random_bases(rng, length)
    regs = []
    for g in range(n):
        txt = to_fasta([(b"g%d" % g, mutate(rng, anc, sub=sub))])
        regs.append(eng.sketch(eng.pack(txt), ks, p=p)[0])
    return torch.stack(regs).contiguous()


@pytest.mark.parametrize("p", [12, 16])
def test_planes_entry_points(eng, p):
    """dd_to_planes once, then prefix unions and pair unions on the planes == the register-input
    entries == the oracle estimator on numpy maxima."""
    rng = np.random.default_rng(20 + p)
    ks, n = [11, 17, 31], 7
    regs = _sketch_set(eng, rng, n, ks, p)
    h = regs.cpu().numpy()
    planes = eng.to_planes(regs, p)
    assert planes.numel() * 4 == eng.lib.dd_planes_bytes(n * len(ks), p)
    orders = [list(range(n)), [6, 5, 4, 3, 2, 1, 0], [3, -1, 3, 0, -1, 6, 2], list(range(n))]
    a = eng.prefix_union_cards(regs, orders, p).cpu().numpy()
    b = eng.prefix_union_cards_planes(planes, n, len(ks), orders, p).cpu().numpy()
    assert np.array_equal(a, b)
    for o, order in enumerate(orders):
        run = np.zeros_like(h[0])
        for s, g in enumerate(order):
            if g >= 0:
                run = np.maximum(run, h[g])
            for i in range(len(ks)):
                assert b[o, s, i] == pytest.approx(orc.card(run[i], p), rel=CARD_RTOL)
    pairs = [(x, y) for x in range(n) for y in range(x + 1, n)]
    c = eng.pairwise_cards(regs, pairs, p).cpu().numpy()
    d = eng.pairwise_cards(None, pairs, p, planes=planes, n_genomes=n, nk=len(ks)).cpu().numpy()
    assert np.array_equal(c, d)
    for j, (x, y) in enumerate(pairs):
        for i in range(len(ks)):
            assert d[j, i] == pytest.approx(orc.card(np.maximum(h[x, i], h[y, i]), p), rel=CARD_RTOL)


def test_prefix_union_without_scratch_uses_byte_kernel(eng):
    """d_ws == NULL is legal: the byte kernel needs no scratch and gives the same numbers."""
    import torch
    from dandd_b200._lib import check
    rng = np.random.default_rng(31)
    ks, n, p = [13, 21], 5, 14
    regs = _sketch_set(eng, rng, n, ks, p)
    orders = np.array([[0, 1, 2, 3, 4], [4, 2, 0, 1, 3]], dtype=np.int32)
    want = eng.prefix_union_cards(regs, orders, p).cpu().numpy()
    order = torch.from_numpy(orders).to(eng.device)
    hist = torch.empty((2, n, len(ks), 64), dtype=torch.int32, device=eng.device)
    cards = torch.empty((2, n, len(ks)), dtype=torch.float64, device=eng.device)
    check(eng.lib.dd_prefix_union_card(regs.data_ptr(), order.data_ptr(), 2, n, n, len(ks), p, 0, cards.data_ptr(),
                                       hist.data_ptr(), None, None, 0, eng.stream))
    assert np.array_equal(cards.cpu().numpy(), want)
    small = torch.empty(1024, dtype=torch.uint8, device=eng.device)
    rc = eng.lib.dd_prefix_union_card(regs.data_ptr(), order.data_ptr(), 2, n, n, len(ks), p, 0, cards.data_ptr(),
                                      hist.data_ptr(), None, small.data_ptr(), small.numel(), eng.stream)
    assert rc == -3 and b"workspace" in eng.lib.dd_last_error()


def test_store_pair_table_serves_two_leaf_unions(eng, tmp_path):
    """GpuSketchStore.pair_unions (one K6 job) then union_sketches for a pair: no new launch, the
    same cardinality as the direct union, and the file DandD expects exists."""
    from dandd_b200 import hllfile
    from dandd_b200.store import GpuSketchStore
    rng = np.random.default_rng(40)
    ks, p, n = [14, 15, 16], 12, 5
    regs = _sketch_set(eng, rng, n, ks, p).cpu().numpy()
    store = GpuSketchStore(engine=eng, union_files="stub")
    leaf_paths = {k: [] for k in ks}
    for g in range(n):
        for i, k in enumerate(ks):
            path = str(tmp_path / f"k{k}" / f"g{g}.hll")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            hllfile.write_hll(path, regs[g, i], p, 0.0)
            leaf_paths[k].append(path)
    table = store.pair_unions(leaf_paths, p)
    assert table.shape == (n * (n - 1) // 2, len(ks))
    launches = store.stats["union_launches"]
    out = {k: str(tmp_path / f"u{k}.hll") for k in ks}
    got = store.union_sketches({k: [leaf_paths[k][3], leaf_paths[k][1]] for k in ks}, p, out)   # order reversed on purpose
    assert store.stats["union_launches"] == launches
    for i, k in enumerate(ks):
        assert got[k] == pytest.approx(orc.card(np.maximum(regs[1, i], regs[3, i]), p), rel=CARD_RTOL)
        assert os.path.getsize(out[k]) > 0
        assert np.array_equal(store.registers(out[k]).cpu().numpy(), np.maximum(regs[1, i], regs[3, i]))   # stub rebuilds
    # three members: not a pair, goes through the kernel
    store.union_sketches({14: leaf_paths[14][:3]}, p, {14: str(tmp_path / "u3.hll")})
    assert store.stats["union_launches"] == launches + 1


# ------------------------------------------------------------------------- parity at config sizes
def _config2_genome(seed):
    rng = np.random.default_rng(seed)
    anc = random_bases(rng, 5_000_000)
    return to_fasta([(b"g", mutate(rng, anc))], width=80)


def test_config2_every_k_bit_exact(eng):
    """Config 2: one 5 Mbp genome, p = 20, EVERY k of 10..32 against the oracle (registers bit-exact,
    cardinalities 1e-9), device-resident path and host-buffer C ABI alike."""
    txt = _config2_genome(22)
    ks = list(range(10, 33))
    sym = orc.fasta_symbols(txt)
    with ThreadPoolExecutor(THREADS) as ex:
        want = list(ex.map(lambda k: orc.hll_sketch(sym, k, 20), ks))
    regs, cards = eng.sketch(eng.pack(txt), ks, p=20)
    hregs, hcards = eng.sketch_fasta_host(txt, ks, p=20)
    regs = regs.cpu().numpy()
    for i, k in enumerate(ks):
        assert np.array_equal(regs[i], want[i]), k
        assert np.array_equal(hregs[i], want[i]), k
        c = orc.card(want[i], 20)
        assert float(cards[i]) == pytest.approx(c, rel=CARD_RTOL)
        assert hcards[i] == pytest.approx(c, rel=CARD_RTOL)


def test_config4_exact_counts_100mbp(eng):
    """Config 4 size: two 100 Mbp genomes (1 % substitutions apart), k in {12, 20, 32}: the GPU count of
    each genome and of their union equals the oracle's sort-unique count."""
    rng = np.random.default_rng(44)
    anc = random_bases(rng, 100_000_000)
    mut = anc.copy()
    hit = rng.random(mut.size) < 0.01
    mut[hit] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(hit.sum()))]
    txts = [to_fasta([(b"a", anc)], width=100), to_fasta([(b"b", mut)], width=100)]
    del anc, mut, hit
    syms = [orc.fasta_symbols(t) for t in txts]
    seqs = [eng.pack(t) for t in txts]
    jobs = [(k, sel) for k in (12, 20, 32) for sel in ((0,), (1,), (0, 1))]
    with ThreadPoolExecutor(min(THREADS, len(jobs))) as ex:
        want = dict(zip(jobs, ex.map(lambda j: orc.exact_count([syms[i] for i in j[1]], j[0]), jobs)))
    for k in (12, 20, 32):
        a_then_ab = eng.exact_counts(seqs, k)
        b_alone = eng.exact_counts([seqs[1]], k)
        assert a_then_ab == [want[(k, (0,))], want[(k, (0, 1))]], k
        assert b_alone == [want[(k, (1,))]], k


def test_config5_every_pair_every_k(eng):
    """Config 5 shape at test size: 40 genomes x 5 Mbp (8 clusters of 5), p = 18, k = 10..32 -- every
    leaf sketch bit-exact, and the union cardinality of EVERY pair at EVERY k (780 x 23) equal to the
    oracle estimator on the numpy maximum of the oracle's registers."""
    import torch
    p, ks, n = 18, list(range(10, 33)), 40
    rng = np.random.default_rng(55)
    txts = []
    for c in range(8):
        anc = random_bases(rng, 5_000_000)
        for _ in range(5):
            s = anc.copy()
            hit = rng.random(s.size) < 0.02
            s[hit] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(hit.sum()))]
            txts.append(to_fasta([(b"g", s)], width=80))
    regs = torch.empty((n, len(ks), 1 << p), dtype=torch.uint8, device=eng.device)
    for g, t in enumerate(txts):
        eng.sketch(eng.pack(t), ks, p=p, out=regs[g])

    def oracle_genome(t):
        sym = orc.fasta_symbols(t)
        return np.stack([orc.hll_sketch(sym, k, p) for k in ks])
    with ThreadPoolExecutor(THREADS) as ex:
        want = np.stack(list(ex.map(oracle_genome, txts)))
    assert np.array_equal(regs.cpu().numpy(), want)
    pairs = [(a, b) for a in range(n) for b in range(a + 1, n)]
    planes = eng.to_planes(regs, p)
    got = eng.pairwise_cards(None, pairs, p, planes=planes, n_genomes=n, nk=len(ks)).cpu().numpy()

    def oracle_pair(ab):
        u = np.maximum(want[ab[0]], want[ab[1]])
        return [orc.card(u[i], p) for i in range(len(ks))]
    with ThreadPoolExecutor(THREADS) as ex:
        exp = np.array(list(ex.map(oracle_pair, pairs)))
    assert got.shape == exp.shape == (780, 23)
    assert np.allclose(got, exp, rtol=CARD_RTOL, atol=0)


# ------------------------------------------------------------------------------------- streaming
def test_streamed_file_sketch_matches_oracle(eng, tmp_path):
    """dandd_b200/streaming.py: disk -> pinned ring -> H2D -> K1 -> K2 in small chunks == the oracle on
    the whole file, and the blake2b it computes from the chunks == hashlib on the file; plain, gzip
    and prefetched-bytes sources; FASTQ is reported, not mis-sketched."""
    import gzip
    import hashlib
    from dandd_b200 import streaming
    from dandd_b200.engine import FastqInput
    rng = np.random.default_rng(77)
    txt = b"preamble junk\n" + adversarial_fasta(rng, n=3_000_000) + kseq_fasta(rng, n=500_000)
    path = tmp_path / "big.fa"
    path.write_bytes(txt)
    sym = orc.fasta_symbols(txt)
    ks = [9, 14, 21, 32]
    want = [orc.hll_sketch(sym, k, 16) for k in ks]
    for chunk in (1 << 20, 64 << 20):
        regs, cards, digest, stats = streaming.sketch_file(eng, str(path), ks, p=16, chunk_bytes=chunk)
        assert digest == hashlib.blake2b(txt).hexdigest()
        assert stats["chunks"] == (-(-len(txt) // (1 << 20)) if chunk == 1 << 20 else 1)
        for i in range(len(ks)):
            assert np.array_equal(regs[i].cpu().numpy(), want[i]), (chunk, ks[i])
            assert cards[i] == pytest.approx(orc.card(want[i], 16), rel=CARD_RTOL)
    gz = tmp_path / "big.fa.gz"
    gz.write_bytes(gzip.compress(txt, 1))
    regs, cards, digest, _ = streaming.sketch_file(eng, str(gz), ks, p=16, chunk_bytes=1 << 20)
    assert digest == hashlib.blake2b(gz.read_bytes()).hexdigest()
    assert all(np.array_equal(regs[i].cpu().numpy(), want[i]) for i in range(len(ks)))
    regs, cards, digest, _ = streaming.sketch_file(eng, str(path), ks, p=16, chunk_bytes=1 << 20, text=txt)
    assert digest is None and all(np.array_equal(regs[i].cpu().numpy(), want[i]) for i in range(len(ks)))
    fq = tmp_path / "reads.fq"
    fq.write_bytes(kseq_fasta(rng, n=2_000_000, fastq=True))
    with pytest.raises(FastqInput):
        streaming.sketch_file(eng, str(fq), ks, p=16, chunk_bytes=1 << 20)


def test_streaming_survives_a_failing_chunk(eng, tmp_path, monkeypatch):
    """A C-ABI error in the middle of a stream (injected: the third dd_sketch_update_sched call reports
    DD_ERR_ARG) is raised to the caller, every pinned slot returns to the ring, the reader and hasher
    threads end, and the next file streams through the same ring bit-identically to the oracle."""
    import threading
    from dandd_b200 import streaming
    rng = np.random.default_rng(79)
    txt = to_fasta([(b"s", random_bases(rng, 6_000_000))], width=70)
    path = tmp_path / "f.fa"
    path.write_bytes(txt)
    ks = [11, 31]
    real = eng.lib.dd_sketch_update_sched
    calls = {"n": 0}

    def flaky(*a):
        calls["n"] += 1
        return -1 if calls["n"] == 3 else real(*a)

    before = threading.active_count()
    monkeypatch.setattr(eng.lib, "dd_sketch_update_sched", flaky, raising=False)
    with pytest.raises(RuntimeError):
        streaming.sketch_file(eng, str(path), ks, p=14, chunk_bytes=1 << 20)
    monkeypatch.setattr(eng.lib, "dd_sketch_update_sched", real, raising=False)
    assert threading.active_count() == before
    ring = streaming._ring_for(eng.device, 1 << 20)
    assert ring.free.qsize() == len(ring.slots)
    regs, cards, digest, stats = streaming.sketch_file(eng, str(path), ks, p=14, chunk_bytes=1 << 20)
    sym = orc.fasta_symbols(txt)
    for i, k in enumerate(ks):
        assert np.array_equal(regs[i].cpu().numpy(), orc.hll_sketch(sym, k, 14))


def test_store_streams_large_fastas(eng, tmp_path, monkeypatch):
    """GpuSketchStore.leaf_sketches takes the streaming path above the size threshold, registers the
    digest for the naming layer, and falls back to the whole-file FASTQ detour when needed."""
    import hashlib
    from dandd_b200 import hllfile, ingest, store as store_mod
    monkeypatch.setattr(store_mod, "STREAM_MIN_BYTES", 1 << 20)
    rng = np.random.default_rng(78)
    texts = {"a.fa": to_fasta([(b"a", random_bases(rng, 2_500_000))], width=80),
             "q.fq": kseq_fasta(rng, n=1_500_000, fastq=True)}
    st = store_mod.GpuSketchStore(engine=eng, prefetch_all_k=False)
    for name, txt in texts.items():
        path = tmp_path / name
        path.write_bytes(txt)
        out = {k: str(tmp_path / "db" / f"k{k}" / (name + ".hll")) for k in (12, 31)}
        cards = st.leaf_sketches(str(path), [12, 31], 14, True, out)
        sym = orc.fasta_symbols(txt)
        for k in (12, 31):
            want = orc.hll_sketch(sym, k, 14)
            assert np.array_equal(hllfile.read_hll(out[k])[0], want), (name, k)
            assert cards[k] == pytest.approx(orc.card(want, 14), rel=CARD_RTOL)
        assert ingest.digest(str(path)) == hashlib.blake2b(txt).hexdigest()
    assert len(st.stream_stats) == 1          # the FASTA streamed; the FASTQ took the detour


# ------------------------------------------------------------------------------------- K2, long streams
def test_midk_bitmaps_long_stream(eng):
    """Pieces that start >= 2^24 symbols into a stream consult the L2-resident presence bitmaps for
    k = 10..12 (sketch.cu, kMidK) and the floor schedule has refreshed several times by then: registers
    for every k must still be bit-identical to the oracle, with the knob on and off, for the k = 2..32 and
    k = 1..32 static variants, canonical and not."""
    from dandd_b200._lib import check
    rng = np.random.default_rng(2024)
    n = 21_000_000
    seq = random_bases(rng, n)
    seq[rng.random(n) < 0.3] |= 0x20                      # soft-masked
    for st in rng.integers(0, n - 2000, 40):
        seq[st:st + 1500] = ord("N")
    seq[18_000_000:18_000_600] = ord("T")                 # low complexity inside the mid-k region
    txt = to_fasta([(b"chr%d" % i, seq[i * (n // 3):(i + 1) * (n // 3)]) for i in range(3)], width=80)
    sym = orc.fasta_symbols(txt)
    p = 14
    packed = eng.pack(txt)
    for ks, canon in ((list(range(2, 33)), True), (list(range(1, 33)), False)):
        check_ks = [2, 9, 10, 11, 12, 13, 32] if canon else [1, 10, 12, 31]
        with ThreadPoolExecutor(THREADS) as ex:
            want = dict(zip(check_ks, ex.map(lambda k: orc.hll_sketch(sym, k, p, canon), check_ks)))
        for knob in (1, 0):
            check(eng.lib.dd_set_option(b"sketch_midk", knob))
            try:
                regs, cards = eng.sketch(packed, ks, p=p, canon=canon)
            finally:
                check(eng.lib.dd_set_option(b"sketch_midk", 1))
            regs = regs.cpu().numpy()
            for k in check_ks:
                assert np.array_equal(regs[ks.index(k)], want[k]), (k, canon, knob)
    hregs, _ = eng.sketch_fasta_host(txt, list(range(2, 33)), p=p)     # chunked host path: 32 MiB text chunks
    for k in (10, 11, 12, 32):
        assert np.array_equal(hregs[k - 2], orc.hll_sketch(sym, k, p)), k


def test_long_stream_schedule_equals_plain_pass(eng):
    """300 Mbp (past 4^13 and 4^14 symbols, so the k = 13 and k = 14 presence bitmaps and several floor
    refreshes are in play): the scheduled path must give exactly the registers of one plain pass with
    no floor and no bitmaps (which the smaller tests pin to the oracle), for both wide static k sets;
    and a second sketch in the SAME workspace must not see the first one's bitmaps."""
    import torch
    from tools.synth import synth_fasta
    text = synth_fasta(300_000_000, 6, seed=11, device=eng.device)
    seq = eng.pack(text, start=0)
    n = seq.nsym
    for ks in (list(range(2, 33)), list(range(1, 33))):
        sched, _ = eng.sketch(seq, ks, p=12)
        plain, _ = eng.sketch(seq, ks, p=12, ranges=[(0, n)])
        assert bool((sched == plain).all()), ks[0]
    other = synth_fasta(80_000_000, 3, seed=12, device=eng.device)
    seq2 = eng.pack(other, start=0)
    a, _ = eng.sketch(seq2, list(range(2, 33)), p=12)                      # same "sketch" workspace as above
    b, _ = eng.sketch(seq2, list(range(2, 33)), p=12, ranges=[(0, seq2.nsym)])
    assert bool((a == b).all())
    del text, other
    torch.cuda.empty_cache()
