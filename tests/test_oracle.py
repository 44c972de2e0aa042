"""The oracle against known answers and against its independent numpy twin.

There are no reference goldens to pin to (the reference has no tests and its
arithmetic lives in absent third-party binaries), so the pins are: hand-computable vectors, the
survey's check values, agreement of two independently written restatements, a high-precision
solve of Ertl's ML equation, and the committed fixtures under tests/golden/."""
import json
import math
import os

import numpy as np
import pytest

from oracle import pyoracle as orc
from oracle import ref_numpy as ref
from tests.util import adversarial_fasta, random_bases, to_fasta

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_wang_known_answers():
    # SURVEY.md A.4 check values
    assert orc.wang(0) == 0x77CFA1EEF01BCA90
    assert orc.wang(1) == 0x5BCA7C69B794F8CE
    xs = np.array([0, 1, 2, 0xFFFFFFFFFFFFFFFF, 0x0123456789ABCDEF], dtype=np.uint64)
    assert [orc.wang(int(x)) for x in xs] == [int(v) for v in ref.wang_np(xs)]


def test_revcomp_and_canonical_by_hand():
    # ACGT (00 01 10 11 = 0x1B) is its own reverse complement; AAAA <-> TTTT
    assert orc.revcomp(0x1B, 4) == 0x1B
    assert orc.revcomp(0x00, 4) == 0xFF
    sym = np.array([0, 1, 2, 3, 3, 3], dtype=np.uint8)  # ACGTTT
    # 3-mers: ACG(6)->min(6,CGT=27)=6, CGT->6, GTT(47)->min(47,AAC=1)=1, TTT(63)->AAA=0
    assert orc.kmers(sym, 3).tolist() == [6, 6, 1, 0]
    assert orc.kmers(sym, 3, canon=False).tolist() == [6, 27, 47, 63]


def test_register_rule_by_hand():
    # one k-mer -> one register: index = top p bits, rank = 1 + leading zeros of the rest
    sym = np.array([0, 0, 0, 0], dtype=np.uint8)
    h = orc.wang(0)
    p = 8
    regs = orc.hll_sketch(sym, 4, p)
    idx = h >> (64 - p)
    rest = h & ((1 << (64 - p)) - 1)
    rank = (64 - p) - rest.bit_length() + 1
    assert regs[idx] == rank and int(regs.sum()) == rank


def test_fasta_symbol_rules():
    s = orc.fasta_symbols(b"junk before\n>h1 x\nACGT\nacgt\r\nNN\n\n>h2\nTT>A\n>h3")
    #            break A C G T a c g t  N  N  break T T > A break
    assert s.tolist() == [4, 0, 1, 2, 3, 0, 1, 2, 3, 4, 4, 4, 3, 3, 4, 0, 4]
    assert orc.fasta_symbols(b"no header here\nACGT\n").size == 0
    assert orc.fasta_symbols(b"").size == 0
    # k-mers never span records or invalid characters; short records contribute nothing
    s = orc.fasta_symbols(b">a\nACG\n>b\nTTT\n")
    assert orc.kmers(s, 3, canon=False).tolist() == [6, 63]
    assert orc.kmers(s, 4).size == 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c_oracle_equals_numpy_twin(seed):
    rng = np.random.default_rng(seed)
    txt = adversarial_fasta(rng)
    s1, s2 = orc.fasta_symbols(txt), ref.fasta_symbols_py(txt)
    assert np.array_equal(s1, s2)
    for k in (1, 2, 7, 15, 16, 17, 24, 31, 32):
        for canon in (True, False):
            assert np.array_equal(orc.kmers(s1, k, canon), ref.kmers_np(s2, k, canon)), (k, canon)
        for p in (8, 12):
            assert np.array_equal(orc.hll_sketch(s1, k, p), ref.hll_sketch_np(s2, k, p)), (k, p)
        assert orc.exact_count([s1], k) == ref.exact_count_np([s2], k)


def test_union_is_sketch_of_concatenation():
    rng = np.random.default_rng(7)
    a = orc.fasta_symbols(to_fasta([(b"a", random_bases(rng, 5000))]))
    b = orc.fasta_symbols(to_fasta([(b"b", random_bases(rng, 4000))]))
    for k in (11, 21):
        ra, rb = orc.hll_sketch(a, k, 10), orc.hll_sketch(b, k, 10)
        both = orc.hll_sketch(b, k, 10, regs=orc.hll_sketch(a, k, 10))
        assert np.array_equal(orc.union_max([ra, rb]), both)
        assert orc.exact_count([a, b], k) == len(set(orc.kmers(a, k).tolist()) | set(orc.kmers(b, k).tolist()))


@pytest.mark.parametrize("p,n", [(8, 50), (10, 5000), (12, 3000), (12, 200000), (14, 1_000_000)])
def test_ertl_mle_matches_exact_root_and_truth(p, n):
    rng = np.random.default_rng(p * 1000 + n % 997)
    h = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    q = 64 - p
    idx = (h >> np.uint64(q)).astype(np.int64)
    rest = h & np.uint64((1 << q) - 1)
    bl = np.array([int(v).bit_length() for v in rest.tolist()])
    regs = np.zeros(1 << p, dtype=np.uint8)
    np.maximum.at(regs, idx, (q - bl + 1).astype(np.uint8))
    c = orc.hist(regs, p)
    est = orc.ertl_mle(c, p)
    assert est == pytest.approx(ref.ertl_mle_py(c, p), rel=1e-12)
    assert est == pytest.approx(ref.ertl_ml_root_mp(c, p), rel=1e-6)      # the north-star tolerance
    assert est == pytest.approx(n, rel=6 * 1.04 / math.sqrt(1 << p))     # and it estimates n


def test_mle_edge_histograms():
    p, q = 10, 54
    m = 1 << p
    c = np.zeros(66, dtype=np.uint32); c[0] = m
    assert orc.ertl_mle(c, p) == 0.0
    c = np.zeros(66, dtype=np.uint32); c[q + 1] = m
    assert math.isinf(orc.ertl_mle(c, p))
    c = np.zeros(66, dtype=np.uint32); c[0] = m - 1; c[1] = 1
    assert orc.ertl_mle(c, p) == pytest.approx(1.0, rel=1e-3)


def test_golden_fixtures():
    """tests/golden/oracle_vectors.json was produced by tests/golden/make_oracle_vectors.py from the
    numpy twin; the C oracle must reproduce it bit for bit (registers via their byte sums and
    histograms, cardinalities to 1e-12)."""
    with open(os.path.join(GOLD, "oracle_vectors.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        rng = np.random.default_rng(case["seed"])
        txt = adversarial_fasta(rng, n=case["n"])
        sym = orc.fasta_symbols(txt)
        assert int(sym.size) == case["nsym"] and int((sym == 4).sum()) == case["nbreak"]
        for row in case["rows"]:
            regs = orc.hll_sketch(sym, row["k"], case["p"], canon=row["canon"])
            assert orc.hist(regs, case["p"])[:len(row["hist"])].tolist() == row["hist"]
            assert int(np.dot(regs.astype(np.int64), np.arange(regs.size) % 251)) == row["checksum"]
            assert orc.card(regs, case["p"]) == pytest.approx(row["card"], rel=1e-12)
            assert orc.exact_count([sym], row["k"], canon=row["canon"]) == row["exact"]


def test_exact_count_wide_k_against_string_sets():
    """k = 33..64 in the C oracle (128-bit values) against sets of Python strings; the input holds a
    reverse-complemented copy of part of itself so that canonical and plain counts differ."""
    rng = np.random.default_rng(64)
    a = orc.fasta_symbols(to_fasta([(b"a", random_bases(rng, 2500))]))
    b = a.copy()
    b[::97] = (b[::97] + 1) % 4
    rc = (3 - a[500:1500])[::-1].copy()
    c = np.concatenate([b[:800], np.array([4], dtype=np.uint8), rc])
    for k in (33, 41, 64):
        plain = orc.exact_count([a, c], k, canon=False)
        canon = orc.exact_count([a, c], k, canon=True)
        assert plain == ref.exact_count_strings([a, c], k, canon=False)
        assert canon == ref.exact_count_strings([a, c], k, canon=True)
        assert canon < plain


# ---- kseq record rules ('@' headers, '+' quality sections) and the A.6 switch ---------------------
KSEQ_CASES = [
    # (text, expected symbols) -- hand-derived from klib kseq_read()
    (b"@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+r2\nJJJJ\n", [4, 0, 1, 2, 3, 4, 2, 2, 1, 1]),          # 4-line FASTQ
    (b"@r1\nACGT\nAC\n+\nII\n@III\n>x\nACGT\n", [4, 0, 1, 2, 3, 0, 1, 4, 0, 1, 2, 3]),          # multi-line; a quality line starting with '@'
    (b"xx@r1\nACGT\n+\nIIII\njunk>r2 c\nAAAA\n@r3\nCC\n", [4, 0, 1, 2, 3, 4, 0, 0, 0, 0, 4, 1, 1]),  # markers found mid-line after a FASTQ record
    (b"@r1\nACGT\n+\nII\n>r2\nAAAA\n", []),                 # quality longer than sequence: kseq_read fails, nothing is read
    (b">ok\nAC\n@r1\nACGT\n+\nII\n", [4, 0, 1]),            # ... but records before the bad one stand
    (b"@r1\nACGT\n+", []),                                     # no quality string
    (b"@r1\n+\n\n>r2\nACGT", [4, 4, 0, 1, 2, 3]),            # empty read, one (empty) quality line
    (b">a\nAC\n+\nxx\n>b\nGG\n", [4, 0, 1, 4, 2, 2]),        # '+' line inside a FASTA file
    (b">a\nACGT\n@b\nTTTT\n", [4, 0, 1, 2, 3, 4, 3, 3, 3, 3]),  # '@' line is a header
    (b">a\nAC@GT+A>C\n", [4, 0, 1, 4, 2, 3, 4, 0, 4, 1]),       # the same bytes in the middle of a line are sequence
]


@pytest.mark.parametrize("text,want", KSEQ_CASES)
def test_kseq_record_rules(text, want):
    assert orc.fasta_symbols(text).tolist() == want
    assert ref.fasta_symbols_py(text).tolist() == want


def test_kseq_rules_fuzz_c_vs_numpy():
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b"ACGTacgtN>@+\n\n\n\r I", dtype=np.uint8)
    for _ in range(3000):
        txt = alpha[rng.integers(0, alpha.size, int(rng.integers(0, 60)))].tobytes()
        assert orc.fasta_symbols(txt).tolist() == ref.fasta_symbols_py(txt).tolist(), txt


def test_polyt_sentinel_switch():
    sym = np.array([0] + [3] * 70 + [4] + [3] * 31 + [1] + [3] * 32, dtype=np.uint8)
    out = orc.polyt_sentinel(sym)
    broken = np.flatnonzero(out != sym).tolist()
    assert broken == [32, 64, 135]            # every 32nd T of a run; a run of 31 is untouched
    assert (out[broken] == 4).all()
    # with the switch on, a k=32 poly-T window never reaches the sketch
    assert orc.kmers(sym, 32, True).size > 0 and not np.any(orc.kmers(out, 32, False) == np.uint64(0xFFFFFFFFFFFFFFFF))
