"""Second, independent restatement of the hot-path arithmetic in numpy (TEST INFRASTRUCTURE ONLY).

Deliberately written differently from dandd_oracle.c -- windowed/vectorised instead of rolling,
a plain-Python FASTA walker, an mpmath root-find of Ertl's ML equation next to the iterative
estimator -- so that agreement between the two is evidence about the *specification*
(SURVEY.md Appendix A/B), not about one implementation.  Parity with real Dashing/KMC stays
unpinned: neither binary nor source is available (see dandd_oracle.c header).
"""
import math

import numpy as np

M64 = (1 << 64) - 1


def fasta_symbols_py(text: bytes) -> np.ndarray:
    """Line-oriented FASTA/FASTQ reader (A.1), written as a state machine over LINES where the C
    oracle walks bytes.  Records open at the first '>' or '@' found anywhere while no record is
    open, afterwards at lines BEGINNING with '>' or '@'; a line beginning with '+' switches to
    quality lines, consumed until they hold as many bytes as the record's sequence; a record whose
    quality length differs ends the parse (kseq_read error); line terminators are not sequence."""
    lut = {ord(c): v for c, v in zip("ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3])}
    out = []
    pos, n = 0, len(text)

    def next_line(at):
        """(line bytes without the terminator, offset after it, whether a terminator was seen)"""
        end = text.find(b"\n", at)
        return (text[at:], n, False) if end < 0 else (text[at:end], end + 1, True)

    while True:
        marks = [m for m in (text.find(b">", pos), text.find(b"@", pos)) if m >= 0]
        if not marks:
            break
        _, pos, _ = next_line(min(marks) + 1)       # header: rest of the marker's line
        mark = len(out)
        out.append(4)
        while True:                                  # one iteration per record opened at a line start
            seq_len, kind = 0, None
            while pos < n:
                first = text[pos:pos + 1]
                if first in (b">", b"@", b"+"):
                    kind = first
                    break
                line, pos, _ = next_line(pos)
                for ch in line:
                    if ch != 13:                     # '\r'
                        out.append(lut.get(ch, 4))
                        seq_len += 1
            if kind in (b">", b"@"):
                _, pos, _ = next_line(pos + 1)
                mark = len(out)
                out.append(4)
                continue
            break
        if kind != b"+":
            break                                    # end of text after a FASTA record
        _, pos, terminated = next_line(pos)          # the '+' line
        if not terminated:
            del out[mark:]
            break
        qual_len, first_line = 0, True
        while pos < n and (first_line or qual_len < seq_len):
            line, pos, _ = next_line(pos)
            qual_len += len(line) - line.count(b"\r")
            first_line = False
        if qual_len != seq_len:
            del out[mark:]
            break
    return np.asarray(out, dtype=np.uint8)


def wang_np(x: np.ndarray) -> np.ndarray:
    """A.4 Thomas Wang 64-bit mix on a uint64 array (wrap-around arithmetic)."""
    k = x.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        k = (~k) + (k << np.uint64(21))
        k ^= k >> np.uint64(24)
        k = k * np.uint64(265)
        k ^= k >> np.uint64(14)
        k = k * np.uint64(21)
        k ^= k >> np.uint64(28)
        k = k + (k << np.uint64(31))
    return k


def kmers_np(sym: np.ndarray, k: int, canon: bool = True) -> np.ndarray:
    """All valid windows of length k (A.2/A.3), by position rather than by rolling."""
    sym = np.asarray(sym, dtype=np.uint8)
    n = sym.size
    if n < k:
        return np.zeros(0, dtype=np.uint64)
    bad = (sym > 3).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(bad)])
    ok = (cs[k:] - cs[:-k]) == 0                      # window [i, i+k) has no break
    codes = (sym & 3).astype(np.uint64)
    nwin = n - k + 1
    fwd = np.zeros(nwin, dtype=np.uint64)
    rc = np.zeros(nwin, dtype=np.uint64)
    for j in range(k):
        col = codes[j:j + nwin]
        fwd |= col << np.uint64(2 * (k - 1 - j))      # first base most significant
        rc |= (np.uint64(3) - col) << np.uint64(2 * j)  # complement, order reversed
    vals = np.minimum(fwd, rc) if canon else fwd
    return vals[ok]


def hll_registers_np(vals: np.ndarray, p: int) -> np.ndarray:
    """A.5: index = top p bits, rank = 1 + leading zeros of the remaining 64-p bits (capped)."""
    regs = np.zeros(1 << p, dtype=np.uint8)
    if vals.size == 0:
        return regs
    h = wang_np(vals)
    idx = (h >> np.uint64(64 - p)).astype(np.int64)
    q = 64 - p
    rest = h & np.uint64((1 << q) - 1)
    # rank = q - floor(log2(rest)) for rest>0, q+1 for rest==0; bit_length via Python ints (exact)
    bl = np.fromiter((int(v).bit_length() for v in rest.tolist()), dtype=np.int64, count=rest.size)
    rank = (q - bl + 1).astype(np.uint8)
    np.maximum.at(regs, idx, rank)
    return regs


def hll_sketch_np(sym, k, p=20, canon=True):
    return hll_registers_np(kmers_np(sym, k, canon), p)


def exact_count_np(syms, k, canon=True) -> int:
    """Appendix B: size of the union of the k-mer sets."""
    if k > 32:
        return exact_count_strings(syms, k, canon)
    s = set()
    for sym in syms:
        s.update(kmers_np(sym, k, canon).tolist())
    return len(s)


def exact_count_strings(syms, k, canon=True) -> int:
    """The same on Python strings (any k): a k-mer and its reverse complement are one element
    when canon.  Small inputs only."""
    comp = str.maketrans("0123", "3210")
    s = set()
    for sym in syms:
        txt = "".join("N" if c > 3 else str(int(c)) for c in sym)
        for i in range(len(txt) - k + 1):
            w = txt[i:i + k]
            if "N" in w:
                continue
            if canon:
                r = w.translate(comp)[::-1]
                if r < w:
                    w = r
            s.add(w)
    return len(s)


def ertl_mle_py(c, p: int) -> float:
    """A.9 in plain Python floats, transcribed from Ertl's paper (Algorithm 8 structure)."""
    q = 64 - p
    m = float(1 << p)
    c = [int(v) for v in c] + [0] * (q + 2 - len(c))
    if c[q + 1] == (1 << p):
        return math.inf
    kmin = next(j for j in range(q + 2) if c[j])
    kmax = next(j for j in range(q + 1, -1, -1) if c[j])
    kminp, kmaxp = max(1, kmin), min(q, kmax)
    z = 0.0
    for j in range(kmaxp, kminp - 1, -1):
        z = 0.5 * z + c[j]
    z = math.ldexp(z, -kminp)
    cp = c[q + 1] + (c[kmaxp] if q >= 1 else 0)
    a = z + c[0]
    mp_ = m - c[0]
    b = z + math.ldexp(c[q + 1], -q)
    x = mp_ / (0.5 * b + a) if b <= 1.5 * a else mp_ / b * math.log1p(b / a)
    gprev, dx, relerr = 0.0, x, 1e-2 / math.sqrt(m)
    while dx > x * relerr:
        kap = math.frexp(x)[1]
        xp = math.ldexp(x, -max(kmaxp + 1, kap + 2))
        xp2 = xp * xp
        h = xp - xp2 / 3.0 + (xp2 * xp2) * (1.0 / 45.0 - xp2 / 472.5)
        for _ in range(kap, kmaxp - 1, -1):
            hp = 1.0 - h
            h = (xp + h * hp) / (xp + hp)
            xp += xp
        g = cp * h
        for j in range(kmaxp - 1, kminp - 1, -1):
            hp = 1.0 - h
            h = (xp + h * hp) / (xp + hp)
            xp += xp
            g += c[j] * h
        g += x * a
        dx = dx * (g - mp_) / (gprev - g) if (gprev < g <= mp_) else 0.0
        x += dx
        gprev = g
    return x * m


def ertl_ml_root_mp(c, p: int, dps: int = 40) -> float:
    """Exact root of the ML equation (last formula) with mpmath -- a second opinion
    on the estimator that shares no code with the secant iteration:
       x*(C0 + sum_j Cj 2^-j) + sum_{j=1..q} Cj h(x/2^j) + C_{q+1} h(x/2^q) = m - C0,
       h(y) = 1 - y/(e^y - 1)."""
    import mpmath as mp
    mp.mp.dps = dps
    q = 64 - p
    m = 1 << p
    c = [int(v) for v in c] + [0] * (q + 2 - len(c))
    if c[0] == m:
        return 0.0

    def h(y):
        return 1 - y / mp.expm1(y) if y != 0 else mp.mpf(0)

    lin = mp.mpf(c[0]) + sum(mp.mpf(c[j]) / (1 << j) for j in range(1, q + 1)) + mp.mpf(c[q + 1]) / (1 << q)

    def f(x):
        s = x * lin
        for j in range(1, q + 1):
            if c[j]:
                s += c[j] * h(x / (1 << j))
        if c[q + 1]:
            s += c[q + 1] * h(x / (1 << q))
        return s - (m - c[0])

    lo, hi = mp.mpf(0), mp.mpf(1)
    while f(hi) < 0:
        hi *= 2
    root = mp.findroot(f, (lo, hi), solver="anderson", tol=mp.mpf(10) ** (-30), maxsteps=200)
    return float(root * m)
