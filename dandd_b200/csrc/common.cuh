// common.cuh -- shared arithmetic of the DandD hot path kernels (sm_100a).
//
// Everything that decides a bit of the result lives here as DD_HD (host+device) inline functions
// so that tests/host_emul.cu can run exactly this code on the CPU against the oracle; the kernels
// add only thread mapping, shared-memory staging and atomics around it.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define DD_HD __host__ __device__ __forceinline__
#else
#define DD_HD inline
#endif

namespace dd {

// ---------------------------------------------------------------------------------------------
// bit helpers with host twins
// ---------------------------------------------------------------------------------------------
DD_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {  // low 32 bits of ((hi:lo) >> s), 0<=s<=32
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(lo, hi, s);
#else
    return s == 0 ? lo : (s >= 32 ? hi : ((lo >> s) | (hi << (32 - s))));
#endif
}
DD_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
DD_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
DD_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
DD_HD int ctz32(uint32_t x) {  // 32 for x == 0
#if defined(__CUDA_ARCH__)
    return x ? (__ffs((int)x) - 1) : 32;
#else
    return x ? __builtin_ctz(x) : 32;
#endif
}

DD_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {  // PRMT, selector nibbles 0..7
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((src >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}

// ---------------------------------------------------------------------------------------------
// Hash + register update (dnbaker/sketch WangHash and hll_t::add)
// ---------------------------------------------------------------------------------------------
// Thomas Wang's 64-bit mix.  The shift-add steps are written as multiplications by constants
// (~x + (x<<21) == x*(2^21-1) - 1, x + (x<<3) + (x<<8) == x*265, x + (x<<2) + (x<<4) == x*21,
// x + (x<<31) == x*(2^31+1)); on the GPU each becomes an IMAD.WIDE + IMAD pair on the FMA pipe,
// leaving the ALU pipe to the xor-shift steps.
DD_HD uint64_t wang64(uint64_t x) {
    x = x * 0x1FFFFFull + 0xFFFFFFFFFFFFFFFFull;
    x ^= x >> 24;
    x *= 265ull;
    x ^= x >> 14;
    x *= 21ull;
    x ^= x >> 28;
    x *= 0x80000001ull;
    return x;
}

// rank = clz(((h << 1) | 1) << (p - 1)) + 1, in [1, 64-p+1]
DD_HD uint32_t hll_rank(uint64_t h, int p) {
    const uint64_t t = (h << p) | (1ull << (p - 1));
    const uint32_t hi = (uint32_t)(t >> 32), lo = (uint32_t)t;
    return (uint32_t)(hi ? clz32(hi) : 32 + clz32(lo)) + 1u;
}
DD_HD uint32_t hll_index(uint64_t h, int p) { return (uint32_t)(h >> (64 - p)); }

// ---------------------------------------------------------------------------------------------
// Packed-stream window extraction (layout in include/dandd_b200.h).
// w0 = code word holding the current symbol, w1/w2 = the one/two words before it.
// ---------------------------------------------------------------------------------------------
// Reverse-complement of one 16-symbol word: symbol order reversed, each code c -> 3-c.
DD_HD uint32_t revcomp_word(uint32_t w) {
    uint32_t y = brev32(w);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return ~y;
}

struct Window {
    uint64_t fwd;  // the 32 symbols ending at the current one, oldest most significant
    uint64_t rc;   // their reverse complement: newest symbol (complemented) most significant
};

// j = position of the current symbol inside w0 (0..15); r* = revcomp_word(w*).
DD_HD Window window_at(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t r0, uint32_t r1, uint32_t r2, int j) {
    const uint32_t sf = 30u - 2u * (uint32_t)j;  // forward stream: (w2:w1:w0) >> sf
    const uint32_t sr = 2u * (uint32_t)j + 2u;   // reverse stream: (r0:r1:r2) >> sr
    Window w;
    w.fwd = ((uint64_t)funnel_r(w1, w2, sf) << 32) | funnel_r(w0, w1, sf);
    w.rc = ((uint64_t)funnel_r(r1, r0, sr) << 32) | funnel_r(r2, r1, sr);
    return w;
}

// k-mer ending at the current symbol: forward value, or min(forward, reverse complement).
template <int K>
DD_HD uint64_t kmer_value(const Window &w, bool canon) {
    const uint64_t mask = (K == 32) ? ~0ull : ((1ull << (2 * (K & 31))) - 1ull);
    const uint64_t f = w.fwd & mask;
    if (!canon) return f;
    const uint64_t r = w.rc >> (64 - 2 * K);
    return f < r ? f : r;
}
DD_HD uint64_t kmer_value_rt(const Window &w, int k, bool canon) {
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const uint64_t f = w.fwd & mask;
    if (!canon) return f;
    const uint64_t r = w.rc >> (64 - 2 * k);
    return f < r ? f : r;
}

// Invalid-bit window: bit b of the result = "symbol (s - b) is invalid", b = 0..31, given the
// invalid word holding s (i0) and the previous word (i1); sm = s % 32.
DD_HD uint32_t invalid_window(uint32_t i0, uint32_t i1, uint32_t sm) { return funnel_r(i0, i1, 31u - sm); }
// Number of valid symbols ending at s (saturates at 32).
DD_HD int valid_run(uint32_t invwin) { return ctz32(invwin); }

// ---- k = 33..64 (exact mode only): 128-bit k-mer values -----------------------------------------
struct U128 {
    uint64_t lo, hi;
};
DD_HD bool u128_eq(const U128 &a, const U128 &b) { return a.lo == b.lo && a.hi == b.hi; }
DD_HD bool u128_less(const U128 &a, const U128 &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
// reverse the order of the 32 two-bit symbols of x
DD_HD uint64_t rev2_64(uint64_t x) {
    const uint64_t r = ((uint64_t)brev32((uint32_t)x) << 32) | brev32((uint32_t)(x >> 32));
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}
// c[t] = packed word (w - t), t = 0..4; the k-mer (33 <= k <= 64) ends at symbol j of word w.
// Value = its symbols as a base-4 number, first symbol most significant; canonical = the smaller
// of that and the same for the reverse complement.
DD_HD U128 kmer128_at(const uint32_t (&c)[5], int j, int k, bool canon) {
    const uint32_t sh = 2u * (15u - (uint32_t)j);
    const uint32_t f0 = funnel_r(c[0], c[1], sh), f1 = funnel_r(c[1], c[2], sh), f2 = funnel_r(c[2], c[3], sh),
                   f3 = funnel_r(c[3], c[4], sh);
    U128 fwd = {(uint64_t)f0 | ((uint64_t)f1 << 32), (uint64_t)f2 | ((uint64_t)f3 << 32)};
    const int drop = 128 - 2 * k;  // 0..62 bits above the k-mer
    if (drop) fwd.hi &= ~(uint64_t)0 >> drop;
    if (!canon) return fwd;
    // complement (the cleared top bits turn into ones), reverse the 64 symbol slots, and shift the
    // slots that came from above the k-mer out at the bottom
    U128 rc = {rev2_64(~fwd.hi), rev2_64(~fwd.lo)};
    if (drop) {
        rc.lo = (rc.lo >> drop) | (rc.hi << (64 - drop));
        rc.hi >>= drop;
    }
    return u128_less(rc, fwd) ? rc : fwd;
}
// Number of valid symbols ending at symbol sm (0..31) of invalid word I[0]; I[t] = invalid word
// (iw - t), t = 0..4.  Saturates at 128.
DD_HD int valid_run_long(const uint32_t (&I)[5], uint32_t sm) {
    int run = 0;
    for (int t = 0; t < 4; ++t) {
        const uint32_t win = funnel_r(I[t], I[t + 1], 31u - sm);
        if (win) return run + ctz32(win);
        run += 32;
    }
    return run;
}

// ---- k = 65..256 (exact mode only): k-mers of up to eight 64-bit words ---------------------------
// out[0] is the LEAST significant word; the value is the k symbols as a base-4 number, first symbol
// most significant, i.e. the same ordering as the 64- and 128-bit forms (and as comparing the symbol
// strings lexicographically with A<C<G<T).  s_end = stream position of the k-mer's last symbol; the
// caller guarantees s_end + 1 >= k.  Returns the number of words W = ceil(2k / 64).
constexpr int kLongWords = 8;
DD_HD int kmer_long_at(const uint32_t *codes, uint64_t s_end, int k, bool canon, uint64_t (&out)[kLongWords]) {
    const int W = (2 * k + 63) / 64;
    uint64_t fwd[kLongWords], rcw[kLongWords];
    // window t = the 32 symbols ending at s_end - 32 t (symbols before the stream read as 0 and are masked / shifted out)
    for (int t = 0; t < W; ++t) {
        const int64_t s = (int64_t)s_end - 32 * (int64_t)t;
        uint32_t w0 = 0, w1 = 0, w2 = 0;
        int j = 0;
        if (s >= 0) {
            const uint64_t w = (uint64_t)s >> 4;
            j = (int)((uint64_t)s & 15);
            w0 = codes[w];
            w1 = w >= 1 ? codes[w - 1] : 0u;
            w2 = w >= 2 ? codes[w - 2] : 0u;
        }
        const Window win = window_at(w0, w1, w2, revcomp_word(w0), revcomp_word(w1), revcomp_word(w2), j);
        fwd[t] = s >= 0 ? win.fwd : 0ull;
        rcw[t] = s >= 0 ? win.rc : 0ull;
    }
    const int top_bits = 2 * k - 64 * (W - 1);          // bits of the k-mer in the most significant word, 2..64
    if (top_bits < 64) fwd[W - 1] &= (~0ull) >> (64 - top_bits);
    if (!canon) {
        for (int t = 0; t < W; ++t) out[t] = fwd[t];
        return W;
    }
    // reverse complement: [rc(window 0) | rc(window 1) | ...] with window 0 most significant, shifted right
    // by the 64 W - 2k bits that belong to symbols before the k-mer
    const int drop = 64 - top_bits;                     // 0..62
    uint64_t rc[kLongWords];
    for (int t = 0; t < W; ++t) {                       // rc word t (little endian) before the shift = rcw[W - 1 - t]
        const uint64_t lo = rcw[W - 1 - t];
        const uint64_t hi = t + 1 < W ? rcw[W - 2 - t] : 0ull;
        rc[t] = drop ? (lo >> drop) | (hi << (64 - drop)) : lo;
    }
    bool take_rc = false;
    for (int t = W - 1; t >= 0; --t)
        if (rc[t] != fwd[t]) {
            take_rc = rc[t] < fwd[t];
            break;
        }
    for (int t = 0; t < W; ++t) out[t] = take_rc ? rc[t] : fwd[t];
    return W;
}
// Number of valid symbols ending at stream position s, saturating at `need` (any k): walks the
// break-bit words backwards.  invalid word i covers symbols 32 i .. 32 i + 31, symbol s at bit 31 - s % 32.
DD_HD int valid_run_upto(const uint32_t *invalid, uint64_t s, int need) {
    int run = 0;
    int64_t iw = (int64_t)(s >> 5);
    uint32_t sm = (uint32_t)(s & 31);
    // bits of word iw at and below symbol s, symbol s in bit 0
    uint32_t win = invalid[iw] >> (31u - sm);
    int have = (int)sm + 1;
    for (;;) {
        if (win) return run + ctz32(win);
        run += have;
        if (run >= need) return run;
        if (--iw < 0) return run;                       // the stream starts here
        win = invalid[iw];                              // next older word: its newest symbol (31) already sits in bit 0
        have = 32;
    }
}

// ---------------------------------------------------------------------------------------------
// FASTA text classification for the packer.  16 text bytes -> bit masks
// (bit i <-> byte i).  A pad byte ('\r') is inert: never a symbol, never changes line state.
// Line rules follow klib's kseq_read(): the first byte of a line decides -- '>' or '@' opens a
// record (header line), '+' opens a FASTQ quality section.  Quality sections are not a finite-state
// matter (their length depends on the record's sequence length), so the packer only REPORTS a
// line-initial '+' (dd_pack_state.reserved bit DD_PACK_FLAG_FASTQ); such text goes through
// dd_fastq_to_fasta_host first.
// ---------------------------------------------------------------------------------------------
constexpr uint8_t kPadByte = 0x0D;

struct ChunkMasks {
    uint32_t nl;     // '\n'
    uint32_t cr;     // '\r'
    uint32_t gt;     // '>' or '@': a record marker when it is the first byte of a line (kseq)
    uint32_t plus;   // '+': first byte of a line => FASTQ quality section (flagged, never packed here)
    uint32_t acgt;   // one of ACGTacgt
    uint32_t codes;  // 2-bit code of byte i at bits [2i+1:2i] (meaningful where acgt is set)
};

// Bit 7 of each byte of the result is set iff that byte of x equals the byte replicated in pat.
DD_HD uint32_t bytes_eq(uint32_t x, uint32_t pat) {
    const uint32_t m = x ^ pat;
    return ~(((m & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m) & 0x80808080u;
}
// Gather bit 7 of bytes 0..3 into bits 0..3.
DD_HD uint32_t gather_bit7(uint32_t z) { return (((z >> 7) * 0x00204081u) >> 21) & 0xFu; }

DD_HD void classify_word(uint32_t x, int wi, ChunkMasks &m) {
    const int sh = 4 * wi;
    m.nl |= gather_bit7(bytes_eq(x, 0x0a0a0a0au)) << sh;
    m.cr |= gather_bit7(bytes_eq(x, 0x0d0d0d0du)) << sh;
    m.gt |= gather_bit7(bytes_eq(x, 0x3e3e3e3eu) | bytes_eq(x, 0x40404040u)) << sh;
    m.plus |= gather_bit7(bytes_eq(x, 0x2b2b2b2bu)) << sh;
    // ACGT test with one table lookup: the low three bits of 'A','C','T','G' are 1,3,4,7 -- all
    // different -- so PRMT with those bits as selectors fetches the only letter each byte could be
    // (filler 0x01 can never equal a byte whose low bits select it), and one compare finishes it.
    const uint32_t u = x & 0xdfdfdfdfu;  // fold case
    const uint32_t y = u & 0x07070707u;
    const uint32_t t = y | (y >> 4);                        // byte0 = b0|b1<<4, byte2 = b2|b3<<4
    const uint32_t sel = (t & 0xffu) | ((t >> 8) & 0xff00u);
    const uint32_t expect = byte_perm(0x43014101u, 0x47010154u, sel);   // [.,A,.,C | T,.,.,G]
    m.acgt |= gather_bit7(bytes_eq(u, expect)) << sh;
    // code = ((c >> 1) ^ (c >> 2)) & 3 : A->0 C->1 G->2 T->3 for either case
    const uint32_t c = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
    // bytes (b0,b1,b2,b3) 2-bit fields -> bits [1:0],[3:2],[5:4],[7:6]
    const uint32_t packed = ((c * 0x00041041u) >> 18) & 0xFFu;
    m.codes |= packed << (8 * wi);
}

DD_HD ChunkMasks classify16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    ChunkMasks m = {0, 0, 0, 0, 0, 0};
    classify_word(x0, 0, m);
    classify_word(x1, 1, m);
    classify_word(x2, 2, m);
    classify_word(x3, 3, m);
    return m;
}

// What a 16-byte chunk emits, given whether its first byte is at a line start and whether the
// text before it ended inside a header line.
struct ChunkSyms {
    uint32_t sym;     // bytes that emit a symbol (sequence characters + the '>' of each header)
    uint32_t brk;     // subset of sym whose symbol is a break (non-ACGT or header marker)
    uint32_t end_hdr; // 1 if the chunk ends inside a header line
};

DD_HD ChunkSyms chunk_symbols(const ChunkMasks &m, bool first_at_line_start, bool start_in_header) {
    // line starts: byte after each '\n', plus byte 0 if the previous text byte was '\n'
    const uint32_t ls = ((m.nl << 1) | (first_at_line_start ? 1u : 0u)) & 0xFFFFu;
    uint32_t hs = ls & m.gt;  // header lines opening inside this chunk
    // A chunk whose first byte starts a line cannot inherit a header.
    const uint32_t inherit = (start_in_header && !first_at_line_start) ? 1u : 0u;
    // Header bytes run from an opening (or bit 0 when inherited) to the next '\n'.  Subtracting the
    // opening bits from the newline bits turns each [open, newline) span into a run of ones; an
    // unterminated header borrows all the way up, setting bit 16 (=> ends in header).
    const uint32_t open = hs | inherit;
    const uint32_t span = (m.nl - open) & ~m.nl;
    const uint32_t hdr = span & 0xFFFFu;
    ChunkSyms r;
    r.end_hdr = (span >> 16) & 1u;
    const uint32_t seq = ~(m.nl | m.cr | hdr) & 0xFFFFu;
    r.sym = seq | hs;
    r.brk = hs | (seq & ~m.acgt);
    return r;
}

// The chunk as a transition function over the two possible incoming states, packed in a u64 so
// block scans can compose it with shuffles:
//   bits  0..23  symbols emitted if the text before the chunk ended in sequence (state 0)
//   bits 24..47  symbols emitted if it ended inside a header             (state 1)
//   bit  48/49   state after the chunk for incoming state 0 / 1
DD_HD uint64_t xfer_make(uint32_t cnt0, uint32_t cnt1, uint32_t e0, uint32_t e1) {
    return (uint64_t)cnt0 | ((uint64_t)cnt1 << 24) | ((uint64_t)e0 << 48) | ((uint64_t)e1 << 49);
}
DD_HD uint32_t xfer_cnt(uint64_t f, uint32_t s) { return (uint32_t)(f >> (24 * s)) & 0xFFFFFFu; }
DD_HD uint32_t xfer_end(uint64_t f, uint32_t s) { return (uint32_t)(f >> (48 + s)) & 1u; }
constexpr uint64_t kXferIdentity = (uint64_t)1 << 49;  // cnt 0/0, end0 = 0, end1 = 1
// f first, then g
DD_HD uint64_t xfer_compose(uint64_t f, uint64_t g) {
    const uint32_t e0 = xfer_end(f, 0), e1 = xfer_end(f, 1);
    return xfer_make(xfer_cnt(f, 0) + xfer_cnt(g, e0), xfer_cnt(f, 1) + xfer_cnt(g, e1), xfer_end(g, e0),
                     xfer_end(g, e1));
}
// bytes of the chunk that are a line-initial '+'
DD_HD uint32_t chunk_fastq_marks(const ChunkMasks &m, bool first_at_line_start) {
    return ((m.nl << 1) | (first_at_line_start ? 1u : 0u)) & 0xFFFFu & m.plus;
}
DD_HD uint64_t chunk_xfer(const ChunkMasks &m, bool first_at_line_start) {
    // both incoming states at once (same terms as chunk_symbols; only the "inherited header" bit differs)
    const uint32_t ls = ((m.nl << 1) | (first_at_line_start ? 1u : 0u)) & 0xFFFFu;
    const uint32_t hs = ls & m.gt;
    const uint32_t body = ~(m.nl | m.cr) & 0xFFFFu;
    const uint32_t span0 = (m.nl - hs) & ~m.nl;
    const uint32_t span1 = first_at_line_start ? span0 : ((m.nl - (hs | 1u)) & ~m.nl);
    return xfer_make((uint32_t)popc32((body & ~span0) | hs), (uint32_t)popc32((body & ~span1) | hs), (span0 >> 16) & 1u,
                     (span1 >> 16) & 1u);
}

// 16 staged symbol bytes (bits 1:0 code, bit 2 break; 4 per word, first symbol in the low byte)
// -> 8 bits of big-endian 2-bit codes / 4 break bits for one word of 4.
DD_HD uint32_t pack_codes4(uint32_t x) { return (((x & 0x03030303u) * 0x40100401u) >> 24) & 0xFFu; }
DD_HD uint32_t pack_breaks4(uint32_t x) { return ((((x >> 2) & 0x01010101u) * 0x08040201u) >> 24) & 0xFu; }

// ---------------------------------------------------------------------------------------------
// Cardinality from a register-value histogram: Ertl's maximum-likelihood estimator, the default
// of `dashing card`.  cnt[j] = #registers == j for j = 0..q+1, q = 64-p.
// The likelihood equation is solved by the secant iteration of Ertl's reference code, stopped at
// a relative step of 1e-2/sqrt(m), so the result agrees with Dashing's to rounding (<= 1e-12 rel.)
// rather than being "the" root.  Pure f64; one thread per sketch is plenty.
// ---------------------------------------------------------------------------------------------
DD_HD double ertl_mle(const uint32_t *cnt, int p) {
    const int q = 64 - p;
    const double m = ldexp(1.0, p);
    if ((double)cnt[q + 1] == m) return INFINITY;  // every register saturated
    int lo = 0, hi = q + 1;
    while (cnt[lo] == 0) ++lo;
    while (hi > 0 && cnt[hi] == 0) --hi;
    const int jlo = lo < 1 ? 1 : lo;   // smallest non-empty rank, at least 1
    const int jhi = hi > q ? q : hi;   // largest non-empty rank, at most q
    // z = sum_{j=jlo..jhi} cnt[j] 2^-j by Horner from the top, scaled once at the end
    double z = 0.0;
    for (int j = jhi; j >= jlo; --j) z = 0.5 * z + (double)cnt[j];
    z = ldexp(z, -jlo);
    double ctop = (double)cnt[q + 1];
    if (q >= 1) ctop += (double)cnt[jhi];
    const double a = z + (double)cnt[0];
    const double occupied = m - (double)cnt[0];
    const double b = z + ldexp((double)cnt[q + 1], -q);
    // starting point: lower bound of the root
    double x = (b <= 1.5 * a) ? occupied / (0.5 * b + a) : occupied / b * log1p(b / a);
    double step = x, g_prev = 0.0;
    const double tol = 1e-2 / sqrt(m);
    while (step > x * tol) {
        int e;
        (void)frexp(x, &e);  // x = f * 2^e, f in [0.5, 1)
        const int down = (jhi + 1 > e + 2) ? jhi + 1 : e + 2;
        double y = ldexp(x, -down);
        const double y2 = y * y;
        // h(t) = 1 - t/(e^t - 1) by its Taylor series at a tiny argument, then carried to each
        // larger power of two with the doubling identity  h <- (y + h(1-h)) / (y + (1-h)),  y <- 2y
        double h = y - y2 / 3.0 + (y2 * y2) * (1.0 / 45.0 - y2 / 472.5);
        for (int j = e; j >= jhi; --j) {
            const double hc = 1.0 - h;
            h = (y + h * hc) / (y + hc);
            y += y;
        }
        double g = ctop * h;
        for (int j = jhi - 1; j >= jlo; --j) {
            const double hc = 1.0 - h;
            h = (y + h * hc) / (y + hc);
            y += y;
            g += (double)cnt[j] * h;
        }
        g += x * a;
        if (g_prev < g && g <= occupied) step *= (g - occupied) / (g_prev - g);
        else step = 0.0;
        x += step;
        g_prev = g;
    }
    return x * m;
}

}  // namespace dd
