/*
 * dandd_oracle.c -- CPU restatement of the arithmetic on DandD's sketch-and-count hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it, and
 * there only as the checker / the timed CPU baseline, never as a fallback for the CUDA path.
 *
 * PARITY UNPINNED.  The reference repository (jessicabonnie/dandd) contains no arithmetic: every
 * number on this path is produced by three external, un-vendored, un-pinned programs that are not
 * installed here and whose sources are absent from /root/reference:
 *     Dashing v1  (github.com/dnbaker/dashing, "latest binary release", reference README.md:17-22)
 *     KMC 3       (bioconda "kmc",                                       reference README.md:23-27)
 *     GNU parallel                                                       (reference README.md:28-32)
 * The reference also ships no tests, fixtures or golden vectors.  This
 * file therefore restates the *published algorithms* of those tools and
 * anchors on the reference's own call sites:
 *     dashing sketch  : lib/sketch_classes.py:351-366      -> orc_hll_sketch()
 *     dashing union   : lib/sketch_classes.py:368-373      -> orc_union_max()
 *     dashing card    : lib/sketch_classes.py:306-321      -> orc_hist() + orc_ertl_mle()
 *     dashing hll     : helpers/allpairs.py:32-35          -> orc_hll_sketch() on several inputs
 *     kmc / kmc_tools : lib/sketch_classes.py:389-399,434-465 -> orc_exact_*()
 * It is cross-checked against an independent numpy restatement (oracle/ref_numpy.py) and a
 * high-precision solve of Ertl's ML equation; both are in tests/test_oracle.py.
 *
 * Build:  make -C oracle        (gcc -O3 -march=x86-64-v3 -shared -fPIC; portable to the GPU box)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_BREAK 4u /* symbol value for "k-mer window must restart here" */

/* ------------------------------------------------------------------------------------------
 * A.1 sequence model.  FASTA / FASTQ text -> symbol stream.
 *   0..3 = A,C,G,T (either case); 4 = break.
 * Restates what kseq_read() (klib kseq.h, the parser inside both Dashing/bonsai and KMC's -fm
 * mode) hands to the k-mer encoder, step for step:
 *   - with no record pending, kseq scans forward for the next '>' OR '@' ANYWHERE in the text
 *     (`while ((c = ks_getc(ks)) != -1 && c != '>' && c != '@');`): everything before the first
 *     marker of a file is ignored, and so is everything between the end of a FASTQ record and the
 *     next marker -- even when that marker sits in the middle of a line;
 *   - the rest of the marker's line is the record name/comment: it contributes no bases and ends
 *     the previous record, so exactly one break symbol is emitted per record (k-mers never span
 *     records);
 *   - sequence lines follow; the FIRST byte of each line decides: '>' or '@' opens the next record,
 *     '+' starts the quality section of a FASTQ record, an empty line is skipped, anything else
 *     makes the whole line sequence ('>', '@' or '+' in the middle of a line are sequence bytes);
 *   - '\n' is never part of the sequence; '\r' is dropped (kseq strips the '\r' of a CRLF line
 *     end; a '\r' elsewhere does not occur in practice and is dropped here as well -- one rule);
 *   - every other sequence byte: ACGTacgt map to 0..3, anything else (N, IUPAC, blanks ...) is a
 *     break  (bonsai cstr_lut == -1  /  KMC "symbols other than ACGT break k-mers");
 *   - after a '+' line kseq reads whole lines as quality until it holds at least as many quality
 *     bytes as the record has sequence bytes (at least one line), whatever those lines start with;
 *     if the two lengths then differ, or the text ends right after the '+' line, kseq_read returns
 *     an error: the caller's `while (kseq_read(ks) >= 0)` loop stops, so that record and the rest
 *     of the file contribute nothing.
 * Returns the number of symbols written; out must hold n bytes (a record marker is >= 1 byte
 * long, so the stream can never be longer than the text).
 * ------------------------------------------------------------------------------------------ */
static inline unsigned orc_base_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return ORC_BREAK;
    }
}

size_t orc_fasta_symbols(const uint8_t *buf, size_t n, uint8_t *out) {
    size_t i = 0, o = 0;
    int pending = 0; /* kseq's last_char: a record marker has already been consumed */
    for (;;) {
        if (!pending) { /* scan for the next record marker, anywhere */
            while (i < n && buf[i] != '>' && buf[i] != '@') ++i;
            if (i >= n) break;
            ++i;
        }
        pending = 0;
        const size_t record_start = o;
        out[o++] = ORC_BREAK;
        while (i < n && buf[i] != '\n') ++i; /* name + comment: the rest of the marker's line */
        if (i < n) ++i;
        /* sequence lines */
        size_t seq_len = 0; /* kseq's seq.l: sequence bytes kept ('\r' excluded) */
        int c = -1;
        while (i < n) {
            c = buf[i];
            if (c == '>' || c == '+' || c == '@') { ++i; break; }
            c = -1;
            if (buf[i] == '\n') { ++i; continue; } /* empty line */
            while (i < n && buf[i] != '\n') {
                if (buf[i] != '\r') { out[o++] = (uint8_t)orc_base_code(buf[i]); ++seq_len; }
                ++i;
            }
            if (i < n) ++i;
        }
        if (c == '>' || c == '@') { pending = 1; continue; }
        if (c != '+') break; /* end of text: last FASTA record */
        /* FASTQ: skip the rest of the '+' line, then read quality lines */
        while (i < n && buf[i] != '\n') ++i;
        if (i >= n) { o = record_start; break; } /* error: no quality string */
        ++i;
        size_t qual_len = 0;
        do {
            if (i >= n) break; /* ks_getuntil2 < 0 at end of text */
            while (i < n && buf[i] != '\n') { if (buf[i] != '\r') ++qual_len; ++i; }
            if (i < n) ++i;
        } while (qual_len < seq_len);
        if (qual_len != seq_len) { o = record_start; break; } /* error: record and the rest are not read */
    }
    return o;
}

/* A.6 (UNVERIFIED, optional): bonsai's unwindowed encoder is recalled to test its 64-bit rolling
 * accumulator against all-ones to detect an invalid base; 32 consecutive T (code 3) make a
 * legitimate accumulator all-ones as well, so the 32nd T of such a run would be taken for an
 * invalid base and the accumulator reset.  This turns every 32nd T of a run of T into a break, in
 * place -- the behaviour dd_pack_polyt_sentinel emulates when switched on. */
void orc_polyt_sentinel(uint8_t *sym, size_t n) {
    unsigned run = 0;
    for (size_t i = 0; i < n; ++i) {
        if (sym[i] != 3) { run = 0; continue; }
        if (++run == 32) { sym[i] = ORC_BREAK; run = 0; }
    }
}

/* ------------------------------------------------------------------------------------------
 * A.4 hash: Thomas Wang's 64-bit mix (dnbaker/sketch hash.h WangHash, Dashing v1 default).
 * ------------------------------------------------------------------------------------------ */
uint64_t orc_wang(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

/* A.3 reverse complement of a k-mer held 2 bits/base, first base most significant. */
uint64_t orc_revcomp(uint64_t v, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; ++i) { r = (r << 2) | (3u - (v & 3u)); v >>= 2; }
    return r;
}

/* A.5 register update.  index = top p bits; rank = clz(((h<<1)|1) << (p-1)) + 1  in [1, 64-p+1]. */
static inline void orc_hll_add(uint8_t *regs, int p, uint64_t h) {
    uint64_t idx = h >> (64 - p);
    uint64_t t = ((h << 1) | 1u) << (p - 1);
    uint8_t rank = (uint8_t)(__builtin_clzll(t) + 1);
    if (regs[idx] < rank) regs[idx] = rank;
}

/* Enumerate the k-mers of a symbol stream the way the encoder does (A.2/A.3): a window becomes
 * valid after k consecutive non-break symbols; every valid window yields one value (canonical =
 * min(kmer, revcomp) as unsigned integers when canon != 0).  The callback style keeps HLL and the
 * exact counter on literally the same enumeration. */
typedef void (*orc_kmer_fn)(uint64_t v, void *ctx);

static void orc_for_each_kmer(const uint8_t *sym, size_t n, int k, int canon, orc_kmer_fn fn, void *ctx) {
    const uint64_t mask = (k == 32) ? ~UINT64_C(0) : ((UINT64_C(1) << (2 * k)) - 1);
    const int rcshift = 2 * (k - 1);
    uint64_t fwd = 0, rc = 0;
    int filled = 0;
    for (size_t i = 0; i < n; ++i) {
        unsigned c = sym[i];
        if (c > 3) { filled = 0; fwd = rc = 0; continue; }
        fwd = ((fwd << 2) | c) & mask;
        rc = (rc >> 2) | ((uint64_t)(3u - c) << rcshift);
        if (filled < k) ++filled;
        if (filled == k) fn(canon ? (fwd < rc ? fwd : rc) : fwd, ctx);
    }
}

struct orc_hll_ctx { uint8_t *regs; int p; };
static void orc_hll_cb(uint64_t v, void *ctx) {
    struct orc_hll_ctx *c = (struct orc_hll_ctx *)ctx;
    orc_hll_add(c->regs, c->p, orc_wang(v));
}

/* `dashing sketch -k K -S p [--no-canon]` on one symbol stream; accumulates into regs (2^p u8,
 * caller zero-initialises), so calling it on several streams gives `dashing hll` / a union. */
void orc_hll_sketch(const uint8_t *sym, size_t n, int k, int p, int canon, uint8_t *regs) {
    struct orc_hll_ctx c = {regs, p};
    orc_for_each_kmer(sym, n, k, canon, orc_hll_cb, &c);
}

/* A.8 `dashing union`: element-wise max. */
void orc_union_max(const uint8_t *const *in, int nin, size_t len, uint8_t *out) {
    memset(out, 0, len);
    for (int j = 0; j < nin; ++j)
        for (size_t i = 0; i < len; ++i)
            if (in[j][i] > out[i]) out[i] = in[j][i];
}

/* Register-value histogram: counts[j] = #registers equal to j, j = 0 .. 64-p+1 (array of 66). */
void orc_hist(const uint8_t *regs, int p, uint32_t *counts /*[66]*/) {
    memset(counts, 0, 66 * sizeof(uint32_t));
    size_t m = (size_t)1 << p;
    for (size_t i = 0; i < m; ++i) counts[regs[i] > 65 ? 65 : regs[i]]++;
}

/* ------------------------------------------------------------------------------------------
 * A.9 cardinality: Ertl's maximum-likelihood estimator (Dashing's default ERTL_MLE), following
 * the structure of Ertl's reference implementation as carried by dnbaker/sketch
 * (detail::ertl_ml_estimate, relerr = 1e-2).  c[j], j = 0..q+1, q = 64-p.
 * ------------------------------------------------------------------------------------------ */
double orc_ertl_mle(const uint32_t *c, int p) {
    const int q = 64 - p;
    const double m = ldexp(1.0, p);
    if ((double)c[q + 1] == m) return INFINITY;
    int kmin = 0, kmax = q + 1;
    while (c[kmin] == 0) ++kmin;
    while (kmax && c[kmax] == 0) --kmax;
    const int kminp = kmin > 1 ? kmin : 1;
    const int kmaxp = kmax < q ? kmax : q;
    double z = 0.0;
    for (int k = kmaxp; k >= kminp; --k) z = 0.5 * z + (double)c[k];
    z = ldexp(z, -kminp);
    double cprime = (double)c[q + 1];
    if (q >= 1) cprime += (double)c[kmaxp];
    const double a = z + (double)c[0];
    const double mprime = m - (double)c[0];
    double x;
    {
        const double b = z + ldexp((double)c[q + 1], -q);
        if (b <= 1.5 * a) x = mprime / (0.5 * b + a);
        else x = mprime / b * log1p(b / a);
    }
    double gprev = 0.0, dx = x;
    const double relerr = 1e-2 / sqrt(m);
    while (dx > x * relerr) {
        int kappam1;
        frexp(x, &kappam1);
        const int sh = (kmaxp + 1 > kappam1 + 2) ? kmaxp + 1 : kappam1 + 2;
        double xp = ldexp(x, -sh);
        const double xp2 = xp * xp;
        double h = xp - xp2 / 3.0 + (xp2 * xp2) * (1.0 / 45.0 - xp2 / 472.5);
        for (int k = kappam1; k >= kmaxp; --k) {
            const double hp = 1.0 - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
        }
        double g = cprime * h;
        for (int k = kmaxp - 1; k >= kminp; --k) {
            const double hp = 1.0 - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
            g += (double)c[k] * h;
        }
        g += x * a;
        if (gprev < g && g <= mprime) dx *= (g - mprime) / (gprev - g);
        else dx = 0.0;
        x += dx;
        gprev = g;
    }
    return x * m;
}

/* `dashing card --presketched`: histogram + MLE of one 2^p-register sketch. */
double orc_card(const uint8_t *regs, int p) {
    uint32_t c[66];
    orc_hist(regs, p, c);
    return orc_ertl_mle(c, p);
}

/* ------------------------------------------------------------------------------------------
 * Appendix B: KMC semantics.  `kmc -ci1 -cs2 -kK [-b] -fm` stores the set of distinct
 * (canonical unless -b) k-mers; `kmc_tools info` prints its size; `kmc_tools complex` with
 * "out = input1 + ... + inputN" is the set union.  The oracle is sort + unique over exactly the
 * enumeration used above, k <= 32.
 * ------------------------------------------------------------------------------------------ */
struct orc_vec { uint64_t *a; size_t n, cap; };
static void orc_vec_cb(uint64_t v, void *ctx) {
    struct orc_vec *vec = (struct orc_vec *)ctx;
    if (vec->n == vec->cap) {
        vec->cap = vec->cap ? vec->cap * 2 : (1u << 16);
        vec->a = (uint64_t *)realloc(vec->a, vec->cap * sizeof(uint64_t));
    }
    vec->a[vec->n++] = v;
}

/* LSD radix sort, 8 passes of 8 bits (skips passes whose digit is constant). */
static void orc_radix_sort(uint64_t *a, size_t n) {
    if (n < 2) return;
    uint64_t *tmp = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *src = a, *dst = tmp;
    for (int pass = 0; pass < 8; ++pass) {
        size_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        const int sh = pass * 8;
        for (size_t i = 0; i < n; ++i) cnt[((src[i] >> sh) & 0xff) + 1]++;
        int trivial = 0;
        for (int d = 0; d < 256; ++d) if (cnt[d + 1] == n) { trivial = 1; break; }
        if (trivial) continue;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; ++i) dst[cnt[(src[i] >> sh) & 0xff]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
    free(tmp);
}

/* k = 33..64: the same enumeration on 128-bit values (KMC accepts k up to 256; the GPU path and
 * this oracle stop at 64, two machine words).  Sorted with qsort -- sizes here are test sizes. */
typedef unsigned __int128 orc_u128;
static int orc_cmp128(const void *a, const void *b) {
    const orc_u128 x = *(const orc_u128 *)a, y = *(const orc_u128 *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
static uint64_t orc_exact_count_wide(const uint8_t *const *sym, const size_t *n, int nseq, int k, int canon) {
    const orc_u128 mask = (k == 64) ? ~(orc_u128)0 : ((((orc_u128)1) << (2 * k)) - 1);
    const int rcshift = 2 * (k - 1);
    size_t cap = 0, cnt = 0;
    for (int s = 0; s < nseq; ++s) cap += n[s];
    orc_u128 *a = (orc_u128 *)malloc((cap ? cap : 1) * sizeof(orc_u128));
    for (int s = 0; s < nseq; ++s) {
        orc_u128 fwd = 0, rc = 0;
        int filled = 0;
        for (size_t i = 0; i < n[s]; ++i) {
            unsigned c = sym[s][i];
            if (c > 3) { filled = 0; fwd = rc = 0; continue; }
            fwd = ((fwd << 2) | c) & mask;
            rc = (rc >> 2) | ((orc_u128)(3u - c) << rcshift);
            if (filled < k) ++filled;
            if (filled == k) a[cnt++] = canon ? (fwd < rc ? fwd : rc) : fwd;
        }
    }
    uint64_t distinct = 0;
    if (cnt) {
        qsort(a, cnt, sizeof(orc_u128), orc_cmp128);
        distinct = 1;
        for (size_t i = 1; i < cnt; ++i) distinct += (a[i] != a[i - 1]);
    }
    free(a);
    return distinct;
}

/* k = 65..256 (KMC's own limit): k-mers as STRINGS of k symbol bytes -- a formulation that shares
 * nothing with the packed multi-word arithmetic of the GPU path.  Canonical = the lexicographically
 * smaller of the k-mer and its reverse complement in A<C<G<T order (identical to the numeric min of
 * the 2-bit encodings, Appendix B).  Sorted with qsort + memcmp -- test sizes only. */
static int orc_long_k = 0;
static int orc_cmp_long(const void *a, const void *b) { return memcmp(a, b, (size_t)orc_long_k); }
static uint64_t orc_exact_count_long(const uint8_t *const *sym, const size_t *n, int nseq, int k, int canon) {
    size_t cap = 0, cnt = 0;
    for (int s = 0; s < nseq; ++s) cap += n[s];
    uint8_t *a = (uint8_t *)malloc((cap ? cap : 1) * (size_t)k);
    uint8_t *rc = (uint8_t *)malloc((size_t)k);
    for (int s = 0; s < nseq; ++s) {
        size_t run = 0;
        for (size_t i = 0; i < n[s]; ++i) {
            if (sym[s][i] > 3) { run = 0; continue; }
            if (++run < (size_t)k) continue;
            const uint8_t *fwd = sym[s] + i + 1 - k;
            const uint8_t *pick = fwd;
            if (canon) {
                for (int j = 0; j < k; ++j) rc[j] = (uint8_t)(3 - fwd[k - 1 - j]);
                if (memcmp(rc, fwd, (size_t)k) < 0) pick = rc;
            }
            memcpy(a + cnt * (size_t)k, pick, (size_t)k);
            ++cnt;
        }
    }
    uint64_t distinct = 0;
    if (cnt) {
        orc_long_k = k;
        qsort(a, cnt, (size_t)k, orc_cmp_long);
        distinct = 1;
        for (size_t i = 1; i < cnt; ++i) distinct += memcmp(a + i * (size_t)k, a + (i - 1) * (size_t)k, (size_t)k) != 0;
    }
    free(a);
    free(rc);
    return distinct;
}

/* Number of distinct (canonical) k-mers in the union of nseq symbol streams. */
uint64_t orc_exact_count(const uint8_t *const *sym, const size_t *n, int nseq, int k, int canon) {
    if (k > 64) return orc_exact_count_long(sym, n, nseq, k, canon);
    if (k > 32) return orc_exact_count_wide(sym, n, nseq, k, canon);
    struct orc_vec v = {0, 0, 0};
    for (int s = 0; s < nseq; ++s) orc_for_each_kmer(sym[s], n[s], k, canon, orc_vec_cb, &v);
    if (v.n == 0) { free(v.a); return 0; }
    orc_radix_sort(v.a, v.n);
    uint64_t distinct = 1;
    for (size_t i = 1; i < v.n; ++i) distinct += (v.a[i] != v.a[i - 1]);
    free(v.a);
    return distinct;
}

/* Convenience for tests: dump the enumerated k-mer values (returns count; out may be NULL). */
size_t orc_kmers(const uint8_t *sym, size_t n, int k, int canon, uint64_t *out, size_t cap) {
    struct orc_vec v = {0, 0, 0};
    orc_for_each_kmer(sym, n, k, canon, orc_vec_cb, &v);
    size_t cnt = v.n;
    if (out) memcpy(out, v.a, (cnt < cap ? cnt : cap) * sizeof(uint64_t));
    free(v.a);
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * Whole-file helpers used by the CPU baseline (one call == one `dashing sketch` process of the
 * reference: parse the FASTA, one k, one sketch, one cardinality).
 * ------------------------------------------------------------------------------------------ */
double orc_sketch_fasta(const uint8_t *buf, size_t n, int k, int p, int canon, uint8_t *regs) {
    uint8_t *sym = (uint8_t *)malloc(n ? n : 1);
    size_t ns = orc_fasta_symbols(buf, n, sym);
    memset(regs, 0, (size_t)1 << p);
    orc_hll_sketch(sym, ns, k, p, canon, regs);
    free(sym);
    return orc_card(regs, p);
}
