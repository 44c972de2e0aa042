// host_emul.cpp -- runs the bit-deciding device functions of dandd_b200/csrc/common.cuh on the CPU.
//
// common.cuh keeps every function that decides a result bit as host+device code.  This file wraps
// them in sequential stand-ins for the kernels' thread mapping (one "thread" per 16-byte text chunk
// for the packer, one per 16-symbol word for the sketch) so the CPU test-suite can compare exactly
// the shipped arithmetic with the oracle without a GPU.  It is compiled by tests/conftest.py with
// g++; it is test infrastructure, not a CPU fallback -- the product library links none of it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../dandd_b200/csrc/common.cuh"

using namespace dd;

extern "C" {

// Packer emulation: returns the number of symbols; codes/invalid must be zero-filled and sized
// for n symbols.  `entry_last` / `entry_hdr` are the carried state (use '\n', 0 for a fresh file);
// the state after the chunk is written back.  Also cross-checks the transition-function algebra:
// returns (size_t)-1 if composing chunk functions disagrees with the sequential walk.
size_t emul_pack(const uint8_t *text, size_t n, uint32_t *codes, uint32_t *invalid, size_t sym_offset,
                 uint32_t *entry_last, uint32_t *entry_hdr) {
    uint32_t state = *entry_hdr;
    uint64_t composed = kXferIdentity;
    size_t nsym = 0;
    for (size_t off = 0; off < n; off += 16) {
        uint32_t w[4];
        for (int i = 0; i < 4; ++i) {
            uint32_t x = 0;
            for (int b = 0; b < 4; ++b) {
                const size_t q = off + 4 * i + b;
                x |= (uint32_t)(q < n ? text[q] : kPadByte) << (8 * b);
            }
            w[i] = x;
        }
        const ChunkMasks m = classify16(w[0], w[1], w[2], w[3]);
        const uint32_t prev = off == 0 ? *entry_last : text[off - 1];
        const bool ls = prev == '\n';
        const uint64_t f = chunk_xfer(m, ls);
        composed = xfer_compose(composed, f);
        const ChunkSyms cs = chunk_symbols(m, ls, state != 0);
        if ((uint32_t)popc32(cs.sym) != xfer_cnt(f, state) || cs.end_hdr != xfer_end(f, state)) return (size_t)-1;
        // stage the chunk's symbols as bytes, then pack them 4 at a time with the kernel's helpers
        uint8_t staged[16];
        int cnt = 0;
        for (int i = 0; i < 16; ++i)
            if ((cs.sym >> i) & 1u) staged[cnt++] = (uint8_t)(((m.codes >> (2 * i)) & 3u) | (((cs.brk >> i) & 1u) << 2));
        for (int i = 0; i < cnt; ++i) {
            const size_t s = sym_offset + nsym + i;
            // pack_codes4 / pack_breaks4 on a word holding just this symbol in lane (s % 4)
            const uint32_t lane = (uint32_t)(s & 3);
            const uint32_t word = (uint32_t)staged[i] << (8 * lane);
            const uint32_t c8 = pack_codes4(word), b4 = pack_breaks4(word);
            codes[s >> 4] |= c8 << (24 - 8 * ((s >> 2) & 3));
            invalid[s >> 5] |= b4 << (28 - 4 * ((s >> 2) & 7));
        }
        nsym += cnt;
        state = cs.end_hdr;
    }
    if (n) {
        if (xfer_cnt(composed, *entry_hdr) != (uint32_t)(nsym & 0xFFFFFFu) && nsym < 0xFFFFFFu) return (size_t)-1;
        if (xfer_end(composed, *entry_hdr) != state) return (size_t)-1;
        *entry_last = text[n - 1];
    }
    *entry_hdr = state;
    return nsym;
}

// Sketch emulation for one k: same per-word / per-symbol walk as sketch_allk_kernel.
void emul_sketch(const uint32_t *codes, const uint32_t *invalid, uint64_t sym_begin, uint64_t sym_end, int k, int p,
                 int canon, uint8_t *regs) {
    if (sym_end <= sym_begin) return;
    for (uint64_t w = sym_begin >> 4; (w << 4) < sym_end; ++w) {
        const uint64_t s0 = w << 4;
        const uint32_t w0 = codes[w], w1 = w >= 1 ? codes[w - 1] : 0u, w2 = w >= 2 ? codes[w - 2] : 0u;
        const uint64_t iw = w >> 1;
        const uint32_t i0 = invalid[iw], i1 = iw >= 1 ? invalid[iw - 1] : 0xffffffffu;
        const uint32_t r0 = revcomp_word(w0), r1 = revcomp_word(w1), r2 = revcomp_word(w2);
        const int j_lo = sym_begin > s0 ? (int)(sym_begin - s0) : 0;
        const int j_hi = sym_end - s0 < 16 ? (int)(sym_end - s0) : 16;
        for (int j = j_lo; j < j_hi; ++j) {
            const int run = valid_run(invalid_window(i0, i1, (uint32_t)(s0 & 31) + (uint32_t)j));
            if (run < k) continue;
            const Window win = window_at(w0, w1, w2, r0, r1, r2, j);
            const uint64_t h = wang64(kmer_value_rt(win, k, canon != 0));
            const uint32_t idx = hll_index(h, p), rank = hll_rank(h, p);
            if (regs[idx] < rank) regs[idx] = (uint8_t)rank;
        }
    }
}

// Wide exact mode (k = 33..64): same per-word walk as exact_insert_wide_kernel; writes the 128-bit
// value of every valid k-mer as (lo, hi) pairs and returns how many there were.
size_t emul_kmers_wide(const uint32_t *codes, const uint32_t *invalid, uint64_t sym_begin, uint64_t sym_end, int k,
                       int canon, uint64_t *out) {
    size_t cnt = 0;
    if (sym_end <= sym_begin) return 0;
    for (uint64_t w = sym_begin >> 4; (w << 4) < sym_end; ++w) {
        const uint64_t s0 = w << 4, iw = w >> 1;
        uint32_t c[5], I[5];
        for (int t = 0; t < 5; ++t) {
            c[t] = w >= (uint64_t)t ? codes[w - t] : 0u;
            I[t] = iw >= (uint64_t)t ? invalid[iw - t] : 0xffffffffu;
        }
        const int j_lo = sym_begin > s0 ? (int)(sym_begin - s0) : 0;
        const int j_hi = sym_end - s0 < 16 ? (int)(sym_end - s0) : 16;
        for (int j = j_lo; j < j_hi; ++j) {
            if (valid_run_long(I, (uint32_t)(s0 & 31) + (uint32_t)j) < k) continue;
            const U128 v = kmer128_at(c, j, k, canon != 0);
            out[2 * cnt] = v.lo;
            out[2 * cnt + 1] = v.hi;
            ++cnt;
        }
    }
    return cnt;
}

// The kernel computes rank/index from the two hash halves; keep that formulation honest too.
uint32_t emul_rank_split(uint64_t h, int p) {
    const uint32_t hi = (uint32_t)(h >> 32), lo = (uint32_t)h;
    const uint32_t rem_hi = hi & (0xffffffffu >> p);
    return (rem_hi ? (uint32_t)clz32(rem_hi) : 32u + (uint32_t)clz32(lo)) + 1u - (uint32_t)p;
}
uint32_t emul_index_split(uint64_t h, int p) { return (uint32_t)(h >> 32) >> (32 - p); }
uint32_t emul_rank(uint64_t h, int p) { return hll_rank(h, p); }
uint32_t emul_index(uint64_t h, int p) { return hll_index(h, p); }
uint64_t emul_wang(uint64_t x) { return wang64(x); }
double emul_mle(const uint32_t *hist64, int p) {
    uint32_t c[66];
    memcpy(c, hist64, 64 * sizeof(uint32_t));
    c[64] = c[65] = 0;
    return ertl_mle(c, p);
}

// k = 65..256 (exact mode): the multi-word k-mer ending at symbol s_end -> out[0..W), returns W; and the
// valid-run length used to decide whether such a k-mer exists there.
int emul_kmer_long(const uint32_t *codes, uint64_t s_end, int k, int canon, uint64_t *out) {
    uint64_t w[kLongWords];
    const int W = kmer_long_at(codes, s_end, k, canon != 0, w);
    for (int t = 0; t < W; ++t) out[t] = w[t];
    return W;
}
int emul_valid_run(const uint32_t *invalid, uint64_t s, int need) { return valid_run_upto(invalid, s, need); }

}  // extern "C"
