// sketch.cu -- K2: fused all-k canonical k-mer HyperLogLog sketch.
//
// Replaces `parallel -j 95% ' dashing sketch -k{} -S p ... ' ::: k...` (reference
// lib/huffman_dandd.py:217 over lib/sketch_classes.py:351-366): ONE pass over the packed symbol
// stream updates the registers of every requested k with Dashing's own hash (Wang) and register
// rule, so the registers are bit-identical to nk separate `dashing sketch` runs.
//
// Mapping: one thread per 16-symbol code word.  The 64-bit forward window and its reverse
// complement for any symbol of the word are two funnel shifts over three consecutive words (the
// reverse-complement words are computed once per thread), so there is no rolling state and no
// dependence between symbols or between k values; validity of a k-mer is "at least k symbols
// since the last break", one funnel shift + ctz on the break-bit words.  The k loop is unrolled
// at compile time (masks and shifts are immediates; k <= 16 keeps the k-mer in one register); the
// symbol loop is not, which keeps the body (~1.1k instructions for 31 k) inside the instruction
// cache.
//
// Registers are accumulated as one u32 per register with RED.MAX (there is no byte-wide atomic
// max); dd_sketch_end narrows them to u8.  The accumulators of one genome (nk x 4 MiB at p=20)
// are L2-resident on B200, so the scattered updates never reach HBM.  Bound: INT32 issue and the
// L2 scattered-update rate -- not HBM (see DESIGN.md for the roofline bookkeeping).
#include <cuda_runtime.h>

#include <utility>

#include "common.cuh"
#include "kernels.cuh"

namespace dd {

constexpr int kSketchThreads = 256;

struct SketchArgs {
    const uint32_t *codes;
    const uint32_t *invalid;
    const dd_pack_state *state;  // if non-null the symbol range is [state->prev_nsym, state->nsym)
    uint64_t sym_begin, sym_end;
    uint32_t kmask;
    int p;
    uint32_t *acc;                // [nk][2^p]
    const SketchWsHeader *hdr;
};

// 64-bit x times 32-bit constant: IMAD.WIDE.U32 + IMAD, both on the FMA pipe (PTX spelled out so
// that ptxas does not split the high-word multiply-add).
__device__ __forceinline__ uint64_t mul64x32(uint64_t x, uint32_t c) {
    uint32_t xl, xh, rl, rh;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(xl), "=r"(xh) : "l"(x));
    uint64_t r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(xl), "r"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(rl), "=r"(rh) : "l"(r));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rh) : "r"(xh), "r"(c), "r"(rh));
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(rl), "r"(rh));
    return r;
}
// ... plus a 64-bit addend folded into the wide multiply.
__device__ __forceinline__ uint64_t mad64x32(uint64_t x, uint32_t c, uint64_t add) {
    uint32_t xl, xh, rl, rh;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(xl), "=r"(xh) : "l"(x));
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(xl), "r"(c), "l"(add));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(rl), "=r"(rh) : "l"(r));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rh) : "r"(xh), "r"(c), "r"(rh));
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(rl), "r"(rh));
    return r;
}

// One (symbol, k) update.  Branch-free apart from the warp-uniform "is this k requested" test:
// an invalid window only predicates the RED off, so neighbouring k bodies can be interleaved.
template <int K, bool kCanon>
__device__ __forceinline__ void update_one_k(const Window &win, int run, uint32_t kmask, int p, uint32_t *acc,
                                             uint32_t &off_k, const uint32_t (&floor4)[8]) {
    if (!((kmask >> (K - 1)) & 1u)) return;  // warp-uniform
    const uint64_t v = kmer_value<K>(win, kCanon);
    // dd::wang64 (common.cuh) with the multiplications pinned to the FMA pipe
    uint64_t h = mad64x32(v, 0x1FFFFFu, 0xFFFFFFFFFFFFFFFFull);
    h ^= h >> 24;
    h = mul64x32(h, 265u);
    h ^= h >> 14;
    h = mul64x32(h, 21u);
    h ^= h >> 28;
    h = mul64x32(h, 0x80000001u);
    const uint32_t hi = (uint32_t)(h >> 32), lo = (uint32_t)h;
    // rank = 1 + leading zeros of the low (64-p) bits, capped at 64-p+1
    const uint32_t rem_hi = hi & (0xffffffffu >> p);
    const uint32_t rank = (rem_hi ? (uint32_t)__clz((int)rem_hi) : 32u + (uint32_t)__clz((int)lo)) + 1u - (uint32_t)p;
    const uint32_t floor_k = (floor4[(K - 1) >> 2] >> (8 * ((K - 1) & 3))) & 0xffu;
    if (run >= K && rank > floor_k) atomicMax(acc + (off_k + (hi >> (32 - p))), rank);
    off_k += 1u << p;
}

template <bool kCanon, int... Ks>
__device__ __forceinline__ void update_all_k(std::integer_sequence<int, Ks...>, const Window &win, int run,
                                             uint32_t kmask, int p, uint32_t *acc, const uint32_t (&floor4)[8]) {
    uint32_t off_k = 0;
    (update_one_k<Ks + 1, kCanon>(win, run, kmask, p, acc, off_k, floor4), ...);
}

template <bool kCanon>
__global__ void __launch_bounds__(kSketchThreads) sketch_allk_kernel(SketchArgs a) {
    // per-k floors (indexed by k-1), four to a register
    uint32_t floor4[8];
    {
        const uint4 f0 = __ldg(reinterpret_cast<const uint4 *>(a.hdr->floor));
        const uint4 f1 = __ldg(reinterpret_cast<const uint4 *>(a.hdr->floor) + 1);
        floor4[0] = f0.x; floor4[1] = f0.y; floor4[2] = f0.z; floor4[3] = f0.w;
        floor4[4] = f1.x; floor4[5] = f1.y; floor4[6] = f1.z; floor4[7] = f1.w;
    }

    uint64_t sym_begin = a.sym_begin, sym_end = a.sym_end;
    if (a.state) {
        sym_begin = a.state->prev_nsym;
        sym_end = a.state->nsym;
    }
    const uint64_t w = (sym_begin >> 4) + (uint64_t)blockIdx.x * kSketchThreads + threadIdx.x;
    const uint64_t s0 = w << 4;
    if (s0 >= sym_end) return;

    const uint32_t w0 = __ldg(a.codes + w);
    const uint32_t w1 = w >= 1 ? __ldg(a.codes + w - 1) : 0u;
    const uint32_t w2 = w >= 2 ? __ldg(a.codes + w - 2) : 0u;
    const uint64_t iw = w >> 1;
    const uint32_t i0 = __ldg(a.invalid + iw);
    const uint32_t i1 = iw >= 1 ? __ldg(a.invalid + iw - 1) : 0xffffffffu;  // before the stream: breaks
    const uint32_t r0 = revcomp_word(w0), r1 = revcomp_word(w1), r2 = revcomp_word(w2);
    const bool all_valid = (i0 | i1) == 0u;

    const int j_lo = sym_begin > s0 ? (int)(sym_begin - s0) : 0;
    const int j_hi = sym_end - s0 < 16 ? (int)(sym_end - s0) : 16;
    const uint32_t sm_base = (uint32_t)(s0 & 31);

#pragma unroll 1
    for (int j = j_lo; j < j_hi; ++j) {
        const Window win = window_at(w0, w1, w2, r0, r1, r2, j);
        const int run = all_valid ? 32 : valid_run(invalid_window(i0, i1, sm_base + (uint32_t)j));
        if (run == 0) continue;
        update_all_k<kCanon>(std::make_integer_sequence<int, 32>{}, win, run, a.kmask, a.p, a.acc, floor4);
    }
}

// ---- per-slot min(register): the "floor" filter ---------------------------------------------------
__global__ void __launch_bounds__(256) floor_min_kernel(const uint32_t *__restrict__ acc, int p, SketchWsHeader *hdr,
                                                        uint32_t *scratch /*[nk]*/) {
    // grid (slices, nk); scratch pre-set to 0xff
    const size_t m = (size_t)1 << p;
    const uint32_t *t = acc + (size_t)blockIdx.y * m;
    uint32_t mn = 0xffu;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m / 4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4 *>(t)[i];
        mn = min(mn, min(min(v.x, v.y), min(v.z, v.w)));
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0) atomicMin(&scratch[blockIdx.y], mn);
}
// scratch[slot] -> floor[k-1] for the slot-th set bit of kmask
__global__ void floor_publish_kernel(SketchWsHeader *hdr, const uint32_t *scratch, uint32_t kmask) {
    const int i = threadIdx.x;  // k - 1
    if (i < 32) {
        const bool on = (kmask >> i) & 1u;
        const int slot = __popc(kmask & ((1u << i) - 1u));
        hdr->floor[i] = on ? (uint8_t)scratch[slot] : 0;
    }
    if (i == 0) hdr->use_floor = 1;
}

// ---- finalisation: u32 accumulators -> u8 registers (+ histogram) -------------------------------
constexpr int kFinThreads = 128;
constexpr int kFinSlices = 8;  // CTAs per table

__global__ void __launch_bounds__(kFinThreads)
finalize_kernel(const uint32_t *__restrict__ acc, int p, uint8_t *__restrict__ regs, uint32_t *__restrict__ hist) {
    // thread-private histograms, [bin][thread] so that a warp's 32 increments hit 32 banks
    __shared__ uint32_t s_hist[DD_HIST_BINS * kFinThreads];
    const size_t m = (size_t)1 << p;
    const int table = blockIdx.y;
    const uint4 *src = reinterpret_cast<const uint4 *>(acc + (size_t)table * m);
    uint32_t *dst = reinterpret_cast<uint32_t *>(regs + (size_t)table * m);
    if (hist)
        for (int i = threadIdx.x; i < DD_HIST_BINS * kFinThreads; i += kFinThreads) s_hist[i] = 0;
    __syncthreads();
    const size_t nquads = m / 4;
    const size_t per = (nquads + gridDim.x - 1) / gridDim.x;
    const size_t q0 = (size_t)blockIdx.x * per, q1 = min(nquads, q0 + per);
    for (size_t q = q0 + threadIdx.x; q < q1; q += kFinThreads) {
        const uint4 v = __ldcs(src + q);
        const uint32_t b0 = min(v.x, 255u), b1 = min(v.y, 255u), b2 = min(v.z, 255u), b3 = min(v.w, 255u);
        dst[q] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        if (hist) {
            s_hist[min(b0, 63u) * kFinThreads + threadIdx.x]++;
            s_hist[min(b1, 63u) * kFinThreads + threadIdx.x]++;
            s_hist[min(b2, 63u) * kFinThreads + threadIdx.x]++;
            s_hist[min(b3, 63u) * kFinThreads + threadIdx.x]++;
        }
    }
    if (!hist) return;
    __syncthreads();
    // bin totals: warp w sums bins w, w+4, ...
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = warp; b < DD_HIST_BINS; b += kFinThreads / 32) {
        uint32_t s = 0;
        for (int t = lane; t < kFinThreads; t += 32) s += s_hist[b * kFinThreads + t];
        s = __reduce_add_sync(0xffffffffu, s);
        if (lane == 0 && s) atomicAdd(&hist[(size_t)table * DD_HIST_BINS + b], s);
    }
}

// ---- host side ----------------------------------------------------------------------------------
size_t sketch_workspace_bytes(int nk, int p) {
    return sizeof(SketchWsHeader) + 256 + (size_t)nk * sizeof(uint32_t) * ((size_t)1 << p);
}
static SketchWsHeader *ws_hdr(void *ws) { return static_cast<SketchWsHeader *>(ws); }
static uint32_t *ws_scratch(void *ws) { return reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ws) + sizeof(SketchWsHeader)); }
static uint32_t *ws_acc(void *ws) { return reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ws) + sizeof(SketchWsHeader) + 256); }

cudaError_t sketch_begin(void *d_ws, int nk, int p, cudaStream_t stream) {
    return cudaMemsetAsync(d_ws, 0, sketch_workspace_bytes(nk, p), stream);
}

cudaError_t sketch_update(const uint32_t *d_codes, const uint32_t *d_invalid, const dd_pack_state *d_state,
                          uint64_t sym_begin, uint64_t sym_end, size_t max_new_symbols, uint32_t kmask, int p,
                          int canon, void *d_ws, cudaStream_t stream) {
    SketchArgs a;
    a.codes = d_codes;
    a.invalid = d_invalid;
    a.state = d_state;
    a.sym_begin = sym_begin;
    a.sym_end = sym_end;
    a.kmask = kmask;
    a.p = p;
    a.acc = ws_acc(d_ws);
    a.hdr = ws_hdr(d_ws);
    const size_t nsym = d_state ? max_new_symbols : (size_t)(sym_end - sym_begin);
    if (nsym == 0) return cudaSuccess;
    // +2 words: the range may start and end in the middle of a word
    const size_t nwords = (nsym + 15) / 16 + 2;
    const unsigned grid = (unsigned)((nwords + kSketchThreads - 1) / kSketchThreads);
    if (canon) sketch_allk_kernel<true><<<grid, kSketchThreads, 0, stream>>>(a);
    else sketch_allk_kernel<false><<<grid, kSketchThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t sketch_refresh_floor(void *d_ws, uint32_t kmask, int p, cudaStream_t stream) {
    const int nk = __builtin_popcount(kmask);
    cudaError_t e = cudaMemsetAsync(ws_scratch(d_ws), 0xff, 32 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    floor_min_kernel<<<dim3(16, nk), 256, 0, stream>>>(ws_acc(d_ws), p, ws_hdr(d_ws), ws_scratch(d_ws));
    floor_publish_kernel<<<1, 32, 0, stream>>>(ws_hdr(d_ws), ws_scratch(d_ws), kmask);
    return cudaGetLastError();
}

cudaError_t sketch_end(void *d_ws, int nk, int p, uint8_t *d_regs, uint32_t *d_hist, double *d_cards,
                       cudaStream_t stream) {
    cudaError_t e;
    if (d_hist && (e = cudaMemsetAsync(d_hist, 0, (size_t)nk * DD_HIST_BINS * sizeof(uint32_t), stream)) != cudaSuccess)
        return e;
    finalize_kernel<<<dim3(kFinSlices, nk), kFinThreads, 0, stream>>>(ws_acc(d_ws), p, d_regs, d_hist);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (d_hist && d_cards) return mle_from_hist(d_hist, nk, p, d_cards, stream);
    return cudaSuccess;
}

}  // namespace dd
