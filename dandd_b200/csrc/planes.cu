// planes.cu -- K3 on bit-sliced sketches: progressive prefix unions with fused histograms.
//
// The byte kernel (card.cu) spends ~7 instructions per register per step, almost all of it in the
// thread-private histogram (extract byte, address, LDS, add, STS).  Here a sketch is first
// transposed once per call into 6 bit-planes: plane b, word g holds bit b of registers 32g..32g+31
// (ranks are <= 64-p+1 < 64).  On planes
//   * the running max of two sketches is a bit-serial compare-and-select, 2 LOP3 per plane
//     (12 per 32 registers), and
//   * "how many registers equal v" is AND-ing the six planes (or their complements) and a POPC:
//     1.5 LOP3 + POPC + IADD per value per 32 registers.
// Values are counted in four buckets of 16 (top two planes); a bucket nobody in the warp populates
// is skipped, a thinly populated one is walked register by register, a dense one runs the unrolled
// 16-value loop -- so the cost follows the actual value range of the data and every register is
// counted exactly once.  ~3.3 instructions per register per step instead of ~7.
//
// Same contract as prefix_union_hist (index mode, no materialised unions); replaces the reference's
// n(n+1)/2 `dashing union` + `dashing card` processes per ordering (lib/huffman_dandd.py:624-663).
#include <cuda_runtime.h>

#include <utility>

#include "common.cuh"
#include "kernels.cuh"

namespace dd {

constexpr int kPlThreads = 64;   // measured (bench prefix unions): 64 -> 1.51 ms, 128 -> 1.54, 256 -> 1.60, 512 -> 1.74
constexpr int kPlanes = 6;

// ---- u8 registers -> bit planes -------------------------------------------------------------------
// One thread per group of 32 registers: two 128-bit loads, then a SWAR transpose -- for each of the
// six bits, the four bytes of a word give a nibble with one mask, one multiply and one shift
// (same gather as dd::gather_bit7), eight nibbles make the plane word.  Stores are one word per
// plane per thread, contiguous across the warp.
__device__ __forceinline__ uint32_t bit_nibble(uint32_t w, int b) {   // bit b of bytes 0..3 -> bits 0..3
    return ((((w >> b) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;
}
__global__ void __launch_bounds__(256)
to_planes_kernel(const uint8_t *__restrict__ regs, uint32_t *__restrict__ planes, int p, size_t total_groups) {
    const size_t ngroups = (size_t)1 << (p - 5);
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups; g += (size_t)gridDim.x * blockDim.x) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(regs) + 2 * g);
        const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(regs) + 2 * g + 1);
        const uint32_t w[8] = {__vminu4(lo.x, 0x3f3f3f3fu), __vminu4(lo.y, 0x3f3f3f3fu), __vminu4(lo.z, 0x3f3f3f3fu),
                               __vminu4(lo.w, 0x3f3f3f3fu), __vminu4(hi.x, 0x3f3f3f3fu), __vminu4(hi.y, 0x3f3f3f3fu),
                               __vminu4(hi.z, 0x3f3f3f3fu), __vminu4(hi.w, 0x3f3f3f3fu)};
        const size_t sketch = g >> (p - 5), group = g & (ngroups - 1);
        uint32_t *dst = planes + sketch * kPlanes * ngroups + group;
#pragma unroll
        for (int b = 0; b < kPlanes; ++b) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) word |= bit_nibble(w[j], b) << (4 * j);
            dst[(size_t)b * ngroups] = word;
        }
    }
}

// running max of bit-sliced values: R = max(R, X).  "R < X" is the borrow of R - X, carried from the
// least significant plane up: where the bits differ the higher plane decides, where they agree the
// verdict so far stands -- a three-input function, one LOP3 per plane; then one select per plane.
__device__ __forceinline__ void plane_max(uint32_t (&R)[kPlanes], const uint32_t (&X)[kPlanes]) {
    uint32_t lt = 0u;
#pragma unroll
    for (int b = 0; b < kPlanes; ++b) lt = (~R[b] & X[b]) | (~(R[b] ^ X[b]) & lt);
#pragma unroll
    for (int b = 0; b < kPlanes; ++b) R[b] = (X[b] & lt) | (R[b] & ~lt);
}

// registers of `members` whose low four bits equal J
template <int J>
__device__ __forceinline__ uint32_t match_low4(const uint32_t (&P)[kPlanes], uint32_t members) {
    uint32_t t = members;
    t &= (J & 1) ? P[0] : ~P[0];
    t &= (J & 2) ? P[1] : ~P[1];
    t &= (J & 4) ? P[2] : ~P[2];
    t &= (J & 8) ? P[3] : ~P[3];
    return t;
}

// Two values per warp reduction (a lane holds at most 128 registers, so a warp sum fits 16 bits);
// lane J keeps the warp's count of value J, and the sixteen lanes issue one shared-memory add
// together at the end -- no branch, vote or atomic per value.
template <int J>
__device__ __forceinline__ uint32_t count_pair(const uint32_t (&P)[4][kPlanes], const uint32_t (&mem)[4]) {
    const uint32_t a = __popc(match_low4<J>(P[0], mem[0])) + __popc(match_low4<J>(P[1], mem[1])) +
                       __popc(match_low4<J>(P[2], mem[2])) + __popc(match_low4<J>(P[3], mem[3]));
    const uint32_t b = __popc(match_low4<J + 1>(P[0], mem[0])) + __popc(match_low4<J + 1>(P[1], mem[1])) +
                       __popc(match_low4<J + 1>(P[2], mem[2])) + __popc(match_low4<J + 1>(P[3], mem[3]));
    return __reduce_add_sync(0xffffffffu, a | (b << 16));
}

template <int... Js>
__device__ __forceinline__ void count_bucket_dense(std::integer_sequence<int, Js...>, const uint32_t (&P)[4][kPlanes],
                                                   const uint32_t (&mem)[4], uint32_t *s_cnt, int lane) {
    uint32_t mine = 0;
    (([&] {
         const uint32_t both = count_pair<2 * Js>(P, mem);
         if (lane == 2 * Js) mine = both & 0xFFFFu;
         if (lane == 2 * Js + 1) mine = both >> 16;
     }()),
     ...);
    if (lane < 16 && mine) atomicAdd(&s_cnt[lane], mine);
}

__device__ __forceinline__ void count_bucket_sparse(const uint32_t (&P)[4][kPlanes], const uint32_t (&mem)[4],
                                                    uint32_t *s_cnt_all) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t left = mem[c];
        while (left) {
            const int i = __ffs((int)left) - 1;
            left &= left - 1;
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < kPlanes; ++b) v |= ((P[c][b] >> i) & 1u) << b;
            atomicAdd(&s_cnt_all[v], 1u);
        }
    }
}

// ---- identical prefixes --------------------------------------------------------------------------
// Two orderings whose first s+1 entries are the same SET have the same union at step s (always true
// for the last step of permutations of one set, frequent at the first and last few steps of random
// ones): only the lowest-numbered ordering of each such class (among the first kDedupWindow orderings)
// counts, the others get a copy of its histogram row.  Sets are 64-bit masks, so this applies to n_genomes <= 64; beyond that every
// ordering is its own representative.
constexpr int kDedupWindow = 256;   // representatives are looked for among the first 256 orderings

__global__ void __launch_bounds__(1024)
prefix_dedup_kernel(const int32_t *__restrict__ order, int n_ord, int n_steps, int n_genomes,
                    unsigned long long *__restrict__ masks, int32_t *__restrict__ rep) {
    const int total = n_ord * n_steps;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int o = i / n_steps, st = i % n_steps;
        unsigned long long m = 0ull;
        for (int j = 0; j <= st; ++j) {
            const int g = order[(size_t)o * n_steps + j];
            if (g >= 0 && g < n_genomes && n_genomes <= 64) m |= 1ull << g;
        }
        masks[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int o = i / n_steps, st = i % n_steps;
        int r = o;
        if (n_genomes <= 64) {
            const unsigned long long m = masks[i];
            const int qmax = o < kDedupWindow ? o : kDedupWindow;   // bounded search: linear in n_ord
            for (int q = 0; q < qmax; ++q)
                if (masks[(size_t)q * n_steps + st] == m) {
                    r = q;
                    break;
                }
        }
        rep[i] = r;
    }
}

// hist rows of non-representative (ordering, step) pairs <- the representative's rows
__global__ void prefix_copy_rows_kernel(const int32_t *__restrict__ rep, int n_ord, int n_steps, int nk, int final_only,
                                        uint32_t *__restrict__ hist) {
    const int out_steps = final_only ? 1 : n_steps;
    const size_t rows = (size_t)n_ord * out_steps * nk;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * DD_HIST_BINS; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / DD_HIST_BINS, bin = i % DD_HIST_BINS;
        const int k = (int)(row % nk);
        const int os = (int)(row / nk);  // o * out_steps + step'
        const int o = os / out_steps, st = final_only ? n_steps - 1 : os % out_steps;
        const int r = rep[(size_t)o * n_steps + st];
        if (r != o) {
            const size_t src = final_only ? (size_t)r * nk + k : ((size_t)r * n_steps + st) * nk + k;
            hist[i] = hist[src * DD_HIST_BINS + bin];
        }
    }
}

// grid (n_ord, slices, nk); a thread owns 4 groups (128 registers), a CTA 128 x kPlThreads registers.
// kCopyFirst: the first member is copied instead of max-ed with the all-zero start (pairs: half of
// their max work; for long orderings the extra branch costs more than the one saved step).
template <bool kCopyFirst>
__global__ void __launch_bounds__(kPlThreads, 1024 / kPlThreads)
prefix_union_planes_kernel(const uint32_t *__restrict__ planes, const int32_t *__restrict__ order, int n_steps,
                           int n_genomes, int nk, int p, int final_only, const int32_t *__restrict__ rep,
                           uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_cnt[DD_HIST_BINS];
    const size_t ngroups = (size_t)1 << (p - 5);
    const size_t nvec = ngroups >> 2;  // uint4 per plane
    const int o = blockIdx.x, k = blockIdx.z;
    const size_t vec = (size_t)blockIdx.y * kPlThreads + threadIdx.x;
    const bool owner = vec < nvec;
    const int lane = threadIdx.x & 31;
    uint32_t R[4][kPlanes];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int b = 0; b < kPlanes; ++b) R[c][b] = 0u;

    for (int step = 0; step < n_steps; ++step) {
        const int g = order[(size_t)o * n_steps + step];
        if (g >= 0 && g < n_genomes && owner) {
            const uint4 *src = reinterpret_cast<const uint4 *>(planes + ((size_t)g * nk + k) * kPlanes * ngroups);
            uint4 x[kPlanes];
#pragma unroll
            for (int b = 0; b < kPlanes; ++b) x[b] = __ldg(src + (size_t)b * nvec + vec);
            if (kCopyFirst && step == 0) {  // max(0, X) = X
#pragma unroll
                for (int b = 0; b < kPlanes; ++b) {
                    R[0][b] = x[b].x;
                    R[1][b] = x[b].y;
                    R[2][b] = x[b].z;
                    R[3][b] = x[b].w;
                }
            } else {
                uint32_t X[kPlanes];
#pragma unroll
                for (int b = 0; b < kPlanes; ++b) X[b] = x[b].x;
                plane_max(R[0], X);
#pragma unroll
                for (int b = 0; b < kPlanes; ++b) X[b] = x[b].y;
                plane_max(R[1], X);
#pragma unroll
                for (int b = 0; b < kPlanes; ++b) X[b] = x[b].z;
                plane_max(R[2], X);
#pragma unroll
                for (int b = 0; b < kPlanes; ++b) X[b] = x[b].w;
                plane_max(R[3], X);
            }
        }
        if (final_only && step != n_steps - 1) continue;
        if (rep && rep[(size_t)o * n_steps + step] != o) continue;  // same set as an earlier ordering: row copied afterwards
        const size_t row = final_only ? (size_t)o * nk + k : ((size_t)o * n_steps + step) * nk + k;

        if (threadIdx.x < DD_HIST_BINS) s_cnt[threadIdx.x] = 0u;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // bucket q: values 16q .. 16q+15
            uint32_t mem[4];
            uint32_t any = 0u, n = 0u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                mem[c] = owner ? (((q & 1) ? R[c][4] : ~R[c][4]) & ((q & 2) ? R[c][5] : ~R[c][5])) : 0u;
                any |= mem[c];
            }
            if (!__any_sync(0xffffffffu, any != 0u)) continue;  // nobody in the warp has such values
#pragma unroll
            for (int c = 0; c < 4; ++c) n += __popc(mem[c]);
            if (__any_sync(0xffffffffu, n > 12u)) count_bucket_dense(std::make_integer_sequence<int, 8>{}, R, mem, s_cnt + 16 * q, lane);
            else count_bucket_sparse(R, mem, s_cnt);
        }
        __syncthreads();
        if (threadIdx.x < DD_HIST_BINS) {
            const uint32_t c = s_cnt[threadIdx.x];
            if (c) atomicAdd(&hist[row * DD_HIST_BINS + threadIdx.x], c);
        }
        __syncthreads();
    }
}

// ---- host side ----------------------------------------------------------------------------------
bool planes_supported(int p) { return p >= 12; }   // whole uint4s of groups per plane, >= 1 CTA of work
size_t planes_bytes(int64_t n_sketches, int p) { return (size_t)n_sketches * kPlanes * (((size_t)1 << p) / 8); }

static size_t dedup_bytes(int n_ord, int n_steps) {
    return (size_t)n_ord * n_steps * (sizeof(unsigned long long) + sizeof(int32_t));
}
// scratch of the register-input entry: [planes | identical-prefix masks | representatives]
size_t prefix_union_workspace_bytes(int n_ord, int n_steps, int n_genomes, int nk, int p) {
    const size_t pl = (planes_bytes((int64_t)n_genomes * nk, p) + 255) / 256 * 256;
    return pl + (dedup_bytes(n_ord, n_steps) + 255) / 256 * 256 + 256;
}

cudaError_t to_planes(const uint8_t *d_regs, int64_t n_sketches, int p, uint32_t *d_planes, cudaStream_t stream) {
    const size_t total_groups = ((size_t)n_sketches << p) >> 5;
    if (total_groups == 0) return cudaSuccess;
    size_t blocks = (total_groups + 255) / 256;
    if (blocks > (size_t)148 * 64) blocks = (size_t)148 * 64;
    DD_COUNT_LAUNCH(), to_planes_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_regs, d_planes, p, total_groups);
    return cudaGetLastError();
}

// d_scratch: dedup_bytes(n_ord, n_steps) bytes, or nullptr (no identical-prefix search)
cudaError_t prefix_union_hist_from_planes(const uint32_t *d_planes, const int32_t *d_order, int n_ord, int n_steps,
                                          int n_genomes, int nk, int p, int final_only, uint32_t *d_hist, void *d_scratch,
                                          cudaStream_t stream) {
    const int out_steps = final_only ? 1 : n_steps;
    const size_t rows = (size_t)n_ord * out_steps * nk;
    cudaError_t e = cudaMemsetAsync(d_hist, 0, rows * DD_HIST_BINS * sizeof(uint32_t), stream);
    if (e != cudaSuccess || rows == 0) return e;
    const size_t m = (size_t)1 << p;
    const size_t pairs = (size_t)n_ord * n_steps;
    unsigned long long *masks = reinterpret_cast<unsigned long long *>(d_scratch);
    int32_t *rep = d_scratch ? reinterpret_cast<int32_t *>(masks + pairs) : nullptr;
    // one CTA compares every (ordering, step) with the first kDedupWindow orderings: the big classes
    // (last steps, first steps) are always found there; pointless for a batch of distinct pairs
    const bool dedup = d_scratch && n_genomes <= 64 && n_ord > 1 && !(final_only && n_steps == 2) &&
                       (double)n_ord * (n_ord < kDedupWindow ? n_ord : kDedupWindow) * n_steps <= 5e7;
    if (dedup) DD_COUNT_LAUNCH(), prefix_dedup_kernel<<<1, 1024, 0, stream>>>(d_order, n_ord, n_steps, n_genomes, masks, rep);
    else rep = nullptr;
    const size_t nvec = (m >> 5) >> 2;
    const unsigned slices = (unsigned)((nvec + kPlThreads - 1) / kPlThreads);
    // the ordering / pair index is the fastest grid dimension (co-resident CTAs share slices in L2);
    // gridDim.x has room for 2^31-1 of them
    const dim3 grid((unsigned)n_ord, slices, (unsigned)nk);
    if (final_only && n_steps == 2)
        DD_COUNT_LAUNCH(), prefix_union_planes_kernel<true><<<grid, kPlThreads, 0, stream>>>(d_planes, d_order, n_steps, n_genomes, nk, p, final_only,
                                                                         rep, d_hist);
    else
        DD_COUNT_LAUNCH(), prefix_union_planes_kernel<false><<<grid, kPlThreads, 0, stream>>>(d_planes, d_order, n_steps, n_genomes, nk, p, final_only,
                                                                          rep, d_hist);
    if (dedup) {
        const size_t cells = rows * DD_HIST_BINS;
        const unsigned cb = (unsigned)((cells + 255) / 256 < 1184 ? (cells + 255) / 256 : 1184);
        DD_COUNT_LAUNCH(), prefix_copy_rows_kernel<<<cb, 256, 0, stream>>>(rep, n_ord, n_steps, nk, final_only, d_hist);
    }
    return cudaGetLastError();
}

// register input: transpose into the caller's workspace, then the planes path
cudaError_t prefix_union_hist_planes(const uint8_t *d_regs, const int32_t *d_order, int n_ord, int n_steps, int n_genomes,
                                     int nk, int p, int final_only, uint32_t *d_hist, void *d_ws, cudaStream_t stream) {
    uint32_t *planes = reinterpret_cast<uint32_t *>(d_ws);
    const size_t pl = (planes_bytes((int64_t)n_genomes * nk, p) + 255) / 256 * 256;
    cudaError_t e = to_planes(d_regs, (int64_t)n_genomes * nk, p, planes, stream);
    if (e != cudaSuccess) return e;
    return prefix_union_hist_from_planes(planes, d_order, n_ord, n_steps, n_genomes, nk, p, final_only, d_hist,
                                         static_cast<uint8_t *>(d_ws) + pl, stream);
}

}  // namespace dd
