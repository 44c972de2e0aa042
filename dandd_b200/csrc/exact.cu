// exact.cu -- K5: --exact mode, the number of distinct (canonical) k-mers, KMC semantics.
//
// Replaces `kmc -ci1 -cs2 -kK [-b] -fm` + `kmc_tools info` (reference lib/sketch_classes.py:389-399,
// 434-449) and `kmc_tools complex` unions (:451-465).  KMC builds an on-disk sorted database; all
// DandD ever reads back is its size, so the GPU keeps a *set* in HBM instead:
//   k <= 16 : a presence bitmap of 4^k bits (<= 512 MiB), filled with test-then-atomicOr, counted
//             with popc -- exact by construction;
//   k  > 16 : an open-addressing table of 64-bit keys (linear probing, atomicCAS); the key is the
//             k-mer value itself, so there are no false merges -- exact as long as the table is
//             not full, which is reported, never ignored;
//   k  > 32 : (to 64) the same with 128-bit keys and `atom.cas.b128`;
//   k  > 64 : (to 256, KMC's own limit) no atomic is wide enough for the key, so a 16-byte entry holds a
//             64-bit fingerprint of the canonical k-mer plus a REFERENCE to one occurrence (stream,
//             position); entries are still claimed with one `atom.cas.b128`, and a fingerprint match
//             is settled by comparing the k-mer at the referenced position word for word -- exact,
//             fingerprints only save comparisons.
// Inserting several genomes into one set gives the union count; reading the count after each
// insertion gives the progressive exact unions in one sweep (the reference re-merges databases
// n(n+1)/2 times).
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dd {

constexpr int kExactThreads = 256;
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

static size_t exact_table_bytes(int k, uint64_t capacity) {
    if (k <= DD_EXACT_BITMAP_MAXK) {
        const size_t bits = (size_t)1 << (2 * k);
        return bits < 128 ? 16 : bits / 8;
    }
    return (size_t)capacity * (k > 32 ? 2 : 1) * sizeof(unsigned long long);
}
constexpr unsigned kMaxStreams = 512;   // k > 64: streams whose k-mers one set may reference
static size_t exact_stream_table_bytes(int k) { return k > 64 ? kMaxStreams * sizeof(const uint32_t *) : 0; }
size_t exact_workspace_bytes(int k, uint64_t capacity) { return 256 + exact_table_bytes(k, capacity) + exact_stream_table_bytes(k); }
static ExactWsHeader *ex_hdr(void *ws) { return static_cast<ExactWsHeader *>(ws); }
static void *ex_tab(void *ws) { return static_cast<uint8_t *>(ws) + 256; }

// fmix64 (MurmurHash3 finaliser) -- only spreads keys over slots, never decides equality
__device__ __forceinline__ uint64_t slot_hash(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// Key-range sharding over ranks: a k-mer belongs to rank (hash >> 40) % world -- the
// high hash bits, so the low bits that pick the slot stay uniform inside every shard.  Each rank
// scans everything but stores only its own keys; the per-rank distinct counts add up exactly.
__device__ __forceinline__ bool mine(uint64_t h, uint32_t shard_rank, uint32_t shard_world) {
    return shard_world <= 1u || (uint32_t)(h >> 40) % shard_world == shard_rank;
}

template <bool kBitmap>
__global__ void __launch_bounds__(kExactThreads)
exact_insert_kernel(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ invalid, uint64_t sym_begin,
                    uint64_t sym_end, int k, int canon, ExactWsHeader *hdr, void *table, uint64_t capacity,
                    uint32_t shard_rank, uint32_t shard_world) {
    const uint64_t w = (sym_begin >> 4) + (uint64_t)blockIdx.x * kExactThreads + threadIdx.x;
    const uint64_t s0 = w << 4;
    unsigned long long fresh = 0;
    if (s0 < sym_end) {
        const uint32_t w0 = __ldg(codes + w);
        const uint32_t w1 = w >= 1 ? __ldg(codes + w - 1) : 0u;
        const uint32_t w2 = w >= 2 ? __ldg(codes + w - 2) : 0u;
        const uint64_t iw = w >> 1;
        const uint32_t i0 = __ldg(invalid + iw);
        const uint32_t i1 = iw >= 1 ? __ldg(invalid + iw - 1) : 0xffffffffu;
        const uint32_t r0 = revcomp_word(w0), r1 = revcomp_word(w1), r2 = revcomp_word(w2);
        const int j_lo = sym_begin > s0 ? (int)(sym_begin - s0) : 0;
        const int j_hi = sym_end - s0 < 16 ? (int)(sym_end - s0) : 16;
        const uint32_t sm_base = (uint32_t)(s0 & 31);
        for (int j = j_lo; j < j_hi; ++j) {
            if (valid_run(invalid_window(i0, i1, sm_base + (uint32_t)j)) < k) continue;
            const Window win = window_at(w0, w1, w2, r0, r1, r2, j);
            const uint64_t v = kmer_value_rt(win, k, canon != 0);
            if (!mine(slot_hash(v), shard_rank, shard_world)) continue;  // another rank's key range
            if (kBitmap) {
                uint32_t *bm = static_cast<uint32_t *>(table);
                const uint32_t bit = 1u << (v & 31);
                uint32_t *word = bm + (v >> 5);
                if (!(__ldcg(word) & bit)) atomicOr(word, bit);
            } else {
                if (v == kEmptyKey) {  // cannot be stored: it is the empty marker
                    if (atomicExch(&hdr->saw_ones, 1ull) == 0ull) ++fresh;
                    continue;
                }
                unsigned long long *tab = static_cast<unsigned long long *>(table);
                const uint64_t mask = capacity - 1;
                uint64_t slot = slot_hash(v) & mask;
                uint64_t probes = 0;
                for (;;) {
                    unsigned long long cur = __ldcg(tab + slot);
                    if (cur == kEmptyKey) cur = atomicCAS(tab + slot, kEmptyKey, (unsigned long long)v);
                    if (cur == kEmptyKey) { ++fresh; break; }
                    if (cur == v) break;
                    slot = (slot + 1) & mask;
                    if (++probes > mask) { hdr->overflow = 1; break; }
                }
            }
        }
    }
    if (!kBitmap) {
        // one global add per warp
        fresh = __reduce_add_sync(0xffffffffu, (unsigned)fresh);
        if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(&hdr->count, fresh);
    }
}

// ---- k = 33..64: 128-bit keys ---------------------------------------------------------------------
__device__ __forceinline__ ulonglong2 cas128(ulonglong2 *addr, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
        "atom.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old.x), "=l"(old.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
        : "memory");
    return old;
}

__global__ void __launch_bounds__(kExactThreads)
exact_insert_wide_kernel(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ invalid, uint64_t sym_begin,
                         uint64_t sym_end, int k, int canon, ExactWsHeader *hdr, ulonglong2 *tab, uint64_t capacity,
                         uint32_t shard_rank, uint32_t shard_world) {
    const uint64_t w = (sym_begin >> 4) + (uint64_t)blockIdx.x * kExactThreads + threadIdx.x;
    const uint64_t s0 = w << 4;
    unsigned long long fresh = 0;
    if (s0 < sym_end) {
        uint32_t c[5], I[5];
        const uint64_t iw = w >> 1;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            c[t] = w >= (uint64_t)t ? __ldg(codes + w - t) : 0u;
            I[t] = iw >= (uint64_t)t ? __ldg(invalid + iw - t) : 0xffffffffu;  // before the stream: breaks
        }
        const int j_lo = sym_begin > s0 ? (int)(sym_begin - s0) : 0;
        const int j_hi = sym_end - s0 < 16 ? (int)(sym_end - s0) : 16;
        const uint32_t sm_base = (uint32_t)(s0 & 31);
        const ulonglong2 empty = make_ulonglong2(kEmptyKey, kEmptyKey);
        const uint64_t mask = capacity - 1;
        for (int j = j_lo; j < j_hi; ++j) {
            if (valid_run_long(I, sm_base + (uint32_t)j) < k) continue;
            const U128 v = kmer128_at(c, j, k, canon != 0);
            if (v.lo == kEmptyKey && v.hi == kEmptyKey) {  // cannot be stored: it is the empty marker (rank 0's)
                if (shard_rank == 0u && atomicExch(&hdr->saw_ones, 1ull) == 0ull) ++fresh;
                continue;
            }
            const uint64_t hv = slot_hash(v.lo ^ slot_hash(v.hi));
            if (!mine(hv, shard_rank, shard_world)) continue;
            const ulonglong2 key = make_ulonglong2(v.lo, v.hi);
            uint64_t slot = hv & mask;
            uint64_t probes = 0;
            for (;;) {
                // a matching read settles it; anything else is decided by the CAS result alone
                const ulonglong2 seen = __ldcg(tab + slot);
                if (seen.x == key.x && seen.y == key.y) break;
                const ulonglong2 old = cas128(tab + slot, empty, key);
                if (old.x == kEmptyKey && old.y == kEmptyKey) { ++fresh; break; }
                if (old.x == key.x && old.y == key.y) break;
                slot = (slot + 1) & mask;
                if (++probes > mask) { hdr->overflow = 1; break; }
            }
        }
    }
    fresh = __reduce_add_sync(0xffffffffu, (unsigned)fresh);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(&hdr->count, fresh);
}

// ---- k = 65..256: (fingerprint, reference) entries ------------------------------------------------
constexpr int kRefPosBits = 48;

// one thread: find or append `codes` in the set's stream table; publish its index for the insert kernel
__global__ void exact_register_stream_kernel(ExactWsHeader *hdr, const uint32_t **streams, const uint32_t *codes) {
    unsigned long long n = hdr->n_streams, i = 0;
    while (i < n && streams[i] != codes) ++i;
    if (i == n) {
        if (n >= kMaxStreams) {
            hdr->overflow = 1;
            i = 0;
        } else {
            streams[n] = codes;
            hdr->n_streams = n + 1;
        }
    }
    hdr->cur_stream = i;
}

__global__ void __launch_bounds__(kExactThreads)
exact_insert_long_kernel(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ invalid, uint64_t sym_begin,
                         uint64_t sym_end, int k, int canon, ExactWsHeader *hdr, ulonglong2 *tab, uint64_t capacity,
                         const uint32_t *const *streams, uint32_t shard_rank, uint32_t shard_world) {
    const uint64_t w = (sym_begin >> 4) + (uint64_t)blockIdx.x * kExactThreads + threadIdx.x;
    const uint64_t s0 = w << 4;
    unsigned long long fresh = 0;
    if (s0 < sym_end) {
        const int j_lo = sym_begin > s0 ? (int)(sym_begin - s0) : 0;
        const int j_hi = sym_end - s0 < 16 ? (int)(sym_end - s0) : 16;
        const ulonglong2 empty = make_ulonglong2(kEmptyKey, kEmptyKey);
        const uint64_t mask = capacity - 1;
        const unsigned long long me = hdr->cur_stream << kRefPosBits;
        int run = 0;
        for (int j = j_lo; j < j_hi; ++j) {
            const uint64_t s = s0 + j;
            // valid symbols ending at s (saturating at k): the first position of the word walks the break
            // bits backwards, the others extend its count
            if (j == j_lo) {
                run = valid_run_upto(invalid, s, k);
            } else {
                const bool bad = (invalid[s >> 5] >> (31u - (uint32_t)(s & 31))) & 1u;
                run = bad ? 0 : (run < k ? run + 1 : k);
            }
            if (run < k) continue;
            uint64_t key[kLongWords];
            const int W = kmer_long_at(codes, s, k, canon != 0, key);
            uint64_t h = 0;
            for (int t = 0; t < W; ++t) h = slot_hash(h ^ key[t]);
            if (!mine(h, shard_rank, shard_world)) continue;
            const ulonglong2 entry = make_ulonglong2(h, me | s);
            uint64_t slot = h & mask;
            uint64_t probes = 0;
            for (;;) {
                ulonglong2 seen = __ldcg(tab + slot);
                if (seen.x == kEmptyKey && seen.y == kEmptyKey) {
                    seen = cas128(tab + slot, empty, entry);
                    if (seen.x == kEmptyKey && seen.y == kEmptyKey) { ++fresh; break; }
                }
                if (seen.x == h) {   // same fingerprint: the k-mer at the referenced occurrence decides
                    uint64_t other[kLongWords];
                    kmer_long_at(streams[seen.y >> kRefPosBits], seen.y & ((1ull << kRefPosBits) - 1ull), k, canon != 0, other);
                    bool same = true;
                    for (int t = 0; t < W; ++t) same = same && other[t] == key[t];
                    if (same) break;
                }
                slot = (slot + 1) & mask;
                if (++probes > mask) { hdr->overflow = 1; break; }
            }
        }
    }
    fresh = __reduce_add_sync(0xffffffffu, (unsigned)fresh);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(&hdr->count, fresh);
}

__global__ void __launch_bounds__(256) bitmap_count_kernel(const uint32_t *__restrict__ bm, size_t nwords,
                                                           unsigned long long *out) {
    unsigned long long c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
        c += __popc(__ldg(bm + i));
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
// Hash-set mode: publish the running count; a full table can never yield a silent wrong answer --
// the count is replaced by the all-ones marker, which the host turns into DD_ERR_WORKSPACE.
__global__ void exact_publish_kernel(const ExactWsHeader *hdr, unsigned long long *out) {
    *out = hdr->overflow ? kEmptyKey : hdr->count;
}

cudaError_t exact_begin(void *d_ws, int k, uint64_t capacity, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_ws, 0, 256, stream);
    if (e != cudaSuccess) return e;
    const int fill = k <= DD_EXACT_BITMAP_MAXK ? 0x00 : 0xff;
    if ((e = cudaMemsetAsync(ex_tab(d_ws), fill, exact_table_bytes(k, capacity), stream)) != cudaSuccess) return e;
    if (k > 64)
        return cudaMemsetAsync(static_cast<uint8_t *>(ex_tab(d_ws)) + exact_table_bytes(k, capacity), 0, exact_stream_table_bytes(k), stream);
    return cudaSuccess;
}

cudaError_t exact_insert(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end,
                         int k, int canon, void *d_ws, uint64_t capacity, uint32_t shard_rank, uint32_t shard_world,
                         cudaStream_t stream) {
    if (sym_end <= sym_begin) return cudaSuccess;
    const size_t nwords = (size_t)((sym_end - sym_begin + 15) / 16 + 2);  // +2: the range need not start on a word boundary
    const unsigned grid = (unsigned)((nwords + kExactThreads - 1) / kExactThreads);
    if (k > 64) {
        if (sym_end >> kRefPosBits) return cudaErrorInvalidValue;
        const uint32_t **streams = reinterpret_cast<const uint32_t **>(static_cast<uint8_t *>(ex_tab(d_ws)) + exact_table_bytes(k, capacity));
        DD_COUNT_LAUNCH(), exact_register_stream_kernel<<<1, 1, 0, stream>>>(ex_hdr(d_ws), streams, d_codes);
        DD_COUNT_LAUNCH(), exact_insert_long_kernel<<<grid, kExactThreads, 0, stream>>>(d_codes, d_invalid, sym_begin, sym_end, k, canon,
                                                                    ex_hdr(d_ws), static_cast<ulonglong2 *>(ex_tab(d_ws)),
                                                                    capacity, streams, shard_rank, shard_world);
    } else if (k > 32)
        DD_COUNT_LAUNCH(), exact_insert_wide_kernel<<<grid, kExactThreads, 0, stream>>>(d_codes, d_invalid, sym_begin, sym_end, k, canon,
                                                                    ex_hdr(d_ws), static_cast<ulonglong2 *>(ex_tab(d_ws)),
                                                                    capacity, shard_rank, shard_world);
    else if (k <= DD_EXACT_BITMAP_MAXK)
        DD_COUNT_LAUNCH(), exact_insert_kernel<true><<<grid, kExactThreads, 0, stream>>>(d_codes, d_invalid, sym_begin, sym_end, k, canon,
                                                                     ex_hdr(d_ws), ex_tab(d_ws), capacity, shard_rank, shard_world);
    else
        DD_COUNT_LAUNCH(), exact_insert_kernel<false><<<grid, kExactThreads, 0, stream>>>(d_codes, d_invalid, sym_begin, sym_end, k,
                                                                      canon, ex_hdr(d_ws), ex_tab(d_ws), capacity, shard_rank,
                                                                      shard_world);
    return cudaGetLastError();
}

cudaError_t exact_count(void *d_ws, int k, uint64_t capacity, uint64_t *d_count, cudaStream_t stream) {
    (void)capacity;
    if (k <= DD_EXACT_BITMAP_MAXK) {
        cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(uint64_t), stream);
        if (e != cudaSuccess) return e;
        const size_t nwords = exact_table_bytes(k, 0) / 4;
        size_t blocks = (nwords + 256 * 8 - 1) / (256 * 8);
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 8) blocks = 148 * 8;
        DD_COUNT_LAUNCH(), bitmap_count_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint32_t *>(ex_tab(d_ws)), nwords,
                                                                 reinterpret_cast<unsigned long long *>(d_count));
    } else {
        DD_COUNT_LAUNCH(), exact_publish_kernel<<<1, 1, 0, stream>>>(ex_hdr(d_ws), reinterpret_cast<unsigned long long *>(d_count));
    }
    return cudaGetLastError();
}

}  // namespace dd
