// Probe: cost of a hypothetical K2 "stage A" that bins (register,rank) items by register slice
// instead of issuing global REDs.  Standalone (not part of the library); inputs are random packed
// words, correctness is not checked here -- only the time against the RED formulation matters.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I dandd_b200/csrc experiments/binA_probe.cu -o /tmp/binA
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <utility>
#include "../dandd_b200/csrc/common.cuh"
using namespace dd;

constexpr int NB = 8, CAP = 4096, T = 256;

__device__ __forceinline__ uint64_t mul64x32(uint64_t x, uint32_t c) { return x * (uint64_t)c; }

template <int K>
__device__ __forceinline__ void one_k(const Window &win, int run, uint32_t kmask, int p, uint32_t *items, uint32_t *s_cur,
                                      uint32_t tile, uint32_t ntiles, int &slot, uint32_t *acc, int mode, uint32_t lt) {
    if (!((kmask >> (K - 1)) & 1u)) return;
    const uint64_t v = kmer_value<K>(win, true);
    uint64_t h = wang64(v);
    const uint32_t hi = (uint32_t)(h >> 32), lo = (uint32_t)h;
    const uint32_t rem_hi = hi & (0xffffffffu >> p);
    const uint32_t rank = (rem_hi ? (uint32_t)__clz((int)rem_hi) : 32u + (uint32_t)__clz((int)lo)) + 1u - (uint32_t)p;
    const uint32_t idx = hi >> (32 - p);
    const bool live = run >= K;
    if (mode == 0) {  // RED formulation (as in sketch.cu)
        if (live) {
            const uint32_t val = rank << ((idx & 1u) * 16u);
            uint32_t *word = acc + ((size_t)slot << (p - 1)) + (idx >> 1);
            asm volatile("{ .reg .b16 l, h; mov.b32 {l, h}, %1; red.global.max.noftz.v2.f16 [%0], {l, h}; }" ::"l"(word), "r"(val) : "memory");
        }
    } else {          // binning formulation
        const uint32_t bin = idx >> (p - 3);
        const uint32_t key = live ? bin : (0x100u | (threadIdx.x & 31u));
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs((int)peers) - 1;
        uint32_t base = 0;
        if ((int)(threadIdx.x & 31) == leader && live) base = atomicAdd(&s_cur[slot * NB + bin], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live) {
            const uint32_t pos = base + __popc(peers & lt);
            items[(((size_t)slot * NB + bin) * ntiles + tile) * CAP + pos] = ((idx & ((1u << (p - 3)) - 1u)) << 8) | rank;
        }
    }
    ++slot;
}
template <int... Ks>
__device__ __forceinline__ void all_k(std::integer_sequence<int, Ks...>, const Window &win, int run, uint32_t kmask, int p,
                                      uint32_t *items, uint32_t *s_cur, uint32_t tile, uint32_t ntiles, uint32_t *acc, int mode, uint32_t lt) {
    int slot = 0;
    (one_k<Ks + 1>(win, run, kmask, p, items, s_cur, tile, ntiles, slot, acc, mode, lt), ...);
}

__global__ void __launch_bounds__(T) probe(const uint32_t *codes, uint32_t nwords, uint32_t kmask, int p, uint32_t *items,
                                           uint32_t *counts, uint32_t *acc, int mode) {
    __shared__ uint32_t s_cur[32 * NB];
    for (int i = threadIdx.x; i < 32 * NB; i += T) s_cur[i] = 0;
    __syncthreads();
    const uint32_t tile = blockIdx.x, ntiles = gridDim.x;
    const uint32_t w = tile * T + threadIdx.x;
    const uint32_t lt = (1u << (threadIdx.x & 31)) - 1u;
    const uint32_t w0 = w < nwords ? codes[w] : 0, w1 = w >= 1 ? codes[w - 1] : 0, w2 = w >= 2 ? codes[w - 2] : 0;
    const uint32_t r0 = revcomp_word(w0), r1 = revcomp_word(w1), r2 = revcomp_word(w2);
#pragma unroll 1
    for (int j = 0; j < 16; ++j) {
        const Window win = window_at(w0, w1, w2, r0, r1, r2, j);
        all_k(std::make_integer_sequence<int, 32>{}, win, w < nwords ? 32 : 0, kmask, p, items, s_cur, tile, ntiles, acc, mode, lt);
    }
    __syncthreads();
    if (mode == 1)
        for (int i = threadIdx.x; i < 32 * NB; i += T) counts[(size_t)i * ntiles + tile] = s_cur[i];
}

int main() {
    const int p = 20;
    const uint32_t nsym = 5000000, nwords = nsym / 16, kmask = 0xFFFFFE00u;  // k = 10..32
    const int nk = 23;
    const uint32_t ntiles = (nwords + T - 1) / T;
    uint32_t *h = (uint32_t *)malloc(nwords * 4);
    srand(1);
    for (uint32_t i = 0; i < nwords; ++i) h[i] = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    uint32_t *codes, *items, *counts, *acc;
    cudaMalloc(&codes, nwords * 4);
    cudaMemcpy(codes, h, nwords * 4, cudaMemcpyHostToDevice);
    const size_t item_bytes = (size_t)nk * NB * ntiles * CAP * 4;
    cudaMalloc(&items, item_bytes);
    cudaMalloc(&counts, (size_t)32 * NB * ntiles * 4);
    cudaMalloc(&acc, (size_t)nk << (p + 1));
    printf("ntiles %u, item region %.2f GB\n", ntiles, item_bytes / 1e9);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            cudaMemset(acc, 0, (size_t)nk << (p + 1));
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            probe<<<ntiles, T>>>(codes, nwords, kmask, p, items, counts, acc, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
        }
        printf("mode %d (%s): %.3f ms  (%s)\n", mode, mode ? "bin" : "RED", best, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
