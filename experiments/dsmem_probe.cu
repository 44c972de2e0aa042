// dsmem_probe.cu -- how fast can a cluster of CTAs OR bits into a table spread over their shared
// memories?  (Question behind a cluster-resident variant of K2: one byte per HLL register, ranks
// one-hot, 2^20 registers over 8 CTAs x 128 KiB.)
//   mode 0: red.shared::cluster.or.b32 to a random CTA of the cluster (DSMEM)
//   mode 1: the same op, always to the own CTA (plain shared-memory atomic)
//   mode 2: scattered red.global.max.noftz.f16x2 over a 46 MB table (what K2 does today)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_probe dsmem_probe.cu
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int kTableBytes = 128 * 1024;

template <int CL>
__global__ void probe(uint32_t *sink, uint32_t *gtab, int iters, int mode) {
    extern __shared__ __align__(16) uint32_t tab[];
    cg::cluster_group cluster = cg::this_cluster();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = 0;
    cluster.sync();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t my = cluster.block_rank();
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t h = x ^ (x >> 15);
        const uint32_t idx = h >> 12;  // 20 bits
        const uint32_t bit = 1u << (8 * (idx & 3) + (h & 7));
        if (mode == 7 || mode == 8) {   // read filter: load the word, RED only ~45 % of the time
            uint32_t *p = gtab + ((h >> 9) % (23u << 19));
            uint32_t cur;
            asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(cur) : "l"(p) : "memory");
            const uint32_t v = 1u + (h & 7);
            const bool go = mode == 7 ? (((cur >> 20) + (h >> 3)) % 100u) < 45u : false;
            if (go) asm volatile("{ .reg .b16 l, hh; mov.b32 {l, hh}, %1; red.global.max.noftz.v2.f16 [%0], {l, hh}; }" ::"l"(p), "r"(v) : "memory");
            else if (cur == 0xdeadbeefu) sink[1] = cur;
        } else if (mode == 3) {   // one-hot byte per register: 4 registers per word, 23 MB table
            uint32_t *p = gtab + ((h >> 9) % (23u << 18));
            asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(bit) : "memory");
        } else if (mode == 4) {   // u32 max over 23 x 2^20 words (92 MB)
            uint32_t *p = gtab + ((h >> 7) % (23u << 20));
            asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(h & 63u) : "memory");
        } else if (mode == 5) {   // u32 max over the 46 MB footprint
            uint32_t *p = gtab + ((h >> 9) % (23u << 19));
            asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(h & 63u) : "memory");
        } else if (mode == 6) {   // or.b32 over the 46 MB footprint
            uint32_t *p = gtab + ((h >> 9) % (23u << 19));
            asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(bit) : "memory");
        } else if (mode == 2) {
            const uint32_t v = 1u + (h & 7);  // subnormal f16 pattern
            uint32_t *p = gtab + ((h >> 9) % (23u << 19));
            asm volatile("{ .reg .b16 l, h; mov.b32 {l, h}, %1; red.global.max.noftz.v2.f16 [%0], {l, h}; }" ::"l"(p), "r"(v) : "memory");
        } else {
            const uint32_t off = (idx & 0x1FFFFu) & ~3u;
            const uint32_t rank = mode == 0 ? (idx >> 17) % CL : my;
            uint32_t raddr;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(base + off), "r"(rank));
            asm volatile("red.relaxed.cluster.shared::cluster.or.b32 [%0], %1;" ::"r"(raddr), "r"(bit) : "memory");
        }
    }
    cluster.sync();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) acc ^= tab[i];
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

template <int CL>
static void run(int threads, int iters, uint32_t *sink, uint32_t *gtab) {
    cudaFuncSetAttribute(probe<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTableBytes);
    if (CL > 8) cudaFuncSetAttribute(probe<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = kTableBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int maxc = 0;
    cfg.gridDim = dim3(CL);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, probe<CL>, &cfg);
    printf("cluster %d, %d threads: max active clusters %d (%s)\n", CL, threads, maxc, cudaGetErrorString(e));
    if (maxc <= 0) return;
    cfg.gridDim = dim3(CL * maxc);
    for (int mode = (CL == 8 ? 0 : 2); mode < 7; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            e = cudaLaunchKernelEx(&cfg, probe<CL>, sink, gtab, iters, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double ops = (double)CL * maxc * threads * iters;
        printf("  mode %d: %s  %.3f ms  %.1f G ops/s  %.2f ops/clk/SM (at 1.965 GHz, %d SMs)\n", mode,
               cudaGetErrorString(cudaGetLastError()), best, ops / best / 1e6, ops / (best * 1e-3) / 1.965e9 / (CL * maxc),
               CL * maxc);
    }
}

static int g_mode = 2;
static void sweep(uint32_t *sink, uint32_t *gtab) {
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTableBytes);
    for (int threads : {256, 1024})
        for (int grid : {148, 296, 592}) {
            const int smem = 16 * 1024;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            float best = 1e9f;
            const int iters = 4096;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                probe<1><<<grid, threads, smem>>>(sink, gtab, iters, g_mode);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("mode %d ", g_mode); printf("sweep threads %4d grid %3d: %.3f ms  %.1f G RED/s (%s)\n", threads, grid, best,
                   (double)grid * threads * iters / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
}

int main() {
    uint32_t *sink, *gtab;
    cudaMalloc(&sink, 4);
    cudaMalloc(&gtab, (size_t)(23u << 20) * 4);
    cudaMemset(gtab, 0, (size_t)(23u << 20) * 4);
    cudaMalloc(&sink, 64);
    for (int m : {2, 7, 8}) { g_mode = m; sweep(sink, gtab); }
    return 0;
}
