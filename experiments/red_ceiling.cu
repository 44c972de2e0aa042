// red_ceiling.cu -- what is the chip-wide rate of scattered 4-byte reductions into an L2-resident
// table?  K2 at config-2 size does exactly one `red.global.max.noftz.v2.f16` per (base, k) into a
// 46 MB table and is bound by this rate, not by HBM or instruction issue; this probe measures the
// ceiling the K2 number is compared with (bench.py: roofline.red_rate).
//   - the same instruction K2 issues, addresses from an LCG (uniform over the table, no locality)
//   - table sizes from 2 MiB to 184 MiB (K2: 2^20 u16 registers x 23 k = 46 MiB; 126 MB L2)
//   - no hashing around it: 4 integer instructions per reduction, so issue is never the limit
//   - full chip, 8 x 256-thread CTAs per SM; also 4 CTAs per SM and a 64-SM grid for shape dependence
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o red_ceiling red_ceiling.cu
// output: one JSON document on stdout (committed as profiles/r02_red_ceiling.json)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) red_kernel(uint32_t *tab, uint32_t words_mask, int iters, int op) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t h = x ^ (x >> 15);
        uint32_t *p = tab + ((h >> 4) & words_mask);
        const uint32_t v = (1u + (h & 7u)) << (16u * ((h >> 3) & 1u));
        if (op == 0)
            asm volatile("{ .reg .b16 l, hh; mov.b32 {l, hh}, %1; red.global.max.noftz.v2.f16 [%0], {l, hh}; }" ::"l"(p), "r"(v) : "memory");
        else if (op == 1)
            asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
        else
            asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
}

static double run(uint32_t *tab, size_t bytes, int grid, int iters, int op) {
    const uint32_t mask = (uint32_t)(bytes / 4 - 1);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    red_kernel<<<grid, 256>>>(tab, mask, iters, op);   // warm-up: table becomes L2-resident
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        red_kernel<<<grid, 256>>>(tab, mask, iters, op);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double rate = (double)grid * 256 * iters / (ms * 1e-3) / 1e9;
        if (rate > best) best = rate;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t *tab;
    const size_t max_bytes = (size_t)256 << 20;
    cudaMalloc(&tab, max_bytes);
    cudaMemset(tab, 0, max_bytes);
    printf("{\n \"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"unit\": \"G reductions/s (chip)\",\n", prop.name, sms, prop.l2CacheSize);
    printf(" \"what\": \"scattered red.global into a table of the given size, uniform random addresses, best of 5 launches of 2048 reductions per thread\",\n");
    const char *ops[3] = {"max.noftz.v2.f16 (K2's instruction)", "max.u32", "or.b32"};
    printf(" \"by_table_size\": [\n");
    const size_t sizes[] = {2, 8, 32, 46, 64, 92, 128, 184};
    for (int s = 0; s < 8; ++s) {
        size_t bytes = (size_t)1 << 20;
        while (bytes < (sizes[s] << 20)) bytes <<= 1;   // power-of-two mask; 46 -> 64 MiB etc. are reported by their real extent
        const double r = run(tab, bytes, sms * 8, 2048, 0);
        printf("  {\"table_mib\": %zu, \"f16x2_max\": %.1f}%s\n", bytes >> 20, r, s == 7 ? "" : ",");
    }
    printf(" ],\n \"by_op_64mib\": {");
    for (int op = 0; op < 3; ++op) printf("\"%s\": %.1f%s", ops[op], run(tab, (size_t)64 << 20, sms * 8, 2048, op), op == 2 ? "" : ", ");
    printf("},\n \"by_grid_64mib\": {");
    const int grids[4] = {sms * 8, sms * 4, sms * 2, 64 * 8};
    const char *gn[4] = {"all SMs x 8 CTAs", "all SMs x 4 CTAs", "all SMs x 2 CTAs", "64 SMs x 8 CTAs"};
    double ceiling = 0;
    for (int g = 0; g < 4; ++g) {
        const double r = run(tab, (size_t)64 << 20, grids[g], 2048, 0);
        if (r > ceiling) ceiling = r;
        printf("\"%s\": %.1f%s", gn[g], r, g == 3 ? "" : ", ");
    }
    printf("},\n \"ceiling_f16x2_64mib\": %.1f,\n \"clock_mhz_max\": %d\n}\n", ceiling, prop.clockRate / 1000);
    return 0;
}
