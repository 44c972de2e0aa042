// kernels.cuh -- host-side launchers shared between the .cu files and the C ABI (api.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/dandd_b200.h"

namespace dd {

// Every kernel launch of the library is counted (dd_kernel_launches(): bench.py reports the number of
// launches inside its timed region as measured, not as a formula).
extern unsigned long long g_kernel_launches;
#define DD_COUNT_LAUNCH() ((void)__atomic_fetch_add(&dd::g_kernel_launches, 1ull, __ATOMIC_RELAXED))

// ---- K1 (pack.cu)
size_t pack_workspace_bytes(size_t chunk_bytes);
cudaError_t pack_reset(uint32_t *d_codes, size_t codes_bytes, uint32_t *d_invalid, size_t invalid_bytes,
                       dd_pack_state *d_state, cudaStream_t stream);
cudaError_t pack_fasta(const uint8_t *d_text, size_t n, uint32_t *d_codes, uint32_t *d_invalid, size_t cap_symbols,
                       dd_pack_state *d_state, void *d_ws, cudaStream_t stream);
cudaError_t pack_polyt_sentinel(const uint32_t *d_codes, uint32_t *d_invalid, const dd_pack_state *d_state,
                                uint64_t sym_begin, uint64_t sym_end, size_t max_symbols, cudaStream_t stream);
extern int g_polyt_sentinel;  // dd_set_option("polyt_sentinel"): the *_host entry points apply the pass

// ---- K2 (sketch.cu).  Workspace = [SketchWsHeader | u16 accumulators [nk][2^p]]
extern int g_k_per_pass;
extern int g_midk;
struct SketchWsHeader {
    uint8_t floor[32];  // floor[k-1] <= min(register of k): updates with rank <= floor are no-ops
    uint32_t use_floor;
    uint32_t pad[7];
};
size_t sketch_workspace_bytes(int nk, int p);
cudaError_t sketch_begin(void *d_ws, int nk, int p, cudaStream_t stream);
// Either d_state != nullptr (range read on the device: [prev_nsym, nsym), grid sized from
// max_new_symbols; a non-empty [sym_begin, sym_end) then selects a sub-range RELATIVE to prev_nsym)
// or an explicit host-known range.
cudaError_t sketch_update(const uint32_t *d_codes, const uint32_t *d_invalid, const dd_pack_state *d_state,
                          uint64_t sym_begin, uint64_t sym_end, size_t max_new_symbols, uint32_t kmask, int p,
                          int canon, void *d_ws, cudaStream_t stream);
cudaError_t sketch_refresh_floor(void *d_ws, uint32_t kmask, int p, cudaStream_t stream);
// sketch_update cut at the floor schedule's boundaries, with the refreshes in between (sketch.cu).
// d_state != nullptr: [sym_begin, sym_end) is ignored, the range is the last pack call's.
cudaError_t sketch_update_sched(const uint32_t *d_codes, const uint32_t *d_invalid, const dd_pack_state *d_state,
                                uint64_t sym_begin, uint64_t sym_end, size_t max_new_symbols, uint64_t seen_before,
                                uint32_t kmask, int p, int canon, void *d_ws, cudaStream_t stream);
cudaError_t sketch_end(void *d_ws, int nk, int p, uint8_t *d_regs, uint32_t *d_hist, double *d_cards,
                       cudaStream_t stream);

// ---- K3 / K4 / K6 (card.cu)
cudaError_t card_hist(const uint8_t *d_regs, int nsk, int p, uint32_t *d_hist, cudaStream_t stream);
cudaError_t mle_from_hist(const uint32_t *d_hist, int nsk, int p, double *d_cards, cudaStream_t stream);
cudaError_t union_max(const uint8_t *const *d_in, int n_in, size_t len, uint8_t *d_out, cudaStream_t stream);
cudaError_t prefix_union_hist(const uint8_t *d_regs, const int32_t *d_order, int n_ord, int n_steps, int n_genomes,
                              int nk, int p, int final_only, uint32_t *d_hist, uint8_t *d_unions,
                              cudaStream_t stream);

// planes.cu: the same contract as prefix_union_hist (without materialised unions) on bit-sliced sketches
bool planes_supported(int p);
size_t planes_bytes(int64_t n_sketches, int p);
size_t prefix_union_workspace_bytes(int n_ord, int n_steps, int n_genomes, int nk, int p);
cudaError_t to_planes(const uint8_t *d_regs, int64_t n_sketches, int p, uint32_t *d_planes, cudaStream_t stream);
cudaError_t prefix_union_hist_from_planes(const uint32_t *d_planes, const int32_t *d_order, int n_ord, int n_steps,
                                          int n_genomes, int nk, int p, int final_only, uint32_t *d_hist, void *d_scratch,
                                          cudaStream_t stream);
cudaError_t prefix_union_hist_planes(const uint8_t *d_regs, const int32_t *d_order, int n_ord, int n_steps, int n_genomes,
                                     int nk, int p, int final_only, uint32_t *d_hist, void *d_ws, cudaStream_t stream);
extern int g_prefix_planes;  // tuning knob: 1 = use the bit-plane kernel where it applies (default)
cudaError_t union_sets_hist(const uint8_t *const *d_members, int n_sets, int n_steps, int p, int final_only,
                            uint32_t *d_hist, uint8_t *d_unions, cudaStream_t stream);

// ---- K5 (exact.cu).  Workspace = [ExactWsHeader | bitmap or key table]
struct ExactWsHeader {
    unsigned long long count;     // distinct keys inserted (hash-set mode; bitmap mode counts on demand)
    unsigned long long overflow;  // table full
    unsigned long long saw_ones;  // the all-ones key (== the empty marker) was inserted
    unsigned long long cur_stream;  // k > 64: index (in the stream table) of the stream being inserted
    unsigned long long n_streams;   // k > 64: streams registered so far
};
size_t exact_workspace_bytes(int k, uint64_t capacity);
cudaError_t exact_begin(void *d_ws, int k, uint64_t capacity, cudaStream_t stream);
cudaError_t exact_insert(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end,
                         int k, int canon, void *d_ws, uint64_t capacity, uint32_t shard_rank, uint32_t shard_world,
                         cudaStream_t stream);
cudaError_t exact_count(void *d_ws, int k, uint64_t capacity, uint64_t *d_count, cudaStream_t stream);

}  // namespace dd
