// pack.cu -- K1: streaming FASTA text -> 2-bit packed symbol stream (+ break bit mask).
//
// Replaces the FASTA parsing half of `dashing sketch` / `kmc -fm` (reference
// lib/sketch_classes.py:351-366, 434-449).  HBM-streaming kernel: every text byte is read with
// 128-bit loads (twice: a counting pass and a writing pass; the chunk is L2-resident for the second
// when the host streams in <= 64 MiB chunks), output is 0.375 B/symbol, staged through shared
// memory so that the global stores are whole, coalesced words.
//
// The only sequential dependence in FASTA text is "am I inside a header line?".  Each 16-byte
// chunk is summarised as a transition function over that one bit (dd::chunk_xfer) and the
// functions are composed with warp-shuffle scans: pass A reduces a 16 KiB tile to one function,
// pass B scans the tile functions (one CTA), pass C re-scans inside the tile and writes.
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dd {

constexpr int kPackThreads = 256;
constexpr int kPackRows = 1;  // 4 KiB tiles: ~1200 CTAs for a 5 MB genome keep every SM busy
constexpr int kRowBytes = kPackThreads * 16;      // 4096
constexpr int kTileBytes = kRowBytes * kPackRows;
constexpr int kScanThreads = 1024;

struct PackWsHeader {
    uint32_t entry_last_byte;  // last text byte before this chunk (for pass C)
    uint32_t seg_len;          // tiles per pass-B thread segment
    uint64_t pad;
};
struct PackTileOut {
    uint64_t local_off;  // symbols emitted by earlier tiles of the same pass-B segment
    uint64_t state;      // header state at the start of the tile
};

// 16 text bytes at offset off (n = chunk length); beyond the end reads as inert padding.
__device__ __forceinline__ uint4 load_text16(const uint8_t *__restrict__ text, size_t off, size_t n, bool aligned) {
    if (aligned && off + 16 <= n) return __ldg(reinterpret_cast<const uint4 *>(text + off));
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t x = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const size_t q = off + 4 * i + b;
            const uint32_t c = (q < n) ? text[q] : (uint32_t)kPadByte;
            x |= c << (8 * b);
        }
        w[i] = x;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint64_t warp_scan_xfer(uint64_t f, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t g = __shfl_up_sync(0xffffffffu, f, d);
        if (lane >= d) f = xfer_compose(g, f);
    }
    return f;
}

// Exclusive scan of xfer functions over the CTA (blockDim.x threads, <= 1024); returns the
// composition of all earlier threads' functions and, through *total, of the whole CTA.
__device__ __forceinline__ uint64_t block_scan_xfer(uint64_t f, uint64_t *s_warp /*[32]*/, uint64_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    const uint64_t incl = warp_scan_xfer(f, lane);
    uint64_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = kXferIdentity;
    __syncthreads();  // s_warp may still be read by the previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t pre = kXferIdentity, all = kXferIdentity;
    for (int w = 0; w < nwarp; ++w) {
        const uint64_t t = s_warp[w];
        if (w < warp) pre = xfer_compose(pre, t);
        all = xfer_compose(all, t);
    }
    *total = all;
    return xfer_compose(pre, excl);
}

__device__ __forceinline__ bool chunk_at_line_start(const uint8_t *__restrict__ text, size_t off, uint32_t entry_last) {
    const uint32_t prev = off == 0 ? entry_last : (uint32_t)text[off - 1];
    return prev == '\n';
}

// ---- pass A: one transition function per tile -------------------------------------------------
__global__ void __launch_bounds__(kPackThreads)
pack_count_kernel(const uint8_t *__restrict__ text, size_t n, const dd_pack_state *__restrict__ st,
                  uint64_t *__restrict__ tile_xfer) {
    __shared__ uint64_t s_warp[32];
    const size_t base = (size_t)blockIdx.x * kTileBytes;
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    const uint32_t entry_last = st->last_byte;
    uint64_t tile = kXferIdentity;
#pragma unroll 1
    for (int r = 0; r < kPackRows; ++r) {
        const size_t off = base + (size_t)r * kRowBytes + (size_t)threadIdx.x * 16;
        uint64_t f = kXferIdentity;
        if (off < n) {
            const uint4 v = load_text16(text, off, n, aligned);
            const ChunkMasks m = classify16(v.x, v.y, v.z, v.w);
            f = chunk_xfer(m, chunk_at_line_start(text, off, entry_last));
        }
        uint64_t row;
        block_scan_xfer(f, s_warp, &row);
        tile = xfer_compose(tile, row);
    }
    if (threadIdx.x == 0) tile_xfer[blockIdx.x] = tile;
}

// ---- pass B: scan tile functions, assign output offsets, advance the stream state --------------
__global__ void __launch_bounds__(kScanThreads)
pack_scan_kernel(const uint8_t *__restrict__ text, size_t n, const uint64_t *__restrict__ tile_xfer, size_t ntiles,
                 PackTileOut *__restrict__ tile_out, uint64_t *__restrict__ seg_base, PackWsHeader *__restrict__ hdr,
                 dd_pack_state *__restrict__ st, size_t cap_symbols) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_sum[1];
    const size_t seg_len = (ntiles + kScanThreads - 1) / kScanThreads;
    const size_t t0 = (size_t)threadIdx.x * seg_len;
    const size_t t1 = t0 + seg_len < ntiles ? t0 + seg_len : ntiles;
    // walk 1: header-state function of my segment (counts dropped so nothing can overflow)
    uint64_t f = kXferIdentity;
    for (size_t t = t0; t < t1; ++t) {
        const uint64_t g = tile_xfer[t];
        f = xfer_make(0, 0, xfer_end(g, xfer_end(f, 0)), xfer_end(g, xfer_end(f, 1)));
    }
    uint64_t all;
    const uint64_t pre = block_scan_xfer(f, s_warp, &all);
    const uint32_t entry_state = st->in_header;
    uint32_t state = xfer_end(pre, entry_state);
    // walk 2: counts under the now-known states
    uint64_t local = 0;
    for (size_t t = t0; t < t1; ++t) {
        const uint64_t g = tile_xfer[t];
        tile_out[t].local_off = local;
        tile_out[t].state = state;
        local += xfer_cnt(g, state);
        state = xfer_end(g, state);
    }
    // exclusive scan of the per-thread symbol counts (warp shuffles + one pass over warp totals)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint64_t incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        __syncthreads();  // s_warp was last read inside block_scan_xfer
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint64_t before = 0, total = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) {
            const uint64_t t = s_warp[w];
            if (w < warp) before += t;
            total += t;
        }
        seg_base[threadIdx.x] = before + incl - local;
        s_sum[0] = total;  // every thread writes the same value
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t run = s_sum[0];
        hdr->entry_last_byte = st->last_byte;
        hdr->seg_len = (uint32_t)seg_len;
        const uint64_t before = st->nsym;
        st->prev_nsym = before;
        st->nsym = before + run;
        st->in_header = xfer_end(all, entry_state);
        if (n > 0) st->last_byte = text[n - 1];
        if (before + run > cap_symbols) st->reserved |= 1;  // overflow: symbols past capacity are dropped
    }
}

// ---- pass C: emit symbols ---------------------------------------------------------------------
__global__ void __launch_bounds__(kPackThreads)
pack_write_kernel(const uint8_t *__restrict__ text, size_t n, const PackTileOut *__restrict__ tile_out,
                  const uint64_t *__restrict__ seg_base, const PackWsHeader *__restrict__ hdr,
                  const dd_pack_state *__restrict__ st, uint32_t *__restrict__ codes, uint32_t *__restrict__ invalid,
                  size_t cap_symbols) {
    __shared__ uint64_t s_warp[32];
    __shared__ __align__(16) uint8_t s_stage[kRowBytes + 64];
    const size_t tile = blockIdx.x;
    const size_t base = tile * kTileBytes;
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    const uint32_t entry_last = hdr->entry_last_byte;
    uint32_t row_state = (uint32_t)tile_out[tile].state;
    uint64_t g_row = st->prev_nsym + seg_base[tile / hdr->seg_len] + tile_out[tile].local_off;
    const size_t cap_words16 = (cap_symbols + 15) >> 4, cap_words32 = (cap_symbols + 31) >> 5;

#pragma unroll 1
    for (int r = 0; r < kPackRows; ++r) {
        const size_t off = base + (size_t)r * kRowBytes + (size_t)threadIdx.x * 16;
        ChunkMasks m = {0, 0, 0, 0, 0};
        bool ls = false;
        uint64_t f = kXferIdentity;
        if (off < n) {
            const uint4 v = load_text16(text, off, n, aligned);
            m = classify16(v.x, v.y, v.z, v.w);
            ls = chunk_at_line_start(text, off, entry_last);
            f = chunk_xfer(m, ls);
        }
        uint64_t row;
        const uint64_t pre = block_scan_xfer(f, s_warp, &row);
        const uint32_t my_state = xfer_end(pre, row_state);
        const uint32_t my_off = xfer_cnt(pre, row_state);
        const uint32_t row_cnt = xfer_cnt(row, row_state);
        const uint32_t lead = (uint32_t)(g_row & 31);

        // zero the staging row (block_scan_xfer ended with a barrier-free read phase; the
        // barrier below orders the previous row's pack-out reads before these writes)
        __syncthreads();
        for (int i = threadIdx.x; i < (kRowBytes + 64) / 16; i += kPackThreads)
            reinterpret_cast<uint4 *>(s_stage)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        if (off < n) {
            const ChunkSyms cs = chunk_symbols(m, ls, my_state != 0);
            uint32_t rem = cs.sym, o = lead + my_off;
            while (rem) {
                const int i = __ffs((int)rem) - 1;
                rem &= rem - 1;
                s_stage[o++] = (uint8_t)(((m.codes >> (2 * i)) & 3u) | (((cs.brk >> i) & 1u) << 2));
            }
        }
        __syncthreads();
        // pack-out: one 32-symbol group per thread
        const uint32_t span = lead + row_cnt;  // staged extent
        const uint32_t ngroups = (span + 31) >> 5;
        const uint64_t g_base = g_row - lead;  // multiple of 32
        for (uint32_t g = threadIdx.x; g < ngroups; g += kPackThreads) {
            const uint4 a = reinterpret_cast<const uint4 *>(s_stage)[2 * g];
            const uint4 b = reinterpret_cast<const uint4 *>(s_stage)[2 * g + 1];
            const uint32_t c0 = (pack_codes4(a.x) << 24) | (pack_codes4(a.y) << 16) | (pack_codes4(a.z) << 8) | pack_codes4(a.w);
            const uint32_t c1 = (pack_codes4(b.x) << 24) | (pack_codes4(b.y) << 16) | (pack_codes4(b.z) << 8) | pack_codes4(b.w);
            const uint32_t iv = (pack_breaks4(a.x) << 28) | (pack_breaks4(a.y) << 24) | (pack_breaks4(a.z) << 20) |
                                (pack_breaks4(a.w) << 16) | (pack_breaks4(b.x) << 12) | (pack_breaks4(b.y) << 8) |
                                (pack_breaks4(b.z) << 4) | pack_breaks4(b.w);
            const uint32_t s0 = 32 * g;
            const uint64_t w32 = (g_base >> 5) + g;
            const uint64_t w16 = w32 * 2;
            // a word wholly produced by this row is stored; a word shared with a neighbouring row,
            // tile or chunk is OR-ed (buffers are zero-filled by dd_pack_reset)
            if (w16 < cap_words16) {
                if (s0 >= lead && s0 + 16 <= span) codes[w16] = c0;
                else if (c0) atomicOr(&codes[w16], c0);
            }
            if (w16 + 1 < cap_words16) {
                if (s0 + 16 >= lead && s0 + 32 <= span) codes[w16 + 1] = c1;
                else if (c1) atomicOr(&codes[w16 + 1], c1);
            }
            if (w32 < cap_words32) {
                if (s0 >= lead && s0 + 32 <= span) invalid[w32] = iv;
                else if (iv) atomicOr(&invalid[w32], iv);
            }
        }
        row_state = xfer_end(row, row_state);
        g_row += row_cnt;
    }
}

__global__ void pack_state_init_kernel(dd_pack_state *st) {
    st->nsym = 0;
    st->prev_nsym = 0;
    st->in_header = 0;
    st->last_byte = '\n';
    st->reserved = 0;
}

// ---- host side ----------------------------------------------------------------------------------
static size_t pack_ntiles(size_t n) { return (n + kTileBytes - 1) / kTileBytes; }
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t pack_workspace_bytes(size_t chunk_bytes) {
    const size_t nt = pack_ntiles(chunk_bytes) + 1;
    return align_up(sizeof(PackWsHeader), 256) + align_up(nt * sizeof(uint64_t), 256) +
           align_up(nt * sizeof(PackTileOut), 256) + align_up(kScanThreads * sizeof(uint64_t), 256);
}

cudaError_t pack_reset(uint32_t *d_codes, size_t codes_bytes, uint32_t *d_invalid, size_t invalid_bytes,
                       dd_pack_state *d_state, cudaStream_t stream) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_codes, 0, codes_bytes, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d_invalid, 0, invalid_bytes, stream)) != cudaSuccess) return e;
    pack_state_init_kernel<<<1, 1, 0, stream>>>(d_state);
    return cudaGetLastError();
}

cudaError_t pack_fasta(const uint8_t *d_text, size_t n, uint32_t *d_codes, uint32_t *d_invalid, size_t cap_symbols,
                       dd_pack_state *d_state, void *d_ws, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const size_t nt = pack_ntiles(n);
    uint8_t *p = static_cast<uint8_t *>(d_ws);
    PackWsHeader *hdr = reinterpret_cast<PackWsHeader *>(p);
    p += align_up(sizeof(PackWsHeader), 256);
    uint64_t *tile_xfer = reinterpret_cast<uint64_t *>(p);
    p += align_up((nt + 1) * sizeof(uint64_t), 256);
    PackTileOut *tile_out = reinterpret_cast<PackTileOut *>(p);
    p += align_up((nt + 1) * sizeof(PackTileOut), 256);
    uint64_t *seg_base = reinterpret_cast<uint64_t *>(p);

    pack_count_kernel<<<(unsigned)nt, kPackThreads, 0, stream>>>(d_text, n, d_state, tile_xfer);
    pack_scan_kernel<<<1, kScanThreads, 0, stream>>>(d_text, n, tile_xfer, nt, tile_out, seg_base, hdr, d_state,
                                                    cap_symbols);
    pack_write_kernel<<<(unsigned)nt, kPackThreads, 0, stream>>>(d_text, n, tile_out, seg_base, hdr, d_state, d_codes,
                                                                d_invalid, cap_symbols);
    return cudaGetLastError();
}

}  // namespace dd
