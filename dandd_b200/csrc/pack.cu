// pack.cu -- K1: streaming FASTA text -> 2-bit packed symbol stream (+ break bit mask).
//
// Replaces the FASTA parsing half of `dashing sketch` / `kmc -fm` (reference
// lib/sketch_classes.py:351-366, 434-449).  HBM-streaming kernel: every text byte is read with
// 128-bit loads (twice: a counting pass and a writing pass; the chunk is L2-resident for the second
// when the host streams in <= 64 MiB chunks), output is 0.375 B/symbol, staged through shared
// memory so that the global stores are whole, coalesced words.  A thread owns 64 contiguous text
// bytes and a CTA a 16 KiB tile, so there is one block scan per 16 KiB of text.  Both text passes
// are persistent kernels (SM count x resident CTAs) that stage their tiles in shared memory with
// the bulk-copy engine (`cp.async.bulk.shared::cluster.global` + mbarrier complete_tx, SASS
// UBLKCP.S.G), double-buffered: tile i+1 streams in while tile i is classified, scanned and packed.
// Text that is not 16-byte aligned is staged by a cooperative byte copy instead (same results).
//
// The only sequential dependence in FASTA text is "am I inside a header line?".  A stretch of text
// is summarised as a transition function over that one bit plus the symbol counts under both
// hypotheses (dd::chunk_xfer, packed in a u64) and the functions are composed with warp-shuffle
// scans: pass A reduces a 16 KiB tile to one function, pass B1 scans the tiles inside groups of
// 512, pass B2 scans the groups (one CTA), pass C re-scans inside the tile and writes.
//
// Two paths per tile.  A tile that holds only letters and '\n' (every tile of a sequence body;
// one SWAR test per word decides, and pass A records it in the tile's function) takes the FAST
// path: no per-byte state machine, the symbols of a 16-byte chunk are validated and 2-bit encoded
// in registers, the newline holes squeezed out, and the result appended to little-endian bit
// streams in shared memory.  Anything else -- headers, CR, blanks, digits, the last partial tile, a
// tile entered inside a header -- takes the GENERAL path built on dd::chunk_symbols.
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dd {

constexpr int kPackThreads = 256;
constexpr int kSpanChunks = 4;                         // 16-byte chunks per thread, contiguous in the text
constexpr int kSpanBytes = 16 * kSpanChunks;           // 64
constexpr int kTileBytes = kPackThreads * kSpanBytes;  // 16 KiB per CTA, one block scan per tile
constexpr int kScanThreads = 1024;

struct PackWsHeader {
    uint32_t entry_last_byte;  // last text byte before this chunk (for pass C)
    uint32_t seg_len;          // groups per pass-B2 thread segment
    uint64_t pad;
};
struct PackTileOut {
    uint64_t local_off;  // symbols emitted by earlier groups of the same pass-B2 segment
    uint64_t state;      // header state at the start of the group
};

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

// A thread's 64-byte span: the masks of its four chunks, whether each chunk starts a line, and
// the span's transition function.  All four 128-bit loads are issued before any is used.
struct Span {
    ChunkMasks m[kSpanChunks];
    uint32_t ls;  // bit c: chunk c starts at a line start
    uint32_t fastq;  // non-zero: the span holds a line that begins with '+'
    uint64_t f;
};

// ---- bulk-copy (TMA) staging of text tiles -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one thread: expect `bytes` on the barrier and start the copy global -> shared (bytes % 16 == 0)
__device__ __forceinline__ void tma_load_tile(uint8_t *dst, const uint8_t *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    if (bytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                     "l"(src), "r"(bytes), "r"(smem_u32(bar))
                     : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}

// Tile pipeline shared by both passes: a CTA walks tiles blockIdx.x, +gridDim.x, ...; buffer i&1
// holds tile i.  stage_issue() is called by thread 0 only, stage_wait() by everyone; the caller's
// end-of-iteration __syncthreads() is what makes a buffer reusable two iterations later.
struct TilePipe {
    uint8_t *buf[2];
    uint64_t *bar;  // [2]
    const uint8_t *text;
    size_t n;
    bool tma;       // text is 16-byte aligned

    __device__ __forceinline__ uint32_t tile_bytes(size_t tile) const {
        const size_t base = tile * kTileBytes;
        return (uint32_t)(n - base < (size_t)kTileBytes ? n - base : (size_t)kTileBytes);
    }
    __device__ __forceinline__ void issue(size_t tile, int it) const {
        if (tma) tma_load_tile(buf[it & 1], text + tile * kTileBytes, tile_bytes(tile) & ~15u, &bar[it & 1]);
    }
    // after this returns every thread may read bytes [0, roundup64(tile_bytes)) of buf[it&1]
    __device__ __forceinline__ const uint8_t *wait(size_t tile, int it) const {
        uint8_t *b = buf[it & 1];
        const uint32_t nb = tile_bytes(tile);
        const uint8_t *src = text + tile * kTileBytes;
        if (tma) {
            mbar_wait(&bar[it & 1], (uint32_t)(it >> 1) & 1u);
            if (nb < (uint32_t)kTileBytes) {  // last tile: the < 16 trailing bytes, then inert padding
                const uint32_t i = (nb & ~15u) + threadIdx.x;
                if (i < ((nb + kSpanBytes - 1) & ~(uint32_t)(kSpanBytes - 1))) b[i] = i < nb ? src[i] : kPadByte;
                __syncthreads();
            }
        } else {  // unaligned text: cooperative byte-safe copy
            const uint32_t lim = (nb + kSpanBytes - 1) & ~(uint32_t)(kSpanBytes - 1);
            for (uint32_t i = threadIdx.x; i < lim; i += kPackThreads) b[i] = i < nb ? src[i] : kPadByte;
            __syncthreads();
        }
        return b;
    }
};

// A thread's span read from the staged tile (shared memory).
__device__ __forceinline__ Span load_span_smem(const uint8_t *tile, uint32_t off_in_tile, uint32_t prev_byte) {
    Span sp;
    sp.f = kXferIdentity;
    sp.ls = 0;
    sp.fastq = 0;
    uint4 v[kSpanChunks];
#pragma unroll
    for (int c = 0; c < kSpanChunks; ++c) v[c] = *reinterpret_cast<const uint4 *>(tile + off_in_tile + 16 * c);
    uint32_t prev = prev_byte;
#pragma unroll
    for (int c = 0; c < kSpanChunks; ++c) {
        sp.m[c] = classify16(v[c].x, v[c].y, v[c].z, v[c].w);
        const bool ls = prev == '\n';
        sp.ls |= (ls ? 1u : 0u) << c;
        sp.f = xfer_compose(sp.f, chunk_xfer(sp.m[c], ls));
        sp.fastq |= chunk_fastq_marks(sp.m[c], ls);
        prev = v[c].w >> 24;
    }
    return sp;
}

// byte before a thread's span: previous span's last byte (same tile), else the text before the tile
__device__ __forceinline__ uint32_t span_prev_byte(const uint8_t *tile, uint32_t off_in_tile, const uint8_t *text,
                                                   size_t tile_base, uint32_t entry_last) {
    if (off_in_tile) return tile[off_in_tile - 1];
    return tile_base ? (uint32_t)text[tile_base - 1] : entry_last;
}

// ---- fast path for "simple" tiles ------------------------------------------------------------------
// A tile is SIMPLE when every byte is either '\n' or has bit 6 set and a non-zero low part (letters:
// 0x41-0x7F, 0xC1-0xFF; '@' = 0x40 is excluded because a line-initial '@' is a record marker):
// no '>' / '@' (so no header can start), no '+', no '\r', no blanks/digits, no padding.  Then the only non-symbol
// bytes are the newlines, the transition function follows from three block reductions, and pass C
// needs no per-byte scatter: each 16-byte chunk is validated and 2-bit encoded with SWAR, its
// newline holes are squeezed out in registers and the result is OR-ed into little-endian bit streams
// in shared memory.  Everything else (headers, CRLF, the last partial tile) takes the general path.
constexpr uint64_t kXferSimple = (uint64_t)1 << 62;

// Per-thread summary of a 64-byte span: newline count and whether every byte without bit 6 is a
// newline (the two byte-wise counters agree exactly then, because a newline itself lacks bit 6).
struct SimpleSpan {
    uint32_t nnl;
    bool simple;
};

__device__ __forceinline__ SimpleSpan scan_simple(const uint8_t *tile, uint32_t off) {
    uint32_t acc_nl = 0, acc_low = 0;  // four byte-wide counters each, <= 16 per byte
    uint32_t at = 0;                   // bit 7 of a byte set if (byte & 0x3f) == 0 there or in a lower byte
#pragma unroll
    for (int c = 0; c < kSpanChunks; ++c) {
        const uint4 v = *reinterpret_cast<const uint4 *>(tile + off + 16 * c);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc_nl += bytes_eq(w[i], 0x0a0a0a0au) >> 7;
            acc_low += (~w[i] & 0x40404040u) >> 6;
            at |= (w[i] & 0x3f3f3f3fu) - 0x01010101u;   // zero-byte test; a borrow only adds false alarms
        }
    }
    SimpleSpan sp;
    sp.simple = acc_nl == acc_low && (at & 0x80808080u) == 0u;
    sp.nnl = (acc_nl * 0x01010101u) >> 24;
    return sp;
}

// byte offset of the first '\n' in a 64-byte span that is known to contain one
__device__ __forceinline__ uint32_t first_newline(const uint8_t *tile, uint32_t off) {
    for (int i = 0; i < kSpanBytes / 4; ++i) {
        const uint32_t z = bytes_eq(*reinterpret_cast<const uint32_t *>(tile + off + 4 * i), 0x0a0a0a0au);
        if (z) return 4 * i + ((__ffs((int)z) - 1) >> 3);
    }
    return kSpanBytes;
}

// block-wide exclusive prefix sum of a small count (blockDim.x == kPackThreads)
__device__ __forceinline__ uint32_t block_scan_u32(uint32_t v, uint32_t *s_warp32 /*[8]*/, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    __syncthreads();
    if (lane == 31) s_warp32[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kPackThreads / 32; ++w) {
        const uint32_t t = s_warp32[w];
        if (w < warp) before += t;
        all += t;
    }
    *total = all;
    return before + incl - v;
}

// ---- pass A: one transition function per tile -------------------------------------------------
__global__ void __launch_bounds__(kPackThreads)
pack_count_kernel(const uint8_t *__restrict__ text, size_t n, size_t ntiles, dd_pack_state *__restrict__ st,
                  uint64_t *__restrict__ tile_xfer) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    __shared__ uint64_t s_warp[32];
    __shared__ __align__(8) uint64_t s_bar[2];
    TilePipe pipe;
    pipe.buf[0] = s_dyn;
    pipe.buf[1] = s_dyn + kTileBytes;
    pipe.bar = s_bar;
    pipe.text = text;
    pipe.n = n;
    pipe.tma = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0]);
        mbar_init(&s_bar[1]);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t entry_last = st->last_byte;
    if (threadIdx.x == 0 && blockIdx.x < ntiles) pipe.issue(blockIdx.x, 0);
    int it = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const size_t next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) pipe.issue(next, it + 1);
        const uint8_t *b = pipe.wait(tile, it);
        const uint32_t off = threadIdx.x * kSpanBytes;
        const bool full_tile = pipe.tile_bytes(tile) == (uint32_t)kTileBytes;
        SimpleSpan ss = {0u, false};
        if (full_tile) ss = scan_simple(b, off);
        if (__syncthreads_and(full_tile && ss.simple)) {
            // symbols = every byte that is not a newline; if the text before the tile ended inside a
            // header, the bytes up to the first newline of the tile belong to that header
            uint32_t total;
            block_scan_u32(kSpanBytes - ss.nnl, reinterpret_cast<uint32_t *>(s_warp), &total);
            // the first newline of the tile matters only if the tile is entered inside a header: the
            // first thread that holds one looks its position up, nobody else does any work for it
            __shared__ uint32_t s_first[kPackThreads / 32];
            const uint32_t holders = __ballot_sync(0xffffffffu, ss.nnl != 0);
            if ((threadIdx.x & 31) == 0)
                s_first[threadIdx.x >> 5] = holders ? (threadIdx.x & ~31u) + (uint32_t)(__ffs((int)holders) - 1) : 0xffffffffu;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t t0 = 0xffffffffu;
                for (int w = 0; w < kPackThreads / 32; ++w) t0 = min(t0, s_first[w]);
                const bool has_nl = t0 != 0xffffffffu;
                // no non-symbol byte precedes the first newline, so it is preceded by f0 symbols
                const uint32_t f0 = has_nl ? t0 * kSpanBytes + first_newline(b, t0 * kSpanBytes) : 0u;
                tile_xfer[tile] = xfer_make(total, has_nl ? total - f0 : 0u, 0u, has_nl ? 0u : 1u) | kXferSimple;
            }
        } else {
            uint64_t f = kXferIdentity;
            if (off < pipe.tile_bytes(tile)) {
                const Span sp = load_span_smem(b, off, span_prev_byte(b, off, text, tile * kTileBytes, entry_last));
                f = sp.f;
                // a line that begins with '+': FASTQ.  Reported, not interpreted (see common.cuh).
                if (sp.fastq) atomicOr(reinterpret_cast<unsigned long long *>(&st->reserved), (unsigned long long)DD_PACK_FLAG_FASTQ);
            }
            uint64_t total;
            block_scan_xfer(f, s_warp, &total);
            if (threadIdx.x == 0) tile_xfer[tile] = total;
        }
        __syncthreads();  // buffer it&1 and s_warp are free again
    }
}

// ---- pass B1: scan tile functions inside groups of kGroupTiles tiles -----------------------------
// One CTA per group, one tile per thread: tile_xfer[t] becomes the composition of the earlier tiles
// of its group (exclusive prefix, simple flag kept), group_agg[g] the whole group.  A group emits at
// most kGroupTiles * kTileBytes = 2^23 symbols, which the 24-bit count fields hold.
constexpr int kGroupTiles = 512;

__global__ void __launch_bounds__(kGroupTiles)
pack_scan_groups_kernel(uint64_t *__restrict__ tile_xfer, size_t ntiles, uint64_t *__restrict__ group_agg,
                        const uint8_t *__restrict__ text, size_t n, PackTileOut *__restrict__ group_out,
                        uint64_t *__restrict__ seg_base, PackWsHeader *__restrict__ hdr, dd_pack_state *__restrict__ st,
                        size_t cap_symbols) {
    __shared__ uint64_t s_warp[32];
    const size_t t = (size_t)blockIdx.x * kGroupTiles + threadIdx.x;
    const uint64_t g = t < ntiles ? tile_xfer[t] : kXferIdentity;
    uint64_t all;
    const uint64_t pre = block_scan_xfer(g & ~kXferSimple, s_warp, &all);
    if (t < ntiles) tile_xfer[t] = pre | (g & kXferSimple);
    if (threadIdx.x == 0) {
        group_agg[blockIdx.x] = all;
        if (gridDim.x == 1) {
            // a single group (<= 8 MiB of text): there is nothing left to scan, so this kernel also does
            // pass B2's bookkeeping and the host skips that launch
            const uint32_t entry_state = st->in_header;
            group_out[0].local_off = 0;
            group_out[0].state = entry_state;
            seg_base[0] = 0;
            hdr->entry_last_byte = st->last_byte;
            hdr->seg_len = 1;
            const uint64_t before = st->nsym, run = xfer_cnt(all, entry_state);
            st->prev_nsym = before;
            st->nsym = before + run;
            st->in_header = xfer_end(all, entry_state);
            if (n > 0) st->last_byte = text[n - 1];
            if (before + run > cap_symbols) st->reserved |= DD_PACK_FLAG_OVERFLOW;  // symbols past capacity are dropped
        }
    }
}

// ---- pass B2: scan the group functions ("tiles" below are groups), assign output offsets,
// advance the stream state ----------------------------------------------------------------------
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
        if (before + run > cap_symbols) st->reserved |= DD_PACK_FLAG_OVERFLOW;  // symbols past capacity are dropped
    }
}

// ---- pass C: emit symbols ---------------------------------------------------------------------
constexpr size_t kWriteSmem = 2 * kTileBytes + (kTileBytes + 64);  // two text buffers + symbol staging
// PRMT without __byte_perm's selector masking (the selectors built here never set a nibble's bit 3)
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// high half of a 32x32 product, pinned to the FMA pipe: a right shift / multiply-gather off the ALU
__device__ __forceinline__ uint32_t mulhi_fma(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
constexpr int kFastCodeWords = 1040;   // >= (kTileBytes + 31) * 2 / 32 + spill word, multiple of 4
constexpr int kFastBreakWords = 520;   // >= (kTileBytes + 31) / 32 + spill word, multiple of 4
// little-endian 2-bit codes (symbol i at bits 2i+1:2i) -> the stream's big-endian order (31-2i:30-2i)
__device__ __forceinline__ uint32_t codes_le_to_be(uint32_t x) {
    const uint32_t r = brev32(x);
    return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

__global__ void __launch_bounds__(kPackThreads)
pack_write_kernel(const uint8_t *__restrict__ text, size_t n, size_t ntiles, const uint64_t *__restrict__ tile_pre,
                  const PackTileOut *__restrict__ group_out, const uint64_t *__restrict__ seg_base, const PackWsHeader *__restrict__ hdr,
                  const dd_pack_state *__restrict__ st, uint32_t *__restrict__ codes, uint32_t *__restrict__ invalid,
                  size_t cap_symbols) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    __shared__ uint64_t s_warp[32];
    __shared__ __align__(8) uint64_t s_bar[2];
    uint8_t *s_stage = s_dyn + 2 * kTileBytes;
    TilePipe pipe;
    pipe.buf[0] = s_dyn;
    pipe.buf[1] = s_dyn + kTileBytes;
    pipe.bar = s_bar;
    pipe.text = text;
    pipe.n = n;
    pipe.tma = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0]);
        mbar_init(&s_bar[1]);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t entry_last = hdr->entry_last_byte;
    const uint32_t seg_len = hdr->seg_len;
    const uint64_t stream_base = st->prev_nsym;
    const size_t cap_words16 = (cap_symbols + 15) >> 4, cap_words32 = (cap_symbols + 31) >> 5;
    if (threadIdx.x == 0 && blockIdx.x < ntiles) pipe.issue(blockIdx.x, 0);
    int it = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const size_t next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) pipe.issue(next, it + 1);
        const size_t group = tile / kGroupTiles;
        const uint64_t pre_fn = tile_pre[tile];  // earlier tiles of the group, as a function of the group's entry state
        const uint32_t group_state = (uint32_t)group_out[group].state;
        const uint32_t tile_state = xfer_end(pre_fn, group_state);
        const bool fast = (pre_fn & kXferSimple) && tile_state == 0;  // simple tile entered outside a header
        const uint64_t g_tile = stream_base + seg_base[group / seg_len] + group_out[group].local_off + xfer_cnt(pre_fn, group_state);
        const uint32_t lead = (uint32_t)(g_tile & 31);
        const uint32_t off = threadIdx.x * kSpanBytes;
        uint32_t tile_cnt;
        uint32_t *s_codes = reinterpret_cast<uint32_t *>(s_stage);  // fast path: little-endian bit streams
        uint32_t *s_brk = s_codes + kFastCodeWords;

        if (fast) {
            for (int i = threadIdx.x; i < (kFastCodeWords + kFastBreakWords) / 4; i += kPackThreads)
                reinterpret_cast<uint4 *>(s_stage)[i] = make_uint4(0, 0, 0, 0);
            const uint8_t *b = pipe.wait(tile, it);
            // In a simple tile a byte without bit 6 is a newline, so one LOP finds them.  The ACGT test
            // is classify_word's table lookup with the free slot 2 holding '\n': the XOR below is zero
            // exactly for ACGTacgt and '\n', so its non-zero bytes are the break symbols.  Newline
            // flags of a chunk are kept transposed (bit 8j+i = byte j of word i; the shifts that put
            // them there run on the FMA pipe, which is idle here) and only turned into symbol
            // positions by the few lanes that actually hold one.
            uint32_t Cw[kSpanChunks], Bw[kSpanChunks / 2], Xw[kSpanChunks / 2];  // codes; break masks; newline flags
            Bw[0] = Bw[1] = 0u;
#pragma unroll
            for (int c = 0; c < kSpanChunks; ++c) {
                const uint4 v = *reinterpret_cast<const uint4 *>(b + off + 16 * c);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                uint32_t d[4], C = 0, X = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t y = w[i] & 0x07070707u;
                    const uint32_t t = y | (y >> 4);
                    d[i] = (w[i] & 0xdfdfdfdfu) ^ prmt_raw(0x430a4101u, 0x47010154u, prmt_raw(t, t, 0x0020u));
                    X |= mulhi_fma(~w[i] & 0x40404040u, 1u << (26 + i));          // >> (6 - i)
                    const uint32_t cc = ((w[i] >> 1) ^ (w[i] >> 2)) & 0x03030303u;
                    C |= (((cc * 0x00041041u) >> 18) & 0xFFu) << (8 * i);
                }
                if (d[0] | d[1] | d[2] | d[3]) {  // a letter that is not ACGT: exact per-byte break mask
                    uint32_t B = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        B |= gather_bit7((((d[i] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d[i]) & 0x80808080u) << (4 * i);
                    Bw[c >> 1] |= B << (16 * (c & 1));
                }
                Cw[c] = C;
                if (c & 1) Xw[c >> 1] |= X << 4;
                else Xw[c >> 1] = X;
            }
            const uint32_t cnt = kSpanBytes - (uint32_t)(__popc(Xw[0]) + __popc(Xw[1]));
            // my first output position (the scan's barriers also order the zeroing above)
            const uint32_t pos0 = lead + block_scan_u32(cnt, reinterpret_cast<uint32_t *>(s_warp), &tile_cnt);
            if (cnt) {
                // code bits are collected in a register and OR-ed into the stream one full word at a
                // time (the tile's stream words are zero and the first/last word of a thread's range is
                // shared with its neighbours, so every flush is an atomic OR); break bits are rare and
                // go straight to the stream
                uint32_t pos = pos0, wc = pos0 >> 4, fill_c = 2 * (pos0 & 15), acc_c = 0;
#pragma unroll
                for (int c = 0; c < kSpanChunks; ++c) {
                    uint32_t C = Cw[c], B = (Bw[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
                    const uint32_t X = (Xw[c >> 1] >> (4 * (c & 1))) & 0x0F0F0F0Fu;
                    uint32_t n = 16;
                    if (X) {  // squeeze the newline positions out of both streams
                        uint32_t nl;
                        if (X & (X - 1)) {  // several newlines in 16 bytes: symbol-order mask, one at a time
                            nl = 0;
#pragma unroll
                            for (int i = 0; i < 4; ++i) nl |= ((((X >> i) & 0x01010101u) * 0x00204081u >> 21) & 0xFu) << (4 * i);
                        } else {
                            const int bit = __ffs((int)X) - 1;
                            nl = 1u << (4 * (bit & 7) + (bit >> 3));
                        }
                        n -= __popc(nl);
                        while (nl) {
                            const int h = __ffs((int)nl) - 1;
                            const uint32_t below = (1u << h) - 1u, below2 = (1u << (2 * h)) - 1u;
                            B = (B & below) | ((B >> (h + 1)) << h);
                            C = (C & below2) | (((C >> (2 * h + 1)) >> 1) << (2 * h));
                            nl = (nl & (nl - 1)) >> 1;
                        }
                    }
                    // append 2n code bits
                    acc_c |= C << fill_c;
                    if (fill_c + 2 * n >= 32) {
                        atomicOr(&s_codes[wc++], acc_c);
                        acc_c = __funnelshift_l(C, 0u, fill_c);  // C >> (32 - fill_c): the bits that did not fit
                        fill_c -= 32;
                    }
                    fill_c += 2 * n;
                    if (B) {
                        const uint32_t sh = pos & 31;
                        atomicOr(&s_brk[pos >> 5], B << sh);
                        if (sh > 16) atomicOr(&s_brk[(pos >> 5) + 1], B >> (32 - sh));
                    }
                    pos += n;
                }
                if (acc_c) atomicOr(&s_codes[wc], acc_c);
            }
        } else {
            // symbols are staged as bytes (bits 1:0 code, bit 2 break) at their position in the output
            // stream relative to the 32-symbol boundary below g_tile, so that pack-out writes whole words
            for (int i = threadIdx.x; i < (kTileBytes + 64) / 16; i += kPackThreads)
                reinterpret_cast<uint4 *>(s_stage)[i] = make_uint4(0, 0, 0, 0);

            const uint8_t *b = pipe.wait(tile, it);
            const bool active = off < pipe.tile_bytes(tile);
            Span sp;
            sp.f = kXferIdentity;
            if (active) sp = load_span_smem(b, off, span_prev_byte(b, off, text, tile * kTileBytes, entry_last));
            uint64_t all;
            const uint64_t pre = block_scan_xfer(sp.f, s_warp, &all);  // its barriers also order the zeroing above
            tile_cnt = xfer_cnt(all, tile_state);
            if (active) {
                uint32_t state = xfer_end(pre, tile_state);
                uint32_t o = lead + xfer_cnt(pre, tile_state);
#pragma unroll
                for (int c = 0; c < kSpanChunks; ++c) {
                    const ChunkSyms cs = chunk_symbols(sp.m[c], (sp.ls >> c) & 1u, state != 0);
                    uint32_t rem = cs.sym;
                    while (rem) {
                        const int i = __ffs((int)rem) - 1;
                        rem &= rem - 1;
                        s_stage[o++] = (uint8_t)(((sp.m[c].codes >> (2 * i)) & 3u) | (((cs.brk >> i) & 1u) << 2));
                    }
                    state = cs.end_hdr;
                }
            }
        }
        __syncthreads();
        // pack-out: one 32-symbol group per thread and iteration
        const uint32_t span = lead + tile_cnt;  // staged extent
        const uint32_t ngroups = (span + 31) >> 5;
        const uint64_t g_base = g_tile - lead;  // multiple of 32
        for (uint32_t g = threadIdx.x; g < ngroups; g += kPackThreads) {
            uint32_t c0, c1, iv;
            if (fast) {
                c0 = codes_le_to_be(s_codes[2 * g]);
                c1 = codes_le_to_be(s_codes[2 * g + 1]);
                iv = brev32(s_brk[g]);
            } else {
                const uint4 a = reinterpret_cast<const uint4 *>(s_stage)[2 * g];
                const uint4 bq = reinterpret_cast<const uint4 *>(s_stage)[2 * g + 1];
                c0 = (pack_codes4(a.x) << 24) | (pack_codes4(a.y) << 16) | (pack_codes4(a.z) << 8) | pack_codes4(a.w);
                c1 = (pack_codes4(bq.x) << 24) | (pack_codes4(bq.y) << 16) | (pack_codes4(bq.z) << 8) | pack_codes4(bq.w);
                iv = (pack_breaks4(a.x) << 28) | (pack_breaks4(a.y) << 24) | (pack_breaks4(a.z) << 20) |
                     (pack_breaks4(a.w) << 16) | (pack_breaks4(bq.x) << 12) | (pack_breaks4(bq.y) << 8) |
                     (pack_breaks4(bq.z) << 4) | pack_breaks4(bq.w);
            }
            const uint32_t s0 = 32 * g;
            const uint64_t w32 = (g_base >> 5) + g;
            const uint64_t w16 = w32 * 2;
            if (s0 >= lead && s0 + 32 <= span && w16 + 1 < cap_words16 && w32 < cap_words32) {
                *reinterpret_cast<uint2 *>(codes + w16) = make_uint2(c0, c1);  // wholly this tile's
                invalid[w32] = iv;
                continue;
            }
            // a word wholly produced by this tile is stored; a word shared with a neighbouring tile or
            // chunk is OR-ed (the buffers are zero-filled by dd_pack_reset)
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
        __syncthreads();  // text buffer it&1 and the staging area are free again
    }
}

// ---- optional: the all-ones accumulator quirk (UNVERIFIED, off by default) ---------
// Every 32nd T of a run of valid T symbols becomes a break.  One thread per 16-symbol word looks
// for run starts among its symbols; the thread that owns a run's start walks the run (runs of T
// are short in real text; a pathological all-T genome is walked by one thread -- this pass is an
// emulation switch, not part of the default path).  Breaks inserted by an earlier call (the range
// grows chunk by chunk) look like run boundaries to a later one, which yields the same positions
// because they are 32 apart by construction; a run that began before `begin` is continued by the
// thread of symbol `begin`, which first counts the at most 31 T behind it.
__device__ __forceinline__ bool sym_is_valid_t(const uint32_t *codes, const uint32_t *invalid, uint64_t s) {
    const uint32_t code = (codes[s >> 4] >> (30u - 2u * (uint32_t)(s & 15))) & 3u;
    const uint32_t bad = (invalid[s >> 5] >> (31u - (uint32_t)(s & 31))) & 1u;
    return code == 3u && !bad;
}
__global__ void __launch_bounds__(256)
pack_polyt_kernel(const uint32_t *__restrict__ codes, uint32_t *invalid, const dd_pack_state *__restrict__ st,
                  uint64_t sym_begin, uint64_t sym_end) {
    if (st) {
        sym_begin = st->prev_nsym;
        sym_end = st->nsym;
    }
    const uint64_t w = (sym_begin >> 4) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int j = 0; j < 16; ++j) {
        const uint64_t s = (w << 4) + j;
        if (s < sym_begin || s >= sym_end) continue;
        if (!sym_is_valid_t(codes, invalid, s)) continue;
        const bool prev_t = s > 0 && sym_is_valid_t(codes, invalid, s - 1);
        if (prev_t && s != sym_begin) continue;       // not a run start, and not the continuation point
        uint32_t run = 0;
        if (prev_t)                                    // s == sym_begin inside a run: count what lies behind
            for (uint64_t b = s; b > 0 && sym_is_valid_t(codes, invalid, b - 1); --b) run = (run + 1) & 31u;
        for (uint64_t t = s; t < sym_end && sym_is_valid_t(codes, invalid, t); ++t)
            if (++run == 32u) {
                atomicOr(&invalid[t >> 5], 1u << (31u - (uint32_t)(t & 31)));
                run = 0;
            }
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

static size_t pack_ngroups(size_t ntiles) { return (ntiles + kGroupTiles - 1) / kGroupTiles; }

size_t pack_workspace_bytes(size_t chunk_bytes) {
    const size_t nt = pack_ntiles(chunk_bytes) + 1, ng = pack_ngroups(nt) + 1;
    return align_up(sizeof(PackWsHeader), 256) + align_up(nt * sizeof(uint64_t), 256) + align_up(ng * sizeof(uint64_t), 256) +
           align_up(ng * sizeof(PackTileOut), 256) + align_up(kScanThreads * sizeof(uint64_t), 256);
}

cudaError_t pack_reset(uint32_t *d_codes, size_t codes_bytes, uint32_t *d_invalid, size_t invalid_bytes,
                       dd_pack_state *d_state, cudaStream_t stream) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_codes, 0, codes_bytes, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d_invalid, 0, invalid_bytes, stream)) != cudaSuccess) return e;
    DD_COUNT_LAUNCH(), pack_state_init_kernel<<<1, 1, 0, stream>>>(d_state);
    return cudaGetLastError();
}

cudaError_t pack_fasta(const uint8_t *d_text, size_t n, uint32_t *d_codes, uint32_t *d_invalid, size_t cap_symbols,
                       dd_pack_state *d_state, void *d_ws, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const size_t nt = pack_ntiles(n);
    uint8_t *p = static_cast<uint8_t *>(d_ws);
    PackWsHeader *hdr = reinterpret_cast<PackWsHeader *>(p);
    p += align_up(sizeof(PackWsHeader), 256);
    const size_t ng = pack_ngroups(nt);
    uint64_t *tile_xfer = reinterpret_cast<uint64_t *>(p);
    p += align_up((nt + 1) * sizeof(uint64_t), 256);
    uint64_t *group_agg = reinterpret_cast<uint64_t *>(p);
    p += align_up((pack_ngroups(nt + 1) + 1) * sizeof(uint64_t), 256);
    PackTileOut *group_out = reinterpret_cast<PackTileOut *>(p);
    p += align_up((pack_ngroups(nt + 1) + 1) * sizeof(PackTileOut), 256);
    uint64_t *seg_base = reinterpret_cast<uint64_t *>(p);

    // persistent grids: SM count x resident CTAs (shared memory decides), never more CTAs than tiles;
    // cached per device (the function attributes are per device as well)
    constexpr int kMaxDev = 64;
    static int grids[kMaxDev][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int local[2] = {0, 0};
    int &grid_count = (dev >= 0 && dev < kMaxDev) ? grids[dev][0] : local[0];
    int &grid_write = (dev >= 0 && dev < kMaxDev) ? grids[dev][1] : local[1];
    if (grid_count == 0) {
        int sms = 148, a = 1, c = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(pack_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kTileBytes));
        cudaFuncSetAttribute(pack_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWriteSmem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, pack_count_kernel, kPackThreads, 2 * kTileBytes);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, pack_write_kernel, kPackThreads, kWriteSmem);
        grid_count = sms * (a > 0 ? a : 1);
        grid_write = sms * (c > 0 ? c : 1);
    }
    const unsigned ga = (unsigned)(nt < (size_t)grid_count ? nt : (size_t)grid_count);
    const unsigned gc = (unsigned)(nt < (size_t)grid_write ? nt : (size_t)grid_write);
    DD_COUNT_LAUNCH(), pack_count_kernel<<<ga, kPackThreads, 2 * kTileBytes, stream>>>(d_text, n, nt, d_state, tile_xfer);
    DD_COUNT_LAUNCH(), pack_scan_groups_kernel<<<(unsigned)ng, kGroupTiles, 0, stream>>>(tile_xfer, nt, group_agg, d_text, n, group_out, seg_base,
                                                                      hdr, d_state, cap_symbols);
    if (ng > 1)
        DD_COUNT_LAUNCH(), pack_scan_kernel<<<1, kScanThreads, 0, stream>>>(d_text, n, group_agg, ng, group_out, seg_base, hdr, d_state,
                                                        cap_symbols);
    DD_COUNT_LAUNCH(), pack_write_kernel<<<gc, kPackThreads, kWriteSmem, stream>>>(d_text, n, nt, tile_xfer, group_out, seg_base, hdr, d_state,
                                                               d_codes, d_invalid, cap_symbols);
    return cudaGetLastError();
}

int g_polyt_sentinel = 0;

cudaError_t pack_polyt_sentinel(const uint32_t *d_codes, uint32_t *d_invalid, const dd_pack_state *d_state,
                                uint64_t sym_begin, uint64_t sym_end, size_t max_symbols, cudaStream_t stream) {
    const size_t nsym = d_state ? max_symbols : (size_t)(sym_end - sym_begin);
    if (nsym == 0) return cudaSuccess;
    const size_t nwords = (nsym + 15) / 16 + 2;   // the range may start and end inside a word
    DD_COUNT_LAUNCH(), pack_polyt_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, stream>>>(d_codes, d_invalid, d_state, sym_begin, sym_end);
    return cudaGetLastError();
}

}  // namespace dd
