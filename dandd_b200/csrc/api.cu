// api.cu -- the C ABI declared in include/dandd_b200.h: argument checking, error reporting and
// the host-buffer convenience path.  No arithmetic lives here.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    return fail(DD_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define DD_CUDA(call, what)                              \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, what); \
    } while (0)

cudaStream_t S(dd_stream s) { return static_cast<cudaStream_t>(s); }
size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
bool bad_p(int p) { return p < 5 || p > 26; }
int popc(uint32_t x) { return __builtin_popcount(x); }

}  // namespace

namespace dd {
unsigned long long g_kernel_launches = 0;
}

extern "C" {

unsigned long long dd_kernel_launches(void) { return __atomic_load_n(&dd::g_kernel_launches, __ATOMIC_RELAXED); }
const char *dd_last_error(void) { return g_err; }
int dd_abi_version(void) { return DD_ABI_VERSION; }

int dd_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(DD_ERR_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(DD_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    DD_CUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(DD_ERR_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major,
                    prop.minor);
    DD_CUDA(cudaSetDevice(device), "cudaSetDevice");
    return DD_OK;
}

int dd_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes, size_t *total_mem) {
    cudaDeviceProp prop;
    DD_CUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return DD_OK;
}

int dd_set_option(const char *name, long value) {
    if (!name) return fail(DD_ERR_ARG, "dd_set_option: null name");
    if (!strcmp(name, "sketch_k_per_pass")) {
        if (value < 0 || value > 32) return fail(DD_ERR_ARG, "dd_set_option: sketch_k_per_pass must be 0..32");
        dd::g_k_per_pass = (int)value;
        return DD_OK;
    }
    if (!strcmp(name, "sketch_midk")) {
        dd::g_midk = value != 0;
        return DD_OK;
    }
    if (!strcmp(name, "prefix_planes")) {
        dd::g_prefix_planes = value != 0;
        return DD_OK;
    }
    if (!strcmp(name, "polyt_sentinel")) {
        dd::g_polyt_sentinel = value != 0;
        return DD_OK;
    }
    return fail(DD_ERR_ARG, "dd_set_option: unknown option '%s'", name);
}

// ---- K1 ------------------------------------------------------------------------------------------
size_t dd_pack_codes_bytes(size_t max_text_bytes) { return align_up(max_text_bytes / 4 + 64, 256); }
size_t dd_pack_invalid_bytes(size_t max_text_bytes) { return align_up(max_text_bytes / 8 + 64, 256); }
size_t dd_pack_workspace_bytes(size_t chunk_bytes) { return dd::pack_workspace_bytes(chunk_bytes); }

int dd_pack_reset(uint32_t *d_codes, size_t codes_bytes, uint32_t *d_invalid, size_t invalid_bytes,
                  dd_pack_state *d_state, dd_stream stream) {
    if (!d_codes || !d_invalid || !d_state) return fail(DD_ERR_ARG, "dd_pack_reset: null pointer");
    DD_CUDA(dd::pack_reset(d_codes, codes_bytes, d_invalid, invalid_bytes, d_state, S(stream)), "dd_pack_reset");
    return DD_OK;
}

int dd_pack_fasta(const uint8_t *d_text, size_t n_bytes, uint32_t *d_codes, uint32_t *d_invalid, size_t cap_symbols,
                  dd_pack_state *d_state, void *d_ws, size_t ws_bytes, dd_stream stream) {
    if (n_bytes == 0) return DD_OK;
    if (!d_text || !d_codes || !d_invalid || !d_state || !d_ws) return fail(DD_ERR_ARG, "dd_pack_fasta: null pointer");
    if (n_bytes > ((size_t)1 << 36)) return fail(DD_ERR_ARG, "dd_pack_fasta: chunk larger than 64 GiB");
    if (ws_bytes < dd::pack_workspace_bytes(n_bytes))
        return fail(DD_ERR_WORKSPACE, "dd_pack_fasta: workspace %zu < %zu", ws_bytes, dd::pack_workspace_bytes(n_bytes));
    DD_CUDA(dd::pack_fasta(d_text, n_bytes, d_codes, d_invalid, cap_symbols, d_state, d_ws, S(stream)), "dd_pack_fasta");
    return DD_OK;
}

int dd_pack_polyt_sentinel(const uint32_t *d_codes, uint32_t *d_invalid, const dd_pack_state *d_state, uint64_t sym_begin,
                           uint64_t sym_end, size_t max_symbols, dd_stream stream) {
    if (!d_codes || !d_invalid) return fail(DD_ERR_ARG, "dd_pack_polyt_sentinel: null pointer");
    if (!d_state && sym_end < sym_begin) return fail(DD_ERR_ARG, "dd_pack_polyt_sentinel: end < begin");
    DD_CUDA(dd::pack_polyt_sentinel(d_codes, d_invalid, d_state, sym_begin, sym_end, max_symbols, S(stream)),
            "dd_pack_polyt_sentinel");
    return DD_OK;
}

// ---- host-side text helpers (no GPU work) ---------------------------------------------------------
size_t dd_fasta_first_record_host(const uint8_t *h_text, size_t n_bytes) {
    if (!h_text) return 0;
    // kseq looks for '>' or '@' (klib kseq.h: `while ((c = ks_getc(ks)) != -1 && c != '>' && c != '@');`)
    const uint8_t *gt = static_cast<const uint8_t *>(memchr(h_text, '>', n_bytes));
    const size_t lim = gt ? (size_t)(gt - h_text) : n_bytes;   // '@' only matters if it comes first
    const uint8_t *at = static_cast<const uint8_t *>(memchr(h_text, '@', lim));
    return at ? (size_t)(at - h_text) : lim;
}

// The walk of kseq_read() over FASTA/FASTQ text, emitting plain FASTA: ">\n" per record, sequence
// lines verbatim.  One record per iteration of the outer loop.
size_t dd_fastq_to_fasta_host(const uint8_t *h_in, size_t n, uint8_t *h_out) {
    if (!h_in || !h_out) return 0;
    const uint8_t *p = h_in, *const end = h_in + n;
    uint8_t *o = h_out;
    auto line_end = [&](const uint8_t *from) {   // one past the line's '\n', or `end`
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(from, '\n', (size_t)(end - from)));
        return nl ? nl + 1 : end;
    };
    bool marker_taken = false;   // the previous record ended on the next record's marker
    while (true) {
        if (!marker_taken) {
            p += dd_fasta_first_record_host(p, (size_t)(end - p));
            if (p >= end) break;
            ++p;
        }
        marker_taken = false;
        uint8_t *const record_out = o;
        *o++ = '>';
        *o++ = '\n';
        p = line_end(p);   // name and comment
        size_t seq_len = 0;
        int stop = -1;
        while (p < end) {
            const uint8_t c = *p;
            if (c == '>' || c == '@' || c == '+') {
                stop = c;
                ++p;
                break;
            }
            const uint8_t *le = line_end(p);
            for (const uint8_t *q = p; q < le; ++q)
                if (*q != '\n' && *q != '\r') ++seq_len;
            memcpy(o, p, (size_t)(le - p));
            o += le - p;
            p = le;
        }
        if (o > record_out + 2 && o[-1] != '\n') *o++ = '\n';   // text ended without a newline
        if (stop == '>' || stop == '@') {
            marker_taken = true;
            continue;
        }
        if (stop != '+') break;   // end of text
        const uint8_t *le = line_end(p);   // the rest of the '+' line
        if (le == end && (le == p || le[-1] != '\n')) {   // no quality string: kseq_read fails
            o = record_out;
            break;
        }
        p = le;
        size_t qual_len = 0;
        bool first = true;
        while (p < end && (first || qual_len < seq_len)) {
            le = line_end(p);
            for (const uint8_t *q = p; q < le; ++q)
                if (*q != '\n' && *q != '\r') ++qual_len;
            p = le;
            first = false;
        }
        if (qual_len != seq_len) {   // kseq_read returns -2: this record and the rest are not read
            o = record_out;
            break;
        }
    }
    return (size_t)(o - h_out);
}

// ---- K2 ------------------------------------------------------------------------------------------
size_t dd_sketch_workspace_bytes(int nk, int p) { return dd::sketch_workspace_bytes(nk, p); }

int dd_sketch_begin(void *d_ws, size_t ws_bytes, int nk, int p, dd_stream stream) {
    if (!d_ws || nk < 1 || nk > 32 || bad_p(p)) return fail(DD_ERR_ARG, "dd_sketch_begin: bad argument (nk=%d p=%d)", nk, p);
    if (ws_bytes < dd::sketch_workspace_bytes(nk, p)) return fail(DD_ERR_WORKSPACE, "dd_sketch_begin: workspace too small");
    DD_CUDA(dd::sketch_begin(d_ws, nk, p, S(stream)), "dd_sketch_begin");
    return DD_OK;
}

static int sketch_check(const void *c, const void *i, uint32_t kmask, int p, void *ws, size_t ws_bytes, const char *fn) {
    if (!c || !i || !ws) return fail(DD_ERR_ARG, "%s: null pointer", fn);
    if (kmask == 0) return fail(DD_ERR_ARG, "%s: empty k mask", fn);
    if (bad_p(p)) return fail(DD_ERR_ARG, "%s: p=%d outside [5,26]", fn, p);
    if (ws_bytes < dd::sketch_workspace_bytes(popc(kmask), p)) return fail(DD_ERR_WORKSPACE, "%s: workspace too small", fn);
    return DD_OK;
}

int dd_sketch_update(const uint32_t *d_codes, const uint32_t *d_invalid, const dd_pack_state *d_state,
                     size_t max_new_symbols, uint32_t kmask, int p, int canon, void *d_ws, size_t ws_bytes,
                     dd_stream stream) {
    if (int rc = sketch_check(d_codes, d_invalid, kmask, p, d_ws, ws_bytes, "dd_sketch_update")) return rc;
    if (!d_state) return fail(DD_ERR_ARG, "dd_sketch_update: null state");
    DD_CUDA(dd::sketch_update(d_codes, d_invalid, d_state, 0, 0, max_new_symbols, kmask, p, canon, d_ws, S(stream)),
            "dd_sketch_update");
    return DD_OK;
}

int dd_sketch_update_range(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end,
                           uint32_t kmask, int p, int canon, void *d_ws, size_t ws_bytes, dd_stream stream) {
    if (int rc = sketch_check(d_codes, d_invalid, kmask, p, d_ws, ws_bytes, "dd_sketch_update_range")) return rc;
    if (sym_end < sym_begin) return fail(DD_ERR_ARG, "dd_sketch_update_range: end < begin");
    DD_CUDA(dd::sketch_update(d_codes, d_invalid, nullptr, sym_begin, sym_end, 0, kmask, p, canon, d_ws, S(stream)),
            "dd_sketch_update_range");
    return DD_OK;
}

int dd_sketch_update_sched(const uint32_t *d_codes, const uint32_t *d_invalid, const dd_pack_state *d_state, uint64_t sym_begin,
                           uint64_t sym_end, size_t max_new_symbols, uint64_t seen_before, uint32_t kmask, int p, int canon,
                           void *d_ws, size_t ws_bytes, dd_stream stream) {
    if (int rc = sketch_check(d_codes, d_invalid, kmask, p, d_ws, ws_bytes, "dd_sketch_update_sched")) return rc;
    if (!d_state && sym_end < sym_begin) return fail(DD_ERR_ARG, "dd_sketch_update_sched: end < begin");
    DD_CUDA(dd::sketch_update_sched(d_codes, d_invalid, d_state, sym_begin, sym_end, max_new_symbols, seen_before, kmask, p, canon,
                                    d_ws, S(stream)),
            "dd_sketch_update_sched");
    return DD_OK;
}

int dd_sketch_refresh_floor(void *d_ws, size_t ws_bytes, uint32_t kmask, int p, dd_stream stream) {
    if (!d_ws || kmask == 0 || bad_p(p)) return fail(DD_ERR_ARG, "dd_sketch_refresh_floor: bad argument");
    if (ws_bytes < dd::sketch_workspace_bytes(popc(kmask), p)) return fail(DD_ERR_WORKSPACE, "dd_sketch_refresh_floor: workspace too small");
    DD_CUDA(dd::sketch_refresh_floor(d_ws, kmask, p, S(stream)), "dd_sketch_refresh_floor");
    return DD_OK;
}

int dd_sketch_end(void *d_ws, size_t ws_bytes, int nk, int p, uint8_t *d_regs, uint32_t *d_hist, double *d_cards,
                  dd_stream stream) {
    if (!d_ws || !d_regs || nk < 1 || nk > 32 || bad_p(p)) return fail(DD_ERR_ARG, "dd_sketch_end: bad argument");
    if (d_cards && !d_hist) return fail(DD_ERR_ARG, "dd_sketch_end: d_cards needs d_hist");
    if (ws_bytes < dd::sketch_workspace_bytes(nk, p)) return fail(DD_ERR_WORKSPACE, "dd_sketch_end: workspace too small");
    DD_CUDA(dd::sketch_end(d_ws, nk, p, d_regs, d_hist, d_cards, S(stream)), "dd_sketch_end");
    return DD_OK;
}

// ---- K4 / K3 / K6 ----------------------------------------------------------------------------------
int dd_card_ertl_mle(const uint8_t *d_regs, int nsk, int p, double *d_cards, uint32_t *d_hist, dd_stream stream) {
    if (nsk == 0) return DD_OK;
    if (!d_regs || !d_cards || !d_hist || nsk < 0 || bad_p(p)) return fail(DD_ERR_ARG, "dd_card_ertl_mle: bad argument");
    DD_CUDA(dd::card_hist(d_regs, nsk, p, d_hist, S(stream)), "dd_card_ertl_mle(hist)");
    DD_CUDA(dd::mle_from_hist(d_hist, nsk, p, d_cards, S(stream)), "dd_card_ertl_mle(mle)");
    return DD_OK;
}

int dd_mle_from_hist(const uint32_t *d_hist, int nsk, int p, double *d_cards, dd_stream stream) {
    if (nsk == 0) return DD_OK;
    if (!d_hist || !d_cards || nsk < 0 || bad_p(p)) return fail(DD_ERR_ARG, "dd_mle_from_hist: bad argument");
    DD_CUDA(dd::mle_from_hist(d_hist, nsk, p, d_cards, S(stream)), "dd_mle_from_hist");
    return DD_OK;
}

int dd_union_max(const uint8_t *const *d_in, int n_in, size_t len, uint8_t *d_out, dd_stream stream) {
    if (!d_in || !d_out || n_in < 1) return fail(DD_ERR_ARG, "dd_union_max: bad argument");
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) != 0) return fail(DD_ERR_ARG, "dd_union_max: output must be 16-byte aligned");
    DD_CUDA(dd::union_max(d_in, n_in, len, d_out, S(stream)), "dd_union_max");
    return DD_OK;
}

size_t dd_prefix_union_workspace_bytes(int n_ord, int n_steps, int n_genomes, int nk, int p) {
    if (n_ord < 0 || n_steps < 0 || n_genomes < 0 || nk < 0 || bad_p(p)) return 0;
    return dd::prefix_union_workspace_bytes(n_ord, n_steps, n_genomes, nk, p);
}
size_t dd_planes_bytes(int64_t n_sketches, int p) {
    if (n_sketches < 0 || bad_p(p)) return 0;
    return dd::planes_bytes(n_sketches, p);
}

int dd_to_planes(const uint8_t *d_regs, int64_t n_sketches, int p, uint32_t *d_planes, dd_stream stream) {
    if (n_sketches == 0) return DD_OK;
    if (!d_regs || !d_planes || n_sketches < 0 || bad_p(p) || !dd::planes_supported(p))
        return fail(DD_ERR_ARG, "dd_to_planes: bad argument (bit planes need p >= 12)");
    DD_CUDA(dd::to_planes(d_regs, n_sketches, p, d_planes, S(stream)), "dd_to_planes");
    return DD_OK;
}

static int prefix_args_ok(const void *data, const int32_t *d_order, const double *d_cards, const uint32_t *d_hist, int n_ord,
                          int n_steps, int n_genomes, int nk, int p) {
    return data && d_order && d_cards && d_hist && n_ord >= 0 && n_steps >= 0 && n_genomes >= 1 && nk >= 1 && nk <= 65535 &&
           !bad_p(p) && (((size_t)1 << p) + 32767) / 32768 <= 65535;
}

int dd_prefix_union_card(const uint8_t *d_regs, const int32_t *d_order, int n_ord, int n_steps, int n_genomes, int nk,
                         int p, int final_only, double *d_cards, uint32_t *d_hist, uint8_t *d_unions, void *d_ws,
                         size_t ws_bytes, dd_stream stream) {
    if (n_ord == 0 || n_steps == 0) return DD_OK;
    if (!prefix_args_ok(d_regs, d_order, d_cards, d_hist, n_ord, n_steps, n_genomes, nk, p))
        return fail(DD_ERR_ARG, "dd_prefix_union_card: bad argument");
    const size_t rows = (size_t)n_ord * (final_only ? 1 : n_steps) * nk;
    if (rows > 0x7fffffff) return fail(DD_ERR_ARG, "dd_prefix_union_card: too many (ordering, step, k) rows");
    if (!d_unions && d_ws && dd::g_prefix_planes && dd::planes_supported(p)) {
        void *base = reinterpret_cast<void *>(align_up(reinterpret_cast<uintptr_t>(d_ws), 256));
        const size_t lost = (size_t)(static_cast<uint8_t *>(base) - static_cast<uint8_t *>(d_ws));
        // pairs (two steps, final union only) never use the identical-prefix table: planes alone
        const bool pairs_only = final_only && n_steps == 2;
        const size_t need = dd::prefix_union_workspace_bytes(pairs_only ? 0 : n_ord, pairs_only ? 0 : n_steps, n_genomes, nk, p);
        if (ws_bytes < lost + need - 256)
            return fail(DD_ERR_WORKSPACE, "dd_prefix_union_card: workspace %zu < %zu", ws_bytes, need);
        DD_CUDA(dd::prefix_union_hist_planes(d_regs, d_order, n_ord, n_steps, n_genomes, nk, p, final_only, d_hist, base, S(stream)),
                "dd_prefix_union_card(planes)");
    } else {
        DD_CUDA(dd::prefix_union_hist(d_regs, d_order, n_ord, n_steps, n_genomes, nk, p, final_only, d_hist, d_unions, S(stream)),
                "dd_prefix_union_card(hist)");
    }
    DD_CUDA(dd::mle_from_hist(d_hist, (int)rows, p, d_cards, S(stream)), "dd_prefix_union_card(mle)");
    return DD_OK;
}

int dd_prefix_union_card_planes(const uint32_t *d_planes, const int32_t *d_order, int n_ord, int n_steps, int n_genomes,
                                int nk, int p, int final_only, double *d_cards, uint32_t *d_hist, void *d_ws, size_t ws_bytes,
                                dd_stream stream) {
    if (n_ord == 0 || n_steps == 0) return DD_OK;
    if (!prefix_args_ok(d_planes, d_order, d_cards, d_hist, n_ord, n_steps, n_genomes, nk, p) || !dd::planes_supported(p))
        return fail(DD_ERR_ARG, "dd_prefix_union_card_planes: bad argument (bit planes need p >= 12)");
    const size_t rows = (size_t)n_ord * (final_only ? 1 : n_steps) * nk;
    if (rows > 0x7fffffff) return fail(DD_ERR_ARG, "dd_prefix_union_card_planes: too many (ordering, step, k) rows");
    void *scratch = nullptr;
    if (d_ws) {
        scratch = reinterpret_cast<void *>(align_up(reinterpret_cast<uintptr_t>(d_ws), 256));
        const size_t lost = (size_t)(static_cast<uint8_t *>(scratch) - static_cast<uint8_t *>(d_ws));
        if (ws_bytes < lost + dd::prefix_union_workspace_bytes(n_ord, n_steps, 0, 0, p) - 256)
            return fail(DD_ERR_WORKSPACE, "dd_prefix_union_card_planes: workspace too small");
    }
    DD_CUDA(dd::prefix_union_hist_from_planes(d_planes, d_order, n_ord, n_steps, n_genomes, nk, p, final_only, d_hist, scratch,
                                              S(stream)),
            "dd_prefix_union_card_planes");
    DD_CUDA(dd::mle_from_hist(d_hist, (int)rows, p, d_cards, S(stream)), "dd_prefix_union_card_planes(mle)");
    return DD_OK;
}

int dd_union_sets_card(const uint8_t *const *d_members, int n_sets, int n_steps, int p, int final_only, double *d_cards,
                       uint32_t *d_hist, uint8_t *d_unions, dd_stream stream) {
    if (n_sets == 0 || n_steps == 0) return DD_OK;
    if (!d_members || !d_cards || !d_hist || n_sets < 0 || n_steps < 0 || bad_p(p))
        return fail(DD_ERR_ARG, "dd_union_sets_card: bad argument");
    const size_t rows = (size_t)n_sets * (final_only ? 1 : n_steps);
    if (rows > 0x7fffffff) return fail(DD_ERR_ARG, "dd_union_sets_card: too many (set, step) rows");
    DD_CUDA(dd::union_sets_hist(d_members, n_sets, n_steps, p, final_only, d_hist, d_unions, S(stream)),
            "dd_union_sets_card(hist)");
    DD_CUDA(dd::mle_from_hist(d_hist, (int)rows, p, d_cards, S(stream)), "dd_union_sets_card(mle)");
    return DD_OK;
}

static int pair_count_ok(int64_t n_pairs, int nk) {
    return n_pairs >= 0 && n_pairs <= 0x7fffffff / (2 * (int64_t)(nk > 0 ? nk : 1));
}

int dd_pairwise_union_card(const uint8_t *d_regs, int n_genomes, int nk, int p, const int32_t *d_pairs, int64_t n_pairs,
                           double *d_cards, uint32_t *d_hist, void *d_ws, size_t ws_bytes, dd_stream stream) {
    if (n_pairs == 0) return DD_OK;
    if (!pair_count_ok(n_pairs, nk)) return fail(DD_ERR_ARG, "dd_pairwise_union_card: bad pair count");
    // a pair is a 2-step ordering of which only the full union is estimated
    return dd_prefix_union_card(d_regs, d_pairs, (int)n_pairs, 2, n_genomes, nk, p, /*final_only=*/1, d_cards, d_hist,
                                nullptr, d_ws, ws_bytes, stream);
}

int dd_pairwise_union_card_planes(const uint32_t *d_planes, int n_genomes, int nk, int p, const int32_t *d_pairs,
                                  int64_t n_pairs, double *d_cards, uint32_t *d_hist, dd_stream stream) {
    if (n_pairs == 0) return DD_OK;
    if (!pair_count_ok(n_pairs, nk)) return fail(DD_ERR_ARG, "dd_pairwise_union_card_planes: bad pair count");
    return dd_prefix_union_card_planes(d_planes, d_pairs, (int)n_pairs, 2, n_genomes, nk, p, /*final_only=*/1, d_cards, d_hist,
                                       nullptr, 0, stream);
}

// ---- K5 ------------------------------------------------------------------------------------------
static int exact_check(int k, uint64_t capacity, const void *ws, size_t ws_bytes, const char *fn) {
    if (!ws) return fail(DD_ERR_ARG, "%s: null workspace", fn);
    if (k < 1 || k > DD_EXACT_MAXK) return fail(DD_ERR_ARG, "%s: k=%d outside [1,%d]", fn, k, DD_EXACT_MAXK);
    if (k > DD_EXACT_BITMAP_MAXK && (capacity < 1024 || (capacity & (capacity - 1))))
        return fail(DD_ERR_ARG, "%s: capacity must be a power of two >= 1024", fn);
    if (ws_bytes < dd::exact_workspace_bytes(k, capacity)) return fail(DD_ERR_WORKSPACE, "%s: workspace too small", fn);
    return DD_OK;
}
size_t dd_exact_workspace_bytes(int k, uint64_t capacity) { return dd::exact_workspace_bytes(k, capacity); }

int dd_exact_begin(void *d_ws, size_t ws_bytes, int k, uint64_t capacity, dd_stream stream) {
    if (int rc = exact_check(k, capacity, d_ws, ws_bytes, "dd_exact_begin")) return rc;
    DD_CUDA(dd::exact_begin(d_ws, k, capacity, S(stream)), "dd_exact_begin");
    return DD_OK;
}
int dd_exact_insert_shard(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end, int k,
                          int canon, void *d_ws, size_t ws_bytes, uint64_t capacity, uint32_t shard_rank,
                          uint32_t shard_world, dd_stream stream) {
    if (!d_codes || !d_invalid) return fail(DD_ERR_ARG, "dd_exact_insert: null pointer");
    if (shard_world < 1 || shard_rank >= shard_world || shard_world > 65536)
        return fail(DD_ERR_ARG, "dd_exact_insert: shard %u of %u", shard_rank, shard_world);
    if (int rc = exact_check(k, capacity, d_ws, ws_bytes, "dd_exact_insert")) return rc;
    DD_CUDA(dd::exact_insert(d_codes, d_invalid, sym_begin, sym_end, k, canon, d_ws, capacity, shard_rank, shard_world,
                             S(stream)),
            "dd_exact_insert");
    return DD_OK;
}
int dd_exact_insert(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end, int k,
                    int canon, void *d_ws, size_t ws_bytes, uint64_t capacity, dd_stream stream) {
    return dd_exact_insert_shard(d_codes, d_invalid, sym_begin, sym_end, k, canon, d_ws, ws_bytes, capacity, 0, 1, stream);
}
int dd_exact_count(void *d_ws, size_t ws_bytes, int k, uint64_t capacity, uint64_t *d_count, dd_stream stream) {
    if (!d_count) return fail(DD_ERR_ARG, "dd_exact_count: null pointer");
    if (int rc = exact_check(k, capacity, d_ws, ws_bytes, "dd_exact_count")) return rc;
    DD_CUDA(dd::exact_count(d_ws, k, capacity, d_count, S(stream)), "dd_exact_count");
    return DD_OK;
}

// ---- host-buffer path -----------------------------------------------------------------------------
namespace {
constexpr size_t kHostChunk = (size_t)32 << 20;  // text bytes per H2D / pack / sketch round
struct HostWs {
    uint8_t *text;
    uint32_t *codes;
    uint32_t *invalid;
    dd_pack_state *state;
    void *pack_ws;
    void *sketch_ws;
    uint8_t *regs;
    uint32_t *hist;
    double *cards;
    size_t codes_bytes, invalid_bytes, pack_ws_bytes, sketch_ws_bytes, total;
};
HostWs carve(void *base, size_t n_bytes, int nk, int p) {
    HostWs w;
    uint8_t *q = static_cast<uint8_t *>(base);
    auto take = [&](size_t bytes) {
        uint8_t *r = q;
        q += align_up(bytes, 256);
        return r;
    };
    const size_t chunk = n_bytes < kHostChunk ? n_bytes : kHostChunk;
    w.codes_bytes = dd_pack_codes_bytes(n_bytes);
    w.invalid_bytes = dd_pack_invalid_bytes(n_bytes);
    w.pack_ws_bytes = dd::pack_workspace_bytes(chunk);
    w.sketch_ws_bytes = dd::sketch_workspace_bytes(nk, p);
    w.text = take(2 * align_up(chunk + 16, 256));  // double buffer
    w.codes = reinterpret_cast<uint32_t *>(take(w.codes_bytes));
    w.invalid = reinterpret_cast<uint32_t *>(take(w.invalid_bytes));
    w.state = reinterpret_cast<dd_pack_state *>(take(sizeof(dd_pack_state)));
    w.pack_ws = take(w.pack_ws_bytes);
    w.sketch_ws = take(w.sketch_ws_bytes);
    w.regs = take((size_t)nk << p);
    w.hist = reinterpret_cast<uint32_t *>(take((size_t)nk * DD_HIST_BINS * sizeof(uint32_t)));
    w.cards = reinterpret_cast<double *>(take((size_t)nk * sizeof(double)));
    w.total = (size_t)(q - static_cast<uint8_t *>(base));
    return w;
}
}  // namespace

size_t dd_sketch_fasta_host_workspace_bytes(size_t n_bytes, int nk, int p) {
    return carve(nullptr, n_bytes, nk, p).total + 256;
}

// cards <- NaN when the packer saw a line beginning with '+': the registers built from FASTQ text
// without dd_fastq_to_fasta_host are not what kseq would have produced, and an asynchronous caller
// has no other way of learning it.
__global__ void fastq_poison_kernel(const dd_pack_state *st, double *cards, int nk) {
    if ((st->reserved & DD_PACK_FLAG_FASTQ) && threadIdx.x < (unsigned)nk) cards[threadIdx.x] = nan("");
}

static int sketch_fasta_host_impl(const uint8_t *h_text, size_t n_bytes, uint32_t kmask, int p, int canon, uint8_t *h_regs,
                                  double *h_cards, uint8_t *d_regs_or_null, void *d_ws, size_t ws_bytes, dd_stream stream,
                                  bool synchronize) {
    const int nk = popc(kmask);
    if (!h_text || !h_cards || !d_ws || nk == 0 || bad_p(p)) return fail(DD_ERR_ARG, "dd_sketch_fasta_host: bad argument");
    if (ws_bytes < dd_sketch_fasta_host_workspace_bytes(n_bytes, nk, p))
        return fail(DD_ERR_WORKSPACE, "dd_sketch_fasta_host: workspace %zu < %zu", ws_bytes,
                    dd_sketch_fasta_host_workspace_bytes(n_bytes, nk, p));
    void *base = reinterpret_cast<void *>(align_up(reinterpret_cast<uintptr_t>(d_ws), 256));
    HostWs w = carve(base, n_bytes, nk, p);
    uint8_t *d_regs = d_regs_or_null ? d_regs_or_null : w.regs;
    cudaStream_t st = S(stream);

    // kseq ignores everything before the first record marker
    const size_t skip = dd_fasta_first_record_host(h_text, n_bytes);
    const uint8_t *text = h_text + skip;
    const size_t n = n_bytes - skip;

    DD_CUDA(dd::pack_reset(w.codes, w.codes_bytes, w.invalid, w.invalid_bytes, w.state, st), "pack_reset");
    DD_CUDA(dd::sketch_begin(w.sketch_ws, nk, p, st), "sketch_begin");
    const size_t chunk = n_bytes < kHostChunk ? n_bytes : kHostChunk;
    const size_t nchunks = n ? (n + chunk - 1) / chunk : 0;
    // More than one chunk: copies run on their own stream, double-buffered against pack + sketch.
    // The stream and the events are created here and destroyed before returning WITHOUT waiting:
    // CUDA releases a destroyed stream / event once the work already enqueued on it has completed,
    // so the call stays asynchronous however large the file is.
    cudaStream_t cs = st;
    cudaEvent_t copied[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr}, entry = nullptr;
    const bool overlap = nchunks > 1;
    int rc = DD_OK;
    auto cleanup = [&]() {
        for (int i = 0; i < 2; ++i) {
            if (copied[i]) cudaEventDestroy(copied[i]);
            if (freed[i]) cudaEventDestroy(freed[i]);
        }
        if (entry) cudaEventDestroy(entry);
        if (overlap && cs != st) cudaStreamDestroy(cs);
    };
#define DD_TRY(call, what)                                   \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) { rc = cuda_fail(e__, what); cleanup(); return rc; } \
    } while (0)
    if (overlap) {
        DD_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking), "create copy stream");
        for (int i = 0; i < 2; ++i) {
            DD_TRY(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming), "event");
            DD_TRY(cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming), "event");
        }
        DD_TRY(cudaEventCreateWithFlags(&entry, cudaEventDisableTiming), "event");
        DD_TRY(cudaEventRecord(entry, st), "event record");      // workspace may still be in use upstream
        DD_TRY(cudaStreamWaitEvent(cs, entry, 0), "stream wait");
    }
    size_t done = 0;
    for (size_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        const size_t len = n - done < chunk ? n - done : chunk;
        uint8_t *d_text = w.text + (size_t)b * align_up(chunk + 16, 256);
        if (overlap && c >= 2) DD_TRY(cudaStreamWaitEvent(cs, freed[b], 0), "stream wait");
        DD_TRY(cudaMemcpyAsync(d_text, text + done, len, cudaMemcpyHostToDevice, cs), "H2D text");
        if (overlap) {
            DD_TRY(cudaEventRecord(copied[b], cs), "event record");
            DD_TRY(cudaStreamWaitEvent(st, copied[b], 0), "stream wait");
        }
        DD_TRY(dd::pack_fasta(d_text, len, w.codes, w.invalid, n_bytes, w.state, w.pack_ws, st), "pack_fasta");
        if (overlap) DD_TRY(cudaEventRecord(freed[b], st), "event record");
        if (dd::g_polyt_sentinel) DD_TRY(dd::pack_polyt_sentinel(w.codes, w.invalid, w.state, 0, 0, len, st), "polyt_sentinel");
        // (text bytes stand in for symbols in the floor schedule: the cuts move by ~1 %, the result does not)
        DD_TRY(dd::sketch_update_sched(w.codes, w.invalid, w.state, 0, 0, len, done, kmask, p, canon, w.sketch_ws, st),
               "sketch_update");
        done += len;
    }
    cleanup();
#undef DD_TRY
    DD_CUDA(dd::sketch_end(w.sketch_ws, nk, p, d_regs, w.hist, w.cards, st), "sketch_end");
    DD_COUNT_LAUNCH(), fastq_poison_kernel<<<1, 32, 0, st>>>(w.state, w.cards, nk);
    DD_CUDA(cudaGetLastError(), "fastq_poison");
    DD_CUDA(cudaMemcpyAsync(h_cards, w.cards, (size_t)nk * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H cards");
    if (h_regs) DD_CUDA(cudaMemcpyAsync(h_regs, d_regs, (size_t)nk << p, cudaMemcpyDeviceToHost, st), "D2H regs");
    if (synchronize) {
        dd_pack_state hs;
        DD_CUDA(cudaMemcpyAsync(&hs, w.state, sizeof hs, cudaMemcpyDeviceToHost, st), "D2H state");
        DD_CUDA(cudaStreamSynchronize(st), "synchronize");
        if (hs.reserved & DD_PACK_FLAG_FASTQ) return DD_ERR_FORMAT;   // the caller normalises and retries
    }
    return DD_OK;
}

int dd_sketch_fasta_host(const uint8_t *h_text, size_t n_bytes, uint32_t kmask, int p, int canon, uint8_t *h_regs,
                         double *h_cards, uint8_t *d_regs_or_null, void *d_ws, size_t ws_bytes, dd_stream stream) {
    int rc = sketch_fasta_host_impl(h_text, n_bytes, kmask, p, canon, h_regs, h_cards, d_regs_or_null, d_ws, ws_bytes, stream,
                                    true);
    if (rc != DD_ERR_FORMAT) return rc;
    // FASTQ (a line begins with '+'): rewrite to FASTA the way kseq walks it, then sketch that
    std::vector<uint8_t> fasta(n_bytes + 16);
    const size_t m = dd_fastq_to_fasta_host(h_text, n_bytes, fasta.data());
    rc = sketch_fasta_host_impl(fasta.data(), m, kmask, p, canon, h_regs, h_cards, d_regs_or_null, d_ws, ws_bytes, stream,
                                true);
    if (rc == DD_ERR_FORMAT) return fail(DD_ERR_FORMAT, "dd_sketch_fasta_host: text still looks like FASTQ after normalisation");
    return rc;
}

int dd_sketch_fasta_host_async(const uint8_t *h_text, size_t n_bytes, uint32_t kmask, int p, int canon, uint8_t *h_regs,
                               double *h_cards, uint8_t *d_regs_or_null, void *d_ws, size_t ws_bytes, dd_stream stream) {
    return sketch_fasta_host_impl(h_text, n_bytes, kmask, p, canon, h_regs, h_cards, d_regs_or_null, d_ws, ws_bytes, stream,
                                  false);
}

}  // extern "C"
