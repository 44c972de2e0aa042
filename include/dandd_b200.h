/*
 * dandd_b200.h -- C ABI of the B200-native DandD sketch-and-count hot path.
 *
 * The reference (jessicabonnie/dandd) has NO FFI: its seam to the arithmetic is a set of shell
 * command strings built by SketchObj subclasses and run through subprocess.  Each entry point
 * below replaces one of those command lines; the citation gives the reference site (paths are
 * relative to the reference repository root).  INTEGRATION.md shows the ctypes stubs a reference
 * maintainer would add at those sites.
 *
 * Conventions
 *   - Every function returns DD_OK (0) or a negative DD_ERR_* code; dd_last_error() returns a
 *     thread-local message for the last failure.  There is no CPU fallback: on a machine without
 *     an sm_100 device dd_init() fails and nothing else may be called.
 *   - The caller owns every buffer.  Pointers named d_* are DEVICE pointers (e.g. torch tensors'
 *     data_ptr()), pointers named h_* are HOST pointers.  The library allocates nothing that
 *     outlives a call; scratch space is passed in as (d_ws, ws_bytes) sized by the matching
 *     *_workspace_bytes() query.
 *   - All device work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream) and is asynchronous unless the name ends in _host.
 *   - k values travel as a bit mask: bit (k-1) set <=> k-mer length k is requested, 1 <= k <= 32
 *     (`maxk<=32 for estimation`, README.md:82; lib/huffman_dandd.py:109-110).  "slot" j of a
 *     [nk][...] output is the j-th set bit of the mask in increasing k.
 *   - A sketch is 2^p one-byte HyperLogLog registers (p = --registers, default 20,
 *     lib/dandd_cmd.py:187), bit-identical to what `dashing sketch -S p` builds.
 */
#ifndef DANDD_B200_H
#define DANDD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_ABI_VERSION 2

#define DD_OK 0
#define DD_ERR_ARG (-1)       /* bad argument (null pointer, k/p out of range, ...) */
#define DD_ERR_CUDA (-2)      /* a CUDA runtime call or kernel launch failed        */
#define DD_ERR_WORKSPACE (-3) /* workspace smaller than *_workspace_bytes() says    */
#define DD_ERR_DEVICE (-4)    /* no usable sm_100 device                            */
#define DD_ERR_FORMAT (-5)    /* input text needs dd_fastq_to_fasta_host first      */

#define DD_HIST_BINS 64 /* register-value histogram: bin j = #registers == j (j <= 64-p+1 < 64) */

typedef void *dd_stream;

#if defined(__GNUC__)
#define DD_API __attribute__((visibility("default")))
#else
#define DD_API
#endif

DD_API const char *dd_last_error(void);
DD_API int dd_abi_version(void);
/* Number of CUDA kernels this library has launched in this process so far (all threads, all streams;
 * memsets and copies are not kernels).  Lets a benchmark report its launches as a measurement. */
DD_API unsigned long long dd_kernel_launches(void);
/* Select `device` for the calling thread and verify it is compute capability 10.x. */
DD_API int dd_init(int device);
/* Tuning knobs (never change results).  "sketch_k_per_pass" = n: K2 updates at most n k values per
 * kernel launch (0 = all in one), trading re-reads of the packed stream for L2 residency of the
 * accumulators.  "prefix_planes" = 0/1: dd_prefix_union_card uses the bit-sliced kernel (default 1)
 * or the byte kernel.  "sketch_midk" = 0/1: dd_sketch_update_sched lets pieces deep inside a long stream
 * consult L2-resident presence bitmaps for k = 10..12 before hashing (default 1).  "polyt_sentinel" = 0/1: the dd_*_host entry points apply dd_pack_polyt_sentinel
 * (this one DOES change results: it selects the SURVEY.md A.6 encoder behaviour; default 0). */
DD_API int dd_set_option(const char *name, long value);
DD_API int dd_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes, size_t *total_mem);

/* ===========================================================================================
 * K1  FASTA text -> packed symbol stream.
 * Replaces the FASTA parsing half of `dashing sketch ... <fasta>` (lib/sketch_classes.py:351-366)
 * and of `kmc ... -fm <fasta>` (lib/sketch_classes.py:434-449).
 *
 * Stream layout (device memory, caller-owned, zero-filled by dd_pack_reset):
 *   codes   : uint32 words, 16 symbols per word, symbol s in word s/16 at bits [31-2(s%16) : 30-2(s%16)]
 *             (first symbol most significant), A/a=0 C/c=1 G/g=2 T/t=3;
 *   invalid : uint32 words, 32 symbols per word, symbol s in word s/32 at bit 31-(s%32); a set bit
 *             means "k-mer windows restart here" (any byte other than ACGTacgt, and one synthetic
 *             symbol per '>' header line so that k-mers never span records);
 *   state   : dd_pack_state, carried from chunk to chunk so a FASTA can be streamed.
 * Text rules (klib kseq_read() as used by Dashing and by KMC's -fm mode): a line
 * whose first byte is '>' or '@' is a record header; '\n' and '\r' are not sequence; text before
 * the first '>' / '@' of a file must be skipped by the caller (dd_*_host do it).  A line whose
 * first byte is '+' opens a FASTQ quality section, whose extent depends on the record's sequence
 * length: the packer does not interpret it, it sets DD_PACK_FLAG_FASTQ in dd_pack_state.reserved and
 * the caller must pass such text through dd_fastq_to_fasta_host first (dd_sketch_fasta_host does).
 * =========================================================================================== */
#define DD_PACK_FLAG_OVERFLOW 1u /* more symbols than cap_symbols: the excess was dropped            */
#define DD_PACK_FLAG_FASTQ 2u    /* a line begins with '+': FASTQ, results are NOT what kseq yields  */
typedef struct dd_pack_state {
    uint64_t nsym;      /* symbols in the stream so far                                   */
    uint64_t prev_nsym; /* value of nsym before the most recent dd_pack_fasta call        */
    uint32_t in_header; /* 1 if the text consumed so far ends inside a header line        */
    uint32_t last_byte; /* last text byte consumed ('\n' initially: start of a line)      */
    uint64_t reserved;  /* DD_PACK_FLAG_* bits (sticky)                                   */
} dd_pack_state;

/* HOST helper, no GPU work: rewrite FASTA/FASTQ text the way kseq_read() walks it into plain
 * FASTA that dd_pack_fasta packs to the same symbols -- record markers found anywhere while no
 * record is open, '+' lines and quality lines (as many bytes as the record has sequence bytes, at
 * least one line) dropped, a record with a quality/sequence length mismatch and everything after
 * it dropped (kseq_read returns an error there and the reader loop ends).  h_out must hold
 * n_bytes + 16 bytes and may not overlap h_in; returns the number of bytes written. */
DD_API size_t dd_fastq_to_fasta_host(const uint8_t *h_in, size_t n_bytes, uint8_t *h_out);
/* Offset of the first record marker ('>' or '@') in host text, n_bytes if there is none. */
DD_API size_t dd_fasta_first_record_host(const uint8_t *h_text, size_t n_bytes);

DD_API size_t dd_pack_codes_bytes(size_t max_text_bytes);
DD_API size_t dd_pack_invalid_bytes(size_t max_text_bytes);
DD_API size_t dd_pack_workspace_bytes(size_t chunk_bytes);
DD_API int dd_pack_reset(uint32_t *d_codes, size_t codes_bytes, uint32_t *d_invalid, size_t invalid_bytes,
                  dd_pack_state *d_state, dd_stream stream);
/* Append the symbols of text chunk [d_text, d_text+n_bytes) to the stream.  d_text should be
 * 16-byte aligned for full speed.  cap_symbols = capacity of codes/invalid in symbols. */
DD_API int dd_pack_fasta(const uint8_t *d_text, size_t n_bytes, uint32_t *d_codes, uint32_t *d_invalid,
                  size_t cap_symbols, dd_pack_state *d_state, void *d_ws, size_t ws_bytes, dd_stream stream);

/* Optional emulation of the encoder quirk recalled in SURVEY.md A.6 (UNVERIFIED, off by default):
 * bonsai's unwindowed encoder tests its 64-bit accumulator against all-ones to detect an invalid
 * base, and 32 consecutive 'T' make a legitimate accumulator all-ones too.  With this pass every
 * 32nd 'T' of a run of T/t becomes a break symbol, for symbols [sym_begin, sym_end) -- or
 * [d_state->prev_nsym, d_state->nsym) when d_state != NULL (max_symbols bounds the grid) -- so
 * call it after each dd_pack_fasta, before dd_sketch_update.  dd_set_option("polyt_sentinel", 1)
 * makes the dd_*_host entry points do so. */
DD_API int dd_pack_polyt_sentinel(const uint32_t *d_codes, uint32_t *d_invalid, const dd_pack_state *d_state, uint64_t sym_begin,
                           uint64_t sym_end, size_t max_symbols, dd_stream stream);

/* ===========================================================================================
 * K2  fused all-k canonical-k-mer HyperLogLog sketch.
 * Replaces the whole fan-out
 *     parallel -j 95% ' dashing sketch [--no-canon] -k{} -S p --prefix DIR fasta ' ::: k1 k2 ...
 * (lib/huffman_dandd.py:217 building lib/sketch_classes.py:351-366): one pass over the packed
 * stream updates the registers of every requested k.
 *   begin  : zero the accumulators held in the workspace;
 *   update : sketch stream symbols [state.prev_nsym, state.nsym) (k-mers ENDING in that range, so
 *            chunked packing + update is exact), or an explicit [sym_begin, sym_end) range;
 *   end    : write the final u8 registers [nk][2^p], and optionally the register histograms and
 *            Ertl-MLE cardinalities (what `dashing card` would print for each of the nk sketches).
 * canon = 1 hashes min(kmer, reverse complement) (default), 0 = --no-canon
 * (lib/sketch_classes.py:20-29).
 * =========================================================================================== */
DD_API size_t dd_sketch_workspace_bytes(int nk, int p);
DD_API int dd_sketch_begin(void *d_ws, size_t ws_bytes, int nk, int p, dd_stream stream);
DD_API int dd_sketch_update(const uint32_t *d_codes, const uint32_t *d_invalid, const dd_pack_state *d_state,
                     size_t max_new_symbols, uint32_t kmask, int p, int canon, void *d_ws, size_t ws_bytes,
                     dd_stream stream);
DD_API int dd_sketch_update_range(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end,
                           uint32_t kmask, int p, int canon, void *d_ws, size_t ws_bytes, dd_stream stream);
/* dd_sketch_update plus the floor schedule: the per-k lower bound min(register) rises by one every
 * time the number of k-mers seen doubles (first at about 14 x 2^p), and an update that cannot exceed
 * it is dropped before the reduction.  This entry cuts the range at 16 x 2^p x 2^i symbols counted
 * from the start of the sketch and refreshes the bound at each cut.  seen_before = symbols sketched
 * into this workspace by earlier calls (a host-side count; an estimate only moves the cuts).  With
 * d_state != NULL the range is the last dd_pack_fasta call's and [sym_begin, sym_end) is ignored. */
DD_API int dd_sketch_update_sched(const uint32_t *d_codes, const uint32_t *d_invalid, const dd_pack_state *d_state,
                           uint64_t sym_begin, uint64_t sym_end, size_t max_new_symbols, uint64_t seen_before,
                           uint32_t kmask, int p, int canon, void *d_ws, size_t ws_bytes, dd_stream stream);
/* Recompute the per-k lower bound min(register) used to skip no-op updates (long genomes). */
DD_API int dd_sketch_refresh_floor(void *d_ws, size_t ws_bytes, uint32_t kmask, int p, dd_stream stream);
DD_API int dd_sketch_end(void *d_ws, size_t ws_bytes, int nk, int p, uint8_t *d_regs, uint32_t *d_hist /*[nk][64] or NULL*/,
                  double *d_cards /*[nk] or NULL*/, dd_stream stream);

/* ===========================================================================================
 * K4  cardinality: register histogram + Ertl maximum-likelihood estimate.
 * Replaces `dashing card --presketched <paths...>` (lib/sketch_classes.py:306-321); d_cards[i] is
 * the number the reference would store in cardkey[path_i].  d_hist ([nsk][64] uint32) receives the
 * histograms and doubles as scratch; it must be provided.
 * =========================================================================================== */
DD_API int dd_card_ertl_mle(const uint8_t *d_regs, int nsk, int p, double *d_cards, uint32_t *d_hist, dd_stream stream);
/* Estimator alone, from histograms (host-callable building block; also used by K3/K6). */
DD_API int dd_mle_from_hist(const uint32_t *d_hist, int nsk, int p, double *d_cards, dd_stream stream);

/* ===========================================================================================
 * K3  unions.
 * dd_union_max replaces `dashing union -z -o OUT in1 .. inN` (lib/sketch_classes.py:368-373):
 * d_in is a DEVICE array of n_in device pointers, each to `len` bytes; d_out = element-wise max.
 *
 * dd_prefix_union_card replaces the progressive-union loop (lib/huffman_dandd.py:624-663): for
 * every ordering o and step i it forms the union of genomes order[o][0..i] and its cardinality
 * for each of the nk sketches per genome -- as a running max, n reads instead of the reference's
 * n(n+1)/2.  Entries of d_order < 0 are skipped (ragged orderings).
 *   d_regs   [n_genomes][nk][2^p] u8         d_order [n_ord][n_steps] int32
 *   d_cards  [n_ord][n_steps][nk] f64        d_hist  [n_ord][n_steps][nk][64] u32 (scratch+output)
 *   d_unions NULL, or [n_ord][n_steps][nk][2^p] u8 to materialise every prefix union
 *   final_only != 0: only the union of each whole ordering is estimated/materialised -- the n-ary
 *   union of a tree node (lib/huffman_dandd.py:412-438); the step dimension of the outputs then
 *   has extent 1: d_cards [n_ord][nk], d_hist [n_ord][nk][64], d_unions [n_ord][nk][2^p].
 * =========================================================================================== */
/* dd_union_sets_card: the same running-max + histogram + MLE over arbitrary sketches given by
 * address.  d_members is a DEVICE array [n_sets][n_steps] of device pointers to 2^p-byte sketches
 * (NULL entries are skipped).  Set s, step i covers members[s][0..i].  This is the form the
 * drop-in sketch objects use: one launch evaluates the union of a tree node for every k
 * (n_sets = nk, final_only = 1), or every prefix of every ordering for every k.
 *   d_cards [n_sets][n_steps or 1] f64   d_hist [n_sets][n_steps or 1][64] u32
 *   d_unions NULL or [n_sets][n_steps or 1][2^p] u8 */
DD_API int dd_union_sets_card(const uint8_t *const *d_members, int n_sets, int n_steps, int p, int final_only,
                              double *d_cards, uint32_t *d_hist, uint8_t *d_unions, dd_stream stream);
DD_API int dd_union_max(const uint8_t *const *d_in, int n_in, size_t len, uint8_t *d_out, dd_stream stream);
/* Scratch: (d_ws, ws_bytes) sized by dd_prefix_union_workspace_bytes() lets the call use the bit-plane
 * kernel (it holds the transposed sketches and the identical-prefix table); with d_ws == NULL, or when
 * d_unions != NULL, the byte kernel runs, which needs no scratch. */
DD_API size_t dd_prefix_union_workspace_bytes(int n_ord, int n_steps, int n_genomes, int nk, int p);
DD_API int dd_prefix_union_card(const uint8_t *d_regs, const int32_t *d_order, int n_ord, int n_steps, int n_genomes,
                         int nk, int p, int final_only, double *d_cards, uint32_t *d_hist, uint8_t *d_unions,
                         void *d_ws, size_t ws_bytes, dd_stream stream);

/* Bit planes: a sketch of 2^p one-byte registers transposed into 6 planes of 2^p bits (plane b, word
 * g = bit b of registers 32g..32g+31; ranks are <= 64-p+1 < 64), 0.75 bytes per register.  A caller
 * that unions the same sketches many times (all pairs, many orderings) transposes them ONCE with
 * dd_to_planes and passes the planes to the *_planes entry points; p >= 12.
 *   d_planes [n_sketches][6][2^p / 32] u32,  dd_planes_bytes(n_sketches, p) bytes
 * dd_prefix_union_card_planes: d_planes laid out [n_genomes][nk]; scratch only for the
 * identical-prefix table, dd_prefix_union_workspace_bytes(n_ord, n_steps, 0, 0, p) bytes (NULL = none). */
DD_API size_t dd_planes_bytes(int64_t n_sketches, int p);
DD_API int dd_to_planes(const uint8_t *d_regs, int64_t n_sketches, int p, uint32_t *d_planes, dd_stream stream);
DD_API int dd_prefix_union_card_planes(const uint32_t *d_planes, const int32_t *d_order, int n_ord, int n_steps,
                                int n_genomes, int nk, int p, int final_only, double *d_cards, uint32_t *d_hist,
                                void *d_ws, size_t ws_bytes, dd_stream stream);

/* ===========================================================================================
 * K6  all-pairs union cardinalities (the shape of lib/huffman_dandd.py:666-695 and
 * helpers/allpairs.py:360-370): for each listed pair (a,b) and each of the nk sketches,
 * card(max(regs[a], regs[b])).
 *   d_pairs [n_pairs][2] int32      d_cards [n_pairs][nk] f64     d_hist [n_pairs][nk][64] u32
 * dd_pairwise_union_card transposes d_regs into (d_ws, ws_bytes) -- dd_prefix_union_workspace_bytes(0, 0,
 * n_genomes, nk, p) bytes -- on every call (NULL: byte kernel); dd_pairwise_union_card_planes takes
 * sketches already transposed by dd_to_planes, so tiles of a large pair matrix share one transpose.
 * =========================================================================================== */
DD_API int dd_pairwise_union_card(const uint8_t *d_regs, int n_genomes, int nk, int p, const int32_t *d_pairs,
                           int64_t n_pairs, double *d_cards, uint32_t *d_hist, void *d_ws, size_t ws_bytes,
                           dd_stream stream);
DD_API int dd_pairwise_union_card_planes(const uint32_t *d_planes, int n_genomes, int nk, int p, const int32_t *d_pairs,
                                  int64_t n_pairs, double *d_cards, uint32_t *d_hist, dd_stream stream);

/* ===========================================================================================
 * K5  --exact: number of distinct (canonical) k-mers, KMC semantics.
 * Replaces `kmc -ci1 -cs2 -kK [-b] -fm ...` + `kmc_tools info` (lib/sketch_classes.py:389-399,
 * 434-449) and, by inserting several streams into one set, `kmc_tools complex` unions (:451-465).
 * The set lives in the workspace: a 4^k-bit presence bitmap when k <= DD_EXACT_BITMAP_MAXK, else
 * an open-addressing table with `capacity` slots (power of two, must exceed the number of distinct
 * k-mers; dd_exact_count reports overflow as DD_ERR_WORKSPACE): 64-bit keys to k = 32, 128-bit keys to
 * k = 64, and for 64 < k <= 256 -- KMC's own limit; the README's "to sweep higher ks you must use KMC
 * via --exact" (README.md:82), helpers/allpairs.py:295 sweeps to 99 -- 16-byte (fingerprint, reference
 * to one occurrence) entries whose matches are verified against the packed stream, so the count stays
 * exact.  In that mode every stream inserted into a set must stay alive and unmodified until the last
 * dd_exact_insert into that set has completed (at most 512 distinct streams per set).
 * 1 <= k <= DD_EXACT_MAXK.
 * =========================================================================================== */
#define DD_EXACT_BITMAP_MAXK 16
#define DD_EXACT_MAXK 256
DD_API size_t dd_exact_workspace_bytes(int k, uint64_t capacity);
DD_API int dd_exact_begin(void *d_ws, size_t ws_bytes, int k, uint64_t capacity, dd_stream stream);
DD_API int dd_exact_insert(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end, int k,
                    int canon, void *d_ws, size_t ws_bytes, uint64_t capacity, dd_stream stream);
/* Multi-GPU exact mode ("per-range local distinct + all-reduce SUM"): the same insert
 * restricted to the k-mers of key range `shard_rank` of `shard_world` (a hash of the k-mer decides).
 * Every rank scans the same streams with its own rank; the per-rank dd_exact_count values add up to
 * the count a single dd_exact_insert pass would give, with 1/shard_world of the table per rank. */
DD_API int dd_exact_insert_shard(const uint32_t *d_codes, const uint32_t *d_invalid, uint64_t sym_begin, uint64_t sym_end,
                          int k, int canon, void *d_ws, size_t ws_bytes, uint64_t capacity, uint32_t shard_rank,
                          uint32_t shard_world, dd_stream stream);
/* Distinct k-mers inserted since dd_exact_begin -> *d_count (device u64); cumulative, so calling
 * it after each genome gives the progressive exact unions. */
DD_API int dd_exact_count(void *d_ws, size_t ws_bytes, int k, uint64_t capacity, uint64_t *d_count, dd_stream stream);

/* ===========================================================================================
 * Host-buffer convenience path (what bench.py's e2e number times): one FASTA held in host memory
 * -> H2D copy -> K1 -> K2 -> K4 -> D2H of cardinalities (and registers if h_regs != NULL);
 * synchronises the stream before returning.  Equivalent to the reference running
 * `dashing sketch` + `dashing card` for every k of the mask on that file.  FASTQ text (a line
 * beginning with '+') is detected by the packer, rewritten with dd_fastq_to_fasta_host and sketched
 * again, so the result is what kseq-based readers produce for it.
 * =========================================================================================== */
DD_API size_t dd_sketch_fasta_host_workspace_bytes(size_t n_bytes, int nk, int p);
DD_API int dd_sketch_fasta_host(const uint8_t *h_text, size_t n_bytes, uint32_t kmask, int p, int canon,
                         uint8_t *h_regs /*[nk][2^p] or NULL*/, double *h_cards /*[nk]*/, uint8_t *d_regs_or_null,
                         void *d_ws, size_t ws_bytes, dd_stream stream);

/* Same, but returns as soon as the work is enqueued: h_text must stay valid, and h_cards / h_regs
 * must not be read, until `stream` has been synchronised by the caller (pinned host memory makes
 * the copies truly asynchronous).  Lets a caller keep several FASTAs in flight on different streams
 * (one workspace per stream) so that H2D copies overlap the kernels of the previous file.  Nothing
 * in the call waits for the device, whatever the file size.  FASTQ text cannot be handled without
 * a host round trip: h_cards[] is then filled with NaN, and the caller falls back to
 * dd_sketch_fasta_host for that file. */
DD_API int dd_sketch_fasta_host_async(const uint8_t *h_text, size_t n_bytes, uint32_t kmask, int p, int canon,
                                      uint8_t *h_regs, double *h_cards, uint8_t *d_regs_or_null, void *d_ws,
                                      size_t ws_bytes, dd_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* DANDD_B200_H */
