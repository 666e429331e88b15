/*
 * evdb.h -- C ABI of libevdb_b200: a B200 (sm_100a) brute-force kNN engine that
 * drops in behind ErlVectorDB's search hot path.
 *
 * The reference (pure Erlang) has no FFI; the seam this ABI replaces is the
 * body of the vector_store gen_server (reference src/vector_store.erl:60-207):
 *
 *   init/1 bulk load            src/vector_store.erl:60-111  -> evdb_store_create + evdb_store_bulk_load_*
 *   handle_call({insert,..})    src/vector_store.erl:113-141 -> evdb_store_upsert_f64/_f32
 *   handle_call({search,..})    src/vector_store.erl:143-150,227-252 -> evdb_store_search_f64/_f32
 *   handle_call({delete,..})    src/vector_store.erl:152-164 -> evdb_store_delete
 *   handle_call(get_stats)      src/vector_store.erl:166-173 -> evdb_store_stats
 *   handle_call(get_all_vectors)src/vector_store.erl:184-190 -> evdb_store_get_f64
 *   terminate/2                 src/vector_store.erl:201-207 -> evdb_store_destroy
 *   vector_utils distances      src/vector_utils.erl:28-43   -> `metric` argument of search
 *   8/4-bit codecs              src/vector_compression.erl:166-204,306-329
 *                                                            -> evdb_quantize_{8,4}bit, evdb_dequantize_{8,4}bit,
 *                                                               evdb_store_bulk_load_codes
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; the caller owns every in/out
 *     buffer, the library owns device memory behind the opaque handle.
 *   - Every function returns EVDB_OK (0) or a negative EVDB_E_* code; nothing
 *     throws across the boundary.  EVDB_E_DIM_MISMATCH / EVDB_E_BAD_VECTOR map
 *     to the reference atoms dimension_mismatch / invalid_vector_format.
 *   - A handle is used by one caller at a time (the store's gen_server
 *     guarantees it); distinct handles are fully concurrent.
 *   - There is NO CPU fallback: without a usable sm_100 device every entry
 *     point that touches a store fails with EVDB_E_NO_DEVICE.
 *   - Rows live in dense slots [0, count).  Id <-> slot and metadata stay with
 *     the caller (Erlang state).  Results order by (distance, slot); the caller
 *     re-applies the reference's {Distance, Id} term order among exact ties.
 */
#ifndef EVDB_H
#define EVDB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVDB_ABI_VERSION 2
#define EVDB_MAX_SHARDS 16

typedef struct evdb_store evdb_store;

/* element type of the device-resident columns */
enum { EVDB_F32 = 0, EVDB_BF16 = 1, EVDB_U8 = 2, EVDB_U4 = 3 };
/* distance (src/vector_store.erl:238-246, src/vector_utils.erl:38-43) */
enum { EVDB_COSINE = 0, EVDB_EUCLIDEAN = 1, EVDB_MANHATTAN = 2 };
/* search plan override (evdb_store_set_plan); AUTO picks by batch size.  GEMM = tensor cores: tcgen05 kind::f16
 * over the fp16 operand column of an F32 store (cosine, euclidean), kind::i8 over the codes of a U8 store
 * (cosine); EVDB_E_UNSUPPORTED for other dtype / metric pairs when forced. */
enum { EVDB_PLAN_AUTO = 0, EVDB_PLAN_SCAN = 1, EVDB_PLAN_GEMM = 2, EVDB_PLAN_EXACT = 3 };

enum {
    EVDB_OK = 0,
    EVDB_E_DIM_MISMATCH = -1, /* {error, dimension_mismatch}     */
    EVDB_E_BAD_VECTOR = -2,   /* {error, invalid_vector_format}  */
    EVDB_E_OOM = -3,
    EVDB_E_CUDA = -4,
    EVDB_E_NCCL = -5,
    EVDB_E_BAD_ARG = -6,
    EVDB_E_NO_DEVICE = -7,
    EVDB_E_UNSUPPORTED = -8,
    EVDB_E_BADARITH = -9 /* Max == Min in a quantizer (reference: badarith); reserved: see upsert */
};

typedef struct evdb_opts {
    int32_t device;         /* CUDA device ordinal (n_shards <= 1) */
    int32_t dtype;          /* EVDB_F32 | EVDB_BF16 | EVDB_U8 | EVDB_U4 */
    int32_t dim;            /* 0 = fixed by the first upsert/bulk load (reference :213-217) */
    int32_t gemm_shadow;    /* F32 stores: keep the fp16 operand column for the tcgen05 path */
    uint64_t capacity_hint; /* rows to reserve up front (0 = grow by doubling) */
    /* ONE store behind ONE handle on n_shards GPUs of this process (the BEAM is one OS process with
     * one gen_server per store, reference src/vector_store.erl:38-39): slot g lives on shard
     * g % n_shards at local slot g / n_shards, so appends and swap-with-last deletes keep every shard
     * dense and balanced.  Searches fan out on one host thread + stream per device, candidates meet
     * over peer memory (NVLink), the merged result comes back through the same host entry points.
     * 0 or 1 = a single-device store on `device`.  The same ordinal may be listed more than once
     * (several shards on one GPU: tests on a one-GPU box).                                        */
    int32_t n_shards;
    int32_t devices[EVDB_MAX_SHARDS];
} evdb_opts;

typedef struct evdb_stats {
    uint64_t count;          /* maps:size(Vectors)                       */
    int32_t dimension;       /* 0 == undefined                           */
    int32_t dtype;
    int32_t device;
    int32_t last_plan;       /* EVDB_PLAN_* actually used by the last search */
    uint64_t capacity;
    uint64_t device_bytes;   /* HBM held by this store                   */
    uint64_t searches;       /* queries answered                         */
    uint64_t rows_scanned;   /* sum over queries of rows visited         */
    uint64_t escalations;    /* candidate windows that had to be widened */
    uint64_t kernel_launches;/* CUDA kernels launched by this store      */
    double last_search_ms;   /* device time of the last search call      */
    /* ---- ABI 2 ---- */
    int32_t n_shards;        /* devices behind this handle (1 = single)  */
    int32_t gemm_disabled;   /* AUTO gave up the tcgen05 plan for this store: its operand column or
                                candidate buffers did not fit in device memory (scan plan from then on) */
    uint64_t shadow_bytes;   /* HBM held by the fp16 operand columns      */
    uint64_t upserts;        /* rows written by upsert/append/bulk load   */
    uint64_t deletes;
    double last_h2d_ms;      /* last host search: query copy, device work, result copy */
    double last_device_ms;
    double last_d2h_ms;
} evdb_stats;

/* ---- process / device ---------------------------------------------------- */
int evdb_abi_version(void);
/* Probe the devices a deployment intends to use; EVDB_E_NO_DEVICE unless each
 * is a compute-capability 10.x GPU.  devices == NULL -> device 0.           */
int evdb_init(const int *devices, int n_dev);
const char *evdb_strerror(int code);
/* last CUDA error text seen by the calling thread ("" if none) */
const char *evdb_last_cuda_error(void);

/* ---- store lifecycle (vector_store_sup:start_store / stop_store) --------- */
int evdb_store_create(const evdb_opts *opts, evdb_store **out);
void evdb_store_destroy(evdb_store *s);
int evdb_store_stats(evdb_store *s, evdb_stats *out);
int evdb_store_set_plan(evdb_store *s, int plan);
/* Measurement aid (bench.py roofline): when enabled, every search brackets its dominant
 * kernel (the scan or the tcgen05 GEMM) with a CUDA event pair on the launching stream.
 * _read synchronises that stream, returns the samples taken since the last read and their
 * summed duration, and resets.  At most 512 samples are kept between reads.           */
int evdb_store_profile(evdb_store *s, int enable);
int evdb_store_profile_read(evdb_store *s, int32_t *n_samples, double *total_ms);

/* ---- ingest --------------------------------------------------------------
 * upsert: slot == count appends, slot < count overwrites (maps:put upsert).
 * First vector fixes the dimension; d != dimension -> EVDB_E_DIM_MISMATCH.
 * Non-finite elements -> EVDB_E_BAD_VECTOR (Erlang floats are always finite).
 * For U8/U4 stores the row is quantized on the device exactly as
 * compress_{8,4}bit_quantization does (fp64).  A row with Max == Min (the reference raises badarith and
 * keeps the RAW constant vector, src/vector_compression.erl:62-64, src/vector_persistence.erl:114-116)
 * is stored as {min, scale = 0} with all-zero codes, which decodes to exactly that vector: the insert
 * succeeds, as it does in the reference.  (EVDB_E_BADARITH is reported by the standalone codecs through
 * their ok[] output, not by a store.)
 * The dimension is fixed only by a vector that passed every check (:128-131): a rejected first vector
 * leaves it undefined.                                                          */
int evdb_store_upsert_f64(evdb_store *s, uint32_t slot, const double *vec, int d);
int evdb_store_upsert_f32(evdb_store *s, uint32_t slot, const float *vec, int d);
/* Batched insert of NEW ids (a run of handle_call({insert,..}) with fresh keys, or a reload in
 * pieces): n rows appended at slots [count, count + n), *first_slot = the first of them.  One
 * transfer and one finalize pass instead of n; same validation and error codes as upsert.      */
int evdb_store_append_f64(evdb_store *s, const double *rows, uint64_t n, int d, uint64_t *first_slot);
int evdb_store_append_f32(evdb_store *s, const float *rows, uint64_t n, int d, uint64_t *first_slot);
/* upsert / append / delete only ENQUEUE their device work (the rows are staged through pinned memory, so
 * the caller's buffer is free on return; searches are ordered behind them).  flush waits for everything
 * enqueued so far and reports a device error that surfaced since -- terminate/2, sync/1 and tests call it. */
int evdb_store_flush(evdb_store *s);
/* Replace the whole content with n rows (vector_store:init/1 bulk load). */
int evdb_store_bulk_load_f32(evdb_store *s, const float *rows, uint64_t n, int d);
int evdb_store_bulk_load_f64(evdb_store *s, const double *rows, uint64_t n, int d);
/* Compressed records straight to device columns, no decompress-to-list
 * (vector_persistence:load_vectors + decompress_if_needed, :157-165,:276-284).
 * codes: n rows of d bytes (U8) or (d+1)/2 bytes (U4, first element in the
 * high nibble); mins/scales: the records' fp64 metadata.                    */
int evdb_store_bulk_load_codes(evdb_store *s, const uint8_t *codes, const double *mins,
                               const double *scales, uint64_t n, int d);
/* Swap-with-last delete.  *moved_from = slot that now lives in `slot`, or -1. */
int evdb_store_delete(evdb_store *s, uint32_t slot, int64_t *moved_from);
/* Read a row back as the reference would see it (fp64; quantized stores give
 * Min + Q*Scale).                                                           */
int evdb_store_get_f64(evdb_store *s, uint32_t slot, double *out, int d);
/* Raw codes + metadata of a quantized row (bit-exact codec checks). */
int evdb_store_get_codes(evdb_store *s, uint32_t slot, uint8_t *codes, double *mn,
                         double *scale);
/* Bench/test only: fill slots [0,n) with the counter-based synthetic corpus
 *  generated on the device; rows are global rows
 * [row0, row0+n) of that corpus (row0 > 0 for a shard).                     */
int evdb_store_fill_synthetic(evdb_store *s, uint64_t seed, uint64_t row0, uint64_t n, int d);

/* ---- search ---------------------------------------------------------------
 * B queries of d numbers each.  For each query b: out_counts[b] = min(k, count)
 * results in out_slots[b*k ..], out_dists[b*k ..], ascending by
 * (distance, slot).  Distances are fp64 and, for F32/U8/U4 stores, bit-equal
 * to the reference's Erlang arithmetic on the stored rows (same operation
 * order, no FMA).  k == 0 -> counts 0.  k < 0 -> EVDB_E_BAD_ARG (the reference
 * crashes with function_clause).  Empty store -> counts 0 for any d.        */
int evdb_store_search_f64(evdb_store *s, const double *queries, int B, int d, int k,
                          int metric, uint32_t *out_slots, double *out_dists,
                          int32_t *out_counts);
int evdb_store_search_f32(evdb_store *s, const float *queries, int B, int d, int k,
                          int metric, uint32_t *out_slots, double *out_dists,
                          int32_t *out_counts);

/* Device-resident variant: d_queries (B x d fp64), d_out_slots (B x k u32),
 * d_out_dists (B x k fp64), d_out_counts (B i32) are DEVICE pointers on the
 * store's device; work is enqueued on `stream` (a cudaStream_t; NULL = the
 * store's own stream; pass cudaStreamLegacy for the default stream) and NOT synchronised.  slot_base is added to every
 * returned slot (global row ids for a row-sharded corpus).  Escalation of the
 * candidate window needs a host decision, so this variant reports a query
 * whose window could not be proven complete in d_out_flags[b] != 0 (may be
 * NULL) instead of retrying; callers re-issue those through the host path.  */
int evdb_store_search_dev(evdb_store *s, const void *d_queries_f64, int B, int d, int k,
                          int metric, uint64_t slot_base, void *d_out_ids_u64,
                          void *d_out_dists_f64, void *d_out_counts_i32,
                          void *d_out_flags_i32, void *stream);

/* The same with the knobs an escalation needs (a sharded caller re-issues flagged queries on every
 * rank: first with a wider window on the scan plan, then through the exhaustive fp64 plan -- the
 * ladder evdb_store_search_f64 climbs by itself, src of the rule: DESIGN.md section 4).          */
typedef struct evdb_search_opts {
    int32_t plan;         /* EVDB_PLAN_*; AUTO = the store's own plan                              */
    int32_t kp_min;       /* smallest candidate window to use (0 = default)                        */
    uint64_t slot_base;   /* returned id = slot_base + slot * slot_stride                          */
    uint64_t slot_stride; /* 0 or 1 = contiguous block; S = this store holds every S-th global row */
} evdb_search_opts;
int evdb_store_search_dev_ex(evdb_store *s, const void *d_queries_f64, int B, int d, int k, int metric,
                             const evdb_search_opts *o, void *d_out_ids_u64, void *d_out_dists_f64,
                             void *d_out_counts_i32, void *d_out_flags_i32, void *stream);

/* G-way merge of per-shard results (the step after the NCCL allgather):
 * in: G lists per query, laid out [G][B][k] (ids u64, dists fp64, counts
 * [G][B]); out: [B][k] ascending by (distance, id).  Device pointers.        */
int evdb_merge_topk_dev(int device, const void *d_ids_u64, const void *d_dists_f64,
                        const void *d_counts_i32, int G, int B, int k, void *d_out_ids_u64,
                        void *d_out_dists_f64, void *d_out_counts_i32, void *stream);

/* Same merge over PACKED per-rank results, the layout of the one-collective exchange: each
 * rank's blob is 2*B*k + B 64-bit words = [B*k ids u64][B*k dists fp64][B counts i32][B flags i32];
 * d_blobs holds G of them back to back (the allgather output), d_out_blob receives one blob
 * (flags OR-ed over the ranks).  evdb_store_search_dev can write straight into a blob.        */
int evdb_merge_topk_packed_dev(int device, const void *d_blobs, int G, int B, int k,
                               void *d_out_blob, void *stream);

/* ---- peer-memory exchange for a row-sharded store (one rank per GPU/process) -------------
 * Instead of an allgather + merge launch: every rank pushes its packed blob into a mailbox in
 * every peer's memory (NVLink stores; peers mapped with CUDA IPC) and publishes an epoch flag;
 * the merge kernel waits for the flags on the device.  create -> exchange the 64-byte handles
 * among the ranks (any transport) -> connect -> per search: push, merge (both only enqueue).   */
typedef struct evdb_exchange evdb_exchange;
int evdb_exchange_create(int device, int rank, int world, uint64_t max_words, evdb_exchange **out,
                         void *ipc_handle_out /* 64 bytes */);
int evdb_exchange_connect(evdb_exchange *x, const void *all_handles /* world * 64 bytes */);
int evdb_exchange_connect_ptrs(evdb_exchange *x, const void *const *mailboxes /* same process */);
void *evdb_exchange_mailbox(evdb_exchange *x);
int evdb_exchange_push(evdb_exchange *x, const void *d_local_blob, int B, int k, void *stream);
int evdb_exchange_merge(evdb_exchange *x, int B, int k, void *d_out_blob, void *stream);
void evdb_exchange_destroy(evdb_exchange *x);

/* ---- row-sharded query batches in two phases (tcgen05 GEMM plan, F32 stores) -----------------
 * Re-ranking every shard's local top-k repeats the exact fp64 work on every rank.  Instead:
 *   phase1: local GEMM + window selection; the approximate window (B x KP keys with global rows,
 *           size, error bound) is pushed to every rank through exchange `xw`;
 *   phase2: every rank merges the world windows into the same global window and re-ranks in exact
 *           fp64 only the candidates it owns; those distances are pushed through exchange `xe`;
 *   phase3: every rank orders the global candidates by the owners' exact distances, emits the
 *           top k into d_out_blob (packed layout of evdb_merge_topk_packed_dev) and proves the window.
 * All three only enqueue.  rank/world come from the exchanges (mailboxes of at least B*KP + B resp.
 * B*KP words, KP <= 128).  EVDB_E_UNSUPPORTED (before anything is enqueued) when the plan does not
 * apply -- callers decide from global facts (dtype, metric, B, k, smallest shard) so that every
 * rank takes the same path.  n_total = rows of the whole store, slot_base = first global row here. */
int evdb_store_search_sharded_phase1(evdb_store *s, evdb_exchange *xw, const void *d_queries_f64, int B,
                                     int d, int k, int metric, uint64_t slot_base, uint64_t n_total,
                                     void *stream);
int evdb_store_search_sharded_phase2(evdb_store *s, evdb_exchange *xw, evdb_exchange *xe,
                                     const void *d_queries_f64, int B, int k, int metric,
                                     uint64_t n_total, void *stream);
int evdb_store_search_sharded_phase3(evdb_store *s, evdb_exchange *xe, int B, int k, int metric,
                                     uint64_t n_total, void *d_out_blob, void *stream);

/* ---- codecs (vector_compression.erl:166-204), computed on the device ------
 * n rows of d fp64 -> codes (+ per-row fp64 min/max/scale).  ok[i] = 0 for a
 * row whose Max == Min (reference: badarith, caller stores it raw).          */
int evdb_quantize_8bit(int device, const double *rows, uint64_t n, int d, uint8_t *codes,
                       double *mins, double *maxs, double *scales, uint8_t *ok);
int evdb_quantize_4bit(int device, const double *rows, uint64_t n, int d, uint8_t *packed,
                       double *mins, double *maxs, double *scales, uint8_t *ok);
int evdb_dequantize_8bit(int device, const uint8_t *codes, const double *mins,
                         const double *scales, uint64_t n, int d, double *out);
int evdb_dequantize_4bit(int device, const uint8_t *packed, const double *mins,
                         const double *scales, uint64_t n, int d, double *out);

/* ---- vector_utils (reference src/vector_utils.erl:28-57), computed on the device ----------------
 * n pairs (a[i], b[i]) of d fp64 numbers -> out[i], in the reference's operation order (independent
 * products, left-to-right lists:sum, no FMA): bit-equal to the Erlang result.  cosine_similarity/2 is
 * Dot/(|a||b|) and 0.0 on a zero norm (:28-36) -- similarity, NOT the distance search uses;
 * EVDB_VU_NORM ignores b (may be NULL).                                                          */
enum { EVDB_VU_COSINE_SIMILARITY = 0, EVDB_VU_COSINE_DISTANCE = 1, EVDB_VU_EUCLIDEAN = 2, EVDB_VU_MANHATTAN = 3,
       EVDB_VU_DOT = 4, EVDB_VU_NORM = 5 };
int evdb_vector_utils_f64(int device, int op, const double *a, const double *b, uint64_t n, int d, double *out);

/* ---- diagnostics -----------------------------------------------------------
 * Host arithmetic only (touches no device): the tile plan the TMA-staged scan over
 * quantization_8bit / _4bit codes (csrc/scan.cu, replaces the maps:fold of
 * src/vector_store.erl:227-231 for compressed stores) would use for a store of `count` rows of
 * `dim` elements with a `window`-key candidate window on a GPU with `sm_count` SMs.
 * Returns 1 and out[8] = {lanes per row, warps per tile, ring stages, bytes per stage, rows per tile,
 * chunk rotation, dynamic shared memory bytes, consumer groups}; 0 when the register-fed scan
 * would run instead; a negative EVDB_E_* for bad arguments.                              */
int evdb_debug_scan_tile_plan(int dtype, int dim, int window, uint64_t count, int sm_count, int32_t *out);
/* The INTEGER part of the quantized scan (replaces dot_product/2 over decompressed lists, reference
 * src/vector_store.erl:248-249 after src/vector_compression.erl:180-183,201-204), observable: for
 * each of n slots of a U8/U4 store, out_sum = sum_i Q_i * c_i where c are the row's codes and
 * Q_i = clamp(rint(query_i * 2^shift)) is the query on the scan's 16-bit fixed-point grid
 * (*out_shift); out_planes[3*i..] = the signed high-digit and unsigned low-digit dp4a sums the
 * kernels accumulate (third = 0 with two planes), out_code_sum = sum_i c_i.  Exact integers: a test
 * compares them with int64 arithmetic on the host (north_star: "integer quantized-code dot
 * products are bit-exact").                                                                  */
int evdb_debug_quant_dots(evdb_store *s, const double *query, int d, const uint32_t *slots, int n,
                          int64_t *out_sum, int32_t *out_planes, int32_t *out_code_sum, int32_t *out_shift);

#ifdef __cplusplus
}
#endif
#endif /* EVDB_H */
