// internal.h -- store layout and kernel launchers shared by the .cu files.
#pragma once
#include "common.cuh"

namespace evdb {

// per-query scalars produced by prep_queries and consumed by the scans
struct QStat {
    float inv_norm;  // 1/||q|| (0 when ||q|| == 0)
    float sum;       // sum(q)   (quantized scans: of the fixed-point query)
    float fx;        // 2^-e: value of one fixed-point unit (quantized scans)
    float norm_sq;   // ||q||^2
};

// Quantized scans: the query sits on a fixed-point grid of 8*kQPlanes bits, one 8-bit digit plane per
// dp4a.  Two planes: 1/3 less integer work per code; the coarser grid is covered by a per-query
// error bound (scan.cu prep_queries_kernel), never by the result (exact fp64 re-rank).
#ifndef EVDB_QPLANES
#define EVDB_QPLANES 2
#endif
constexpr int kQPlanes = EVDB_QPLANES;
static_assert(kQPlanes == 2 || kQPlanes == 3, "digit planes");
constexpr int kMaxKP = 1024;   // widest candidate window the scan/select path carries
constexpr int kMinKP = 16;
constexpr int kScanWarps = 8;  // warps per scan CTA
constexpr int kProfMax = 512;

}  // namespace evdb

namespace evdb { struct MStore; }

namespace evdb {
// what a captured search graph was built for: shape, store state, and the buffers baked into its nodes
struct GraphKey {
    int B = -1, k = 0, metric = 0, f64 = 0;
    uint64_t count = 0, epoch = 0;
    const void *q = nullptr, *out = nullptr, *pin = nullptr;
    uint64_t ws = 0;     // signature of every workspace / column pointer the plan's kernels were given
    bool operator==(const GraphKey &o) const {
        return B == o.B && k == o.k && metric == o.metric && f64 == o.f64 && count == o.count && epoch == o.epoch &&
               q == o.q && out == o.out && pin == o.pin && ws == o.ws;
    }
};
}  // namespace evdb

// Device-resident store.  One owner thread at a time (the store's gen_server).
struct evdb_store {
    evdb::MStore *multi = nullptr;   // n_shards > 1: this handle only fronts the shard stores (mstore.cu)
    int device = 0;
    int dtype = EVDB_F32;
    int dim = 0;    // 0 = undefined (reference: dimension :: undefined)
    int dpad = 0;   // elements per row incl. zero padding (row = nch 16-byte chunks)
    int nch = 0;    // 16-byte chunks per row of the main column
    size_t row_bytes = 0;
    int sm_count = 148;
    int plan = EVDB_PLAN_AUTO;
    int gemm_shadow = 0;
    uint64_t count = 0, capacity = 0, capacity_hint = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;   // host search: start, end, after H2D, before D2H

    // ---- columns (all [capacity]) ----
    uint8_t *rows = nullptr;       // F32: f32[dpad]; BF16: bf16[dpad]; U8: u8[dpad]; U4: u8[dpad/2]
    double *norm64 = nullptr;      // exact reference vector_norm of the stored row
    float *inv_norm = nullptr;     // (float)(1/norm64), 0 for a zero row
    float *norm_sq = nullptr;      // (float)(norm64^2)   (GEMM euclidean)
    float2 *qcoef = nullptr;       // U8/U4: {scale/||y||, min/||y||}
    double2 *qms64 = nullptr;      // U8/U4: {min, scale} fp64 (exact re-rank, read-back)
    __half *shadow = nullptr;      // F32 + gemm_shadow: fp16 of v/||v||, [capacity][spitch] (tcgen05 path)
    int spitch = 0;                // shadow row pitch in elements (dim rounded up to 64: 128-byte aligned rows)
    uint64_t shadow_valid = 0;     // rows [0, shadow_valid) of the shadow are current
    __half *shadow_l2 = nullptr;   // euclidean tcgen05 operand: fp16(sigma*v) [capacity][spitch], built on first use
    __half *l2_tail = nullptr;     // its norm K-step: [capacity][16] = {h1, h2, h3, 0...}, h1+h2+h3 = -sigma^2*||v||^2/2
    int l2_pitch = 0;              // row pitch of shadow_l2 in elements (== spitch)
    uint64_t l2_cap = 0, l2_valid = 0;  // rows allocated / rows [0, l2_valid) current
    float l2_sigma = 0.f;          // power-of-two scale: sigma * max||v|| in [64, 128)
    double max_norm = 0.0;         // upper bound on the largest row norm (valid when !max_norm_dirty)
    int max_norm_dirty = 1;
    uint64_t slot_mul = 1;        // returned id = slot_base + slot * slot_mul (> 1: round-robin shard of a multi-device store)
    unsigned int *d_arrive = nullptr;   // [8] arrival counters of the fused small-store kernel (zero between launches)
    int gemm_oom = 0;            // the GEMM plan ran out of device memory once: AUTO stays on the scan plan
    void *d_scalar = nullptr;      // 64-byte device scratch (reductions)

    // ---- workspace (grown on demand) ----
    double *w_q64 = nullptr;   size_t w_q64_cap = 0;    // [B][dim]
    float *w_q32 = nullptr;    size_t w_q32_cap = 0;    // [B][dpad32]
    uint8_t *w_qdig = nullptr; size_t w_qdig_cap = 0;   // [B][3][dpad] digit planes
    void *w_seed = nullptr;    size_t w_seed_cap = 0;   // [Bpad][S] sampled scores + [Bpad] seeded thresholds (GEMM)
    void *w_qh = nullptr;      size_t w_qh_cap = 0;     // [Bpad][spitch] fp16 unit-norm queries (GEMM)
    evdb::QStat *w_qstat = nullptr; size_t w_qstat_cap = 0;
    float *w_qeps = nullptr; size_t w_qeps_cap = 0;   // quantized scans: per-query bound of the query grid error
    uint64_t *w_partial = nullptr; size_t w_partial_cap = 0; // [B][G][KP]
    uint64_t *w_ids = nullptr;  size_t w_ids_cap = 0;   // [B][k]
    double *w_dists = nullptr;  size_t w_dists_cap = 0; // [B][k]
    int32_t *w_counts = nullptr; size_t w_counts_cap = 0; // [B] counts + [B] flags
    void *w_tmp = nullptr;      size_t w_tmp_cap = 0;   // ingest staging / exact plan
    void *w_shard = nullptr;    size_t w_shard_cap = 0; // sharded two-phase search: window blob, exact blob, global window, meta
    void *h_pin = nullptr;      size_t h_pin_cap = 0;   // pinned host staging
    // ingest at rate: pinned + device staging ring (two halves), enqueue-only upserts / deletes
    uint8_t *h_ring = nullptr, *d_ring = nullptr;
    cudaEvent_t ring_ev[2] = {nullptr, nullptr}, ev_ing = nullptr;
    int ring_busy[2] = {0, 0};
    uint64_t ring_pos = 0;
    size_t ring_off = 0;        // bytes of the current half already handed to the stream
    int ingest_pending = 0;     // work enqueued on `stream` since the last flush
    // small host searches replayed as one CUDA graph (store.cu search_host)
    evdb::GraphKey gkey, last_key;
    cudaGraphExec_t gexec = nullptr;
    uint64_t graph_epoch = 0, graph_launches = 0;
    int graph_plan = 0, graph_broken = 0;
    void *h_gq = nullptr; size_t h_gq_cap = 0;   // pinned slot the graph copies its queries from

    // ---- dominant-kernel profiling (evdb_store_profile) ----
    int prof_on = 0, prof_n = 0;
    cudaEvent_t *prof_ev = nullptr;  // [2 * kProfMax]
    cudaStream_t prof_stream = nullptr;

    // ---- counters ----
    uint64_t n_searches = 0, n_rows_scanned = 0, n_escalations = 0, n_launches = 0, n_upserts = 0, n_deletes = 0;
    double last_h2d_ms = 0.0, last_device_ms = 0.0, last_d2h_ms = 0.0;
    int last_plan = 0;
    double last_search_ms = 0.0;
};

// Peer-memory mailbox of one rank (exchange.cu): [2 parities][world slots][slot_words] u64, then
// [2][world] epoch flags, then a block counter.
struct evdb_exchange {
    int device = 0, rank = 0, world = 1;
    uint64_t max_words = 0;              // blob capacity of a mailbox slot (u64 words)
    uint64_t slot_words = 0;             // slot pitch (max_words rounded up to 32 words)
    unsigned long long epoch = 0;
    uint64_t *mailbox = nullptr;
    uint64_t **d_peer_box = nullptr;     // device array [world]: base of each rank's mailbox (mine included)
    void *peer_base[64] = {nullptr};     // IPC-opened pointers (to close)
    bool opened[64] = {false};
};

namespace evdb {

// what a consumer kernel of the CURRENT epoch needs: the world slots, their pitch, the flags to wait on
struct ExchangeView {
    const uint64_t *slots;
    size_t stride;
    const unsigned long long *flags;
    unsigned long long epoch;
};
int exchange_push_words(evdb_exchange *x, const void *d_blob, size_t words, cudaStream_t st);
// Fused push: the PRODUCING kernel stores its output words straight into every peer's mailbox slot
// (push_store) and its last warp publishes the epoch flags (push_arrive) -- no separate copy kernel.
struct PushTarget {
    uint64_t *const *peer_box;        // device array [world]
    unsigned long long slot_off;      // word offset of my slot of this epoch's parity in every mailbox
    unsigned long long flag_off;      // word offset of my flag there
    unsigned long long epoch;
    unsigned int *counter;            // local arrival counter (zero between uses)
    int world;
};
PushTarget exchange_begin_push(evdb_exchange *x);   // starts the next epoch
ExchangeView exchange_view(const evdb_exchange *x);

struct ScanArgs {
    const uint8_t *rows;
    size_t row_bytes;
    int nch;                 // 16-byte chunks per row
    uint64_t n;
    const float *inv_norm;   // F32/BF16 cosine
    const float2 *qcoef;     // U8/U4 cosine
    const float *q32;        // [B][q32_stride]
    int q32_stride;
    const uint8_t *qdig;     // [B][3][qdig_stride]
    int qdig_stride;
    const QStat *qstat;      // [B]
    uint64_t *partial;       // [B][G][KP]
    int KP;
    int G;
    int B;
    int tma_stages = 0;      // TMA-staged quantized scan: ring depth, bytes per stage, chunk rotation
    int tma_stage_bytes = 0;
    int bank_mul = 0;
    int tma_wt = 8;          // consumer warps per tile
};

// Unsorted candidate buffers left by the tcgen05 GEMM plan: query q = (sweep c, block mb, thread et)
// has one buffer per (row group ng, column part): set l = (c*nCTA + ng*MB + mb)*parts + part, its i-th
// key at cand[(l*cap + i)*gm + et], fill count at cnt[l*gm + et].
struct RawCands {
    const uint64_t *cand;
    const int *cnt;
    int cap, nCTA, MB, NG, parts, gm;
};
constexpr int kRawMaxLists = 640;  // NG * parts <= 148 * 4

// ---- launchers (each returns EVDB_OK or an error; all async on `st`) ----
int launch_prep_queries(evdb_store *s, const double *d_q64, int B, int metric, cudaStream_t st);
typedef void (*scan_fn_t)(const ScanArgs);
scan_fn_t pick_float_mq_f32(int metric, int tpr, int Q);    // scan_mq_f32.cu
scan_fn_t pick_float_mq_bf16(int metric, int tpr, int Q);   // scan_mq_bf16.cu
int scan_grid_size(evdb_store *s, int metric, int KP, int B, int *G_out);
int debug_quant_dots(evdb_store *s, const double *d_q64, const uint32_t *d_slots, int n, long long *d_S, int *d_planes,
                     int *d_csum, float *h_fx, cudaStream_t st);
int launch_scan(evdb_store *s, int metric, const ScanArgs &a, cudaStream_t st);
// gemm_i8.cu: query batches against a U8 store as a tcgen05 kind::i8 GEMM over the codes
bool qgemm_plan_supported(evdb_store *s, int metric, int B, int KP);
int launch_seed_thresholds(const float *dump, int pooled, int Bpad, int KP, uint32_t *thr, cudaStream_t st);   // gemm_tcgen05.cu
int launch_qgemm_topk(evdb_store *s, const double *d_q64, int B, int KP, int *lists_per_query,
                      const float **d_eps_q, RawCands *raw, cudaStream_t st);
// eps_q: optional per-query absolute bound added to eps_abs; squared: key scores are squared
// distances (euclidean GEMM plan)
int launch_select(evdb_store *s, const double *d_q64, const uint64_t *partial, const RawCands *raw, int lists_per_query,
                  int KP, int B, int kk, int kstride, int metric, float eps_abs, float eps_rel,
                  const float *eps_q, int squared, uint64_t slot_base, uint64_t *d_out_ids, double *d_out_dists,
                  int32_t *d_out_counts, int32_t *d_out_flags, cudaStream_t st);
// d_src != NULL (float stores): rows [slot0, slot0+n) are first narrowed from the staged n x dim source (fused ingest)
int launch_small_fused(evdb_store *s, const double *d_q64, int B, int KP, int kk, int kstride, int metric, int G, int tpr,
                       uint64_t *partial, float *eps_q, unsigned int *arrive, float eps_abs, float eps_rel,
                       uint64_t slot_base, uint64_t *d_out_ids, double *d_out_dists, int32_t *d_out_counts,
                       int32_t *d_out_flags, cudaStream_t st);
int scan_small_plan(evdb_store *s, int metric, int KP, int *G_out, int *tpr_out);   // grid / lane group the one-query float scan uses
int launch_finalize_rows(evdb_store *s, uint64_t slot0, uint64_t n, cudaStream_t st, const void *d_src = nullptr, bool src_f64 = false);
int launch_fill_synthetic(evdb_store *s, uint64_t seed, uint64_t row0, uint64_t rstride, uint64_t n, cudaStream_t st);
int launch_quantize_rows(int dtype, const double *d_rows64, const float *d_rows32, uint64_t n,
                         int d, uint8_t *codes, size_t code_row_bytes, double2 *ms64,
                         double *maxs, uint8_t *ok, cudaStream_t st);
int launch_dequantize_rows(int dtype, const uint8_t *codes, size_t code_row_bytes,
                           const double2 *ms64, uint64_t n, int d, double *out, cudaStream_t st);
int launch_vector_utils(int op, const double *d_a, const double *d_b, uint64_t n, int d, double *d_out, cudaStream_t st);
int exact_plan_search(evdb_store *s, const double *d_q64, int B, int kk, int kstride, int metric,
                      uint64_t slot_base, uint64_t *d_out_ids, double *d_out_dists,
                      int32_t *d_out_counts, int32_t *d_out_flags, cudaStream_t st);
int launch_merge_topk(const uint64_t *ids, const double *dists, const int32_t *counts, int G, int B,
                      int k, uint64_t *out_ids, double *out_dists, int32_t *out_counts,
                      cudaStream_t st);
int launch_merge_topk_packed(const uint64_t *blobs, size_t blob_stride, int G, int B, int k, uint64_t *out_blob,
                             cudaStream_t st, const unsigned long long *arrived = nullptr,
                             unsigned long long epoch = 0);
int check_device_public(int dev);
// row-sharded two-phase GEMM search (select.cu)
int launch_shard_window(evdb_store *s, const double *d_q64, const RawCands *raw, int L, int KP, int B, int kk,
                        int metric, const float *eps_q, uint64_t slot_base, const PushTarget &push, cudaStream_t st);
size_t shard_gmeta_bytes(int B);
int launch_shard_rerank(evdb_store *s, const double *d_q64, int B, int KP, int k, int kk, int metric, int rank,
                        int world, uint64_t n_total, const ExchangeView &win, const PushTarget &e_push, uint64_t *g_out,
                        void *g_meta, cudaStream_t st);
int launch_shard_final(evdb_store *s, int B, int KP, int k, int kk, int metric, int rank, int world, uint64_t n_total,
                       const ExchangeView &ex, const uint64_t *g_out, const void *g_meta, uint64_t *out_blob,
                       cudaStream_t st);
// tcgen05 path (gemm_tcgen05.cu)
bool gemm_plan_supported(evdb_store *s, int metric, int B, int KP);
int gemm_kp(int KP);
int gemm_max_batch();
int launch_gemm_topk(evdb_store *s, const double *d_q64, int B, int KP, int metric, int *lists_per_query,
                     const float **d_eps_q, RawCands *raw, cudaStream_t st);
int launch_l2_shadow_rows(evdb_store *s, uint64_t slot0, uint64_t n, cudaStream_t st);

// store.cu internals shared with the multi-device store (mstore.cu)
int choose_kp(int kk, int kp_min);
int search_core(evdb_store *s, const double *d_q64, int B, int k, int kstride, int metric, int kp_min, int plan,
                uint64_t slot_base, uint64_t *d_ids, double *d_dists, int32_t *d_counts, int32_t *d_flags, cudaStream_t st);
int store_put_rows(evdb_store *s, uint64_t slot0, const void *rows, bool is_f64, uint64_t n, size_t pitch, int d);
int store_load_codes(evdb_store *s, const uint8_t *codes, size_t code_pitch, const double *mins, const double *scales,
                     size_t ms_stride, uint64_t n, int d);
int store_fill_synthetic(evdb_store *s, uint64_t seed, uint64_t row0, uint64_t row_stride, uint64_t n, int d);
void store_drop_last(evdb_store *s);
int store_refinalize(evdb_store *s, uint64_t slot0, uint64_t n, cudaStream_t st);
int store_ensure_capacity(evdb_store *s, uint64_t need);
uint64_t store_device_bytes(const evdb_store *s);
// one handle, N devices (mstore.cu); `owner` mirrors count / dimension for the ABI's argument checks
int mstore_create(evdb_store *owner, const evdb_opts *o);
void mstore_destroy(MStore *m);
int m_put(MStore *m, uint64_t slot0, const void *rows, bool is_f64, uint64_t n, int d, bool replace_all);
int m_bulk_codes(MStore *m, const uint8_t *codes, const double *mins, const double *scales, uint64_t n, int d);
int m_delete(MStore *m, uint32_t slot, int64_t *moved_from);
int m_get_f64(MStore *m, uint32_t slot, double *out, int d);
int m_get_codes(MStore *m, uint32_t slot, uint8_t *codes, double *mn, double *scale);
int m_fill_synthetic(MStore *m, uint64_t seed, uint64_t row0, uint64_t n, int d);
int m_stats(MStore *m, evdb_stats *out);
int m_set_plan(MStore *m, int plan);
int m_flush(MStore *m);
int m_profile(MStore *m, int enable);
int m_profile_read(MStore *m, int32_t *n_samples, double *total_ms);
int m_search_host(MStore *m, const void *queries, bool is_f64, int B, int d, int k, int metric, uint32_t *out_slots,
                  double *out_dists, int32_t *out_counts);

int ensure_bytes(void **p, size_t *cap, size_t need, bool pinned = false);
int ensure_func_smem(const void *fn, size_t smem);                        // cached cudaFuncSetAttribute(MaxDynamicSharedMemorySize)
int cached_occupancy(const void *fn, int threads, size_t smem, int *occ);  // cached cudaOccupancyMaxActiveBlocksPerMultiprocessor
void prof_begin(evdb_store *s, cudaStream_t st);
void prof_end(evdb_store *s, cudaStream_t st);

}  // namespace evdb
