// store.cu -- device-resident store, search orchestration and the C ABI.
//
// The store is what a vector_store gen_server (reference
// src/vector_store.erl:21-35,60-207) keeps in its `vectors` map, re-laid-out as
// packed device columns.  See include/evdb.h for the entry-point <-> reference
// mapping.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>

#include "internal.h"

namespace evdb {

static thread_local char g_cuda_err[256] = "";

void set_cuda_error(cudaError_t e, const char *file, int line) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (%s) at %s:%d", cudaGetErrorName(e),
             cudaGetErrorString(e), file, line);
    cudaGetLastError();  // clear the sticky-less error state
}

int ensure_bytes(void **p, size_t *cap, size_t need, bool pinned) {
    if (need <= *cap && *p) return EVDB_OK;
    size_t ncap = *cap ? *cap : 4096;
    while (ncap < need) ncap *= 2;
    if (*p) {
        if (pinned) cudaFreeHost(*p); else cudaFree(*p);
        *p = nullptr;
        *cap = 0;
    }
    if (pinned) EVDB_CUDA(cudaMallocHost(p, ncap));
    else EVDB_CUDA(cudaMalloc(p, ncap));
    *cap = ncap;
    return EVDB_OK;
}

void prof_begin(evdb_store *s, cudaStream_t st) {
    if (!s->prof_on || s->prof_n >= kProfMax) return;
    cudaEventRecord(s->prof_ev[2 * s->prof_n], st);
    s->prof_stream = st;
}
void prof_end(evdb_store *s, cudaStream_t st) {
    if (!s->prof_on || s->prof_n >= kProfMax) return;
    cudaEventRecord(s->prof_ev[2 * s->prof_n + 1], st);
    s->prof_n++;
}

// cudaFuncSetAttribute / occupancy queries cost microseconds each and a lone query is launch-bound:
// remember, per (device, kernel), the largest dynamic-smem opt-in made and the occupancy found.
struct FuncCacheEntry { const void *fn; int dev; size_t smem_set; size_t occ_smem; int occ_threads; int occ; };
static FuncCacheEntry g_fcache[128];
static int g_fcache_n = 0;
static std::mutex g_fcache_mu;

static FuncCacheEntry *fcache_get(const void *fn) {
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < g_fcache_n; ++i)
        if (g_fcache[i].fn == fn && g_fcache[i].dev == dev) return &g_fcache[i];
    if (g_fcache_n >= 128) return nullptr;
    FuncCacheEntry *e = &g_fcache[g_fcache_n++];
    e->fn = fn; e->dev = dev; e->smem_set = 0; e->occ_smem = ~(size_t)0; e->occ_threads = 0; e->occ = 0;
    return e;
}

int ensure_func_smem(const void *fn, size_t smem) {
    std::lock_guard<std::mutex> lk(g_fcache_mu);
    FuncCacheEntry *e = fcache_get(fn);
    if (e && e->smem_set >= smem) return EVDB_OK;
    EVDB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (e) e->smem_set = smem;
    return EVDB_OK;
}

int cached_occupancy(const void *fn, int threads, size_t smem, int *occ) {
    std::lock_guard<std::mutex> lk(g_fcache_mu);
    FuncCacheEntry *e = fcache_get(fn);
    if (e && e->occ_smem == smem && e->occ_threads == threads) { *occ = e->occ; return EVDB_OK; }
    EVDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, fn, threads, smem));
    if (e) { e->occ_smem = smem; e->occ_threads = threads; e->occ = *occ; }
    return EVDB_OK;
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("EVDB_PDL"); on = e ? (atoi(e) != 0) : 1; }
    return on != 0;
}

static bool is_quant(const evdb_store *s) { return s->dtype == EVDB_U8 || s->dtype == EVDB_U4; }

static int set_device(const evdb_store *s) {
    EVDB_CUDA(cudaSetDevice(s->device));
    return EVDB_OK;
}

static void set_dim(evdb_store *s, int d) {
    s->dim = d;
    switch (s->dtype) {
        case EVDB_F32: s->dpad = round_up(d, 4); s->nch = s->dpad / 4; s->row_bytes = (size_t)s->dpad * 4; break;
        case EVDB_BF16: s->dpad = round_up(d, 8); s->nch = s->dpad / 8; s->row_bytes = (size_t)s->dpad * 2; break;
        case EVDB_U8: s->dpad = round_up(d, 16); s->nch = s->dpad / 16; s->row_bytes = (size_t)s->dpad; break;
        default: s->dpad = round_up(d, 32); s->nch = s->dpad / 32; s->row_bytes = (size_t)s->dpad / 2; break;
    }
    s->spitch = round_up(d, 64);  // 128-byte row pitch: every TMA box row is one aligned line
}

template <typename T>
static int regrow(T **p, uint64_t old_rows, uint64_t new_rows, size_t per_row, cudaStream_t st) {
    T *np = nullptr;
    EVDB_CUDA(cudaMalloc((void **)&np, new_rows * per_row + 256));   // slack: the i8 plan's code tiles may read up to 127 bytes past the last row
    if (*p && old_rows) EVDB_CUDA(cudaMemcpyAsync(np, *p, old_rows * per_row, cudaMemcpyDeviceToDevice, st));
    if (*p) {
        EVDB_CUDA(cudaStreamSynchronize(st));
        cudaFree(*p);
    }
    *p = np;
    return EVDB_OK;
}

static int ensure_capacity(evdb_store *s, uint64_t need) {
    if (need <= s->capacity) return EVDB_OK;
    uint64_t ncap = s->capacity ? s->capacity * 2 : 1024;
    if (ncap < need) ncap = need;
    if (s->capacity == 0 && ncap < s->capacity_hint) ncap = s->capacity_hint;   // evdb_opts.capacity_hint, also when the dimension came later
    uint64_t live = s->count;
    EVDB_TRY(regrow(&s->rows, live, ncap, s->row_bytes, s->stream));
    EVDB_TRY(regrow(&s->norm64, live, ncap, sizeof(double), s->stream));
    EVDB_TRY(regrow(&s->inv_norm, live, ncap, sizeof(float), s->stream));
    EVDB_TRY(regrow(&s->norm_sq, live, ncap, sizeof(float), s->stream));
    if (is_quant(s)) {
        EVDB_TRY(regrow(&s->qcoef, live, ncap, sizeof(float2), s->stream));
        EVDB_TRY(regrow(&s->qms64, live, ncap, sizeof(double2), s->stream));
    }
    if (s->dtype == EVDB_F32 && s->gemm_shadow)
        EVDB_TRY(regrow(&s->shadow, live, ncap, (size_t)s->spitch * sizeof(__half), s->stream));
    s->capacity = ncap;
    s->graph_epoch++;
    return EVDB_OK;
}

static uint64_t device_bytes(const evdb_store *s) {
    uint64_t per = s->row_bytes + sizeof(double) + 2 * sizeof(float);
    if (is_quant(s)) per += sizeof(float2) + sizeof(double2);
    if (s->shadow) per += (uint64_t)s->spitch * 2;
    return per * s->capacity + (s->shadow_l2 ? s->l2_cap * (uint64_t)(s->l2_pitch + 16) * 2 : 0) + s->w_q64_cap + s->w_q32_cap + s->w_qdig_cap + s->w_qstat_cap +
           s->w_partial_cap + s->w_qh_cap + s->w_seed_cap + s->w_ids_cap + s->w_dists_cap + s->w_counts_cap + s->w_tmp_cap + s->w_shard_cap;
}

// ----------------------------------------------------------------------------
// search orchestration (device side, asynchronous on `st`)
// ----------------------------------------------------------------------------
int choose_kp(int kk, int kp_min) {
    int slack = kk / 4 > 6 ? kk / 4 : 6;
    int want = kk + slack;
    if (want < kp_min) want = kp_min;
    if (want < kMinKP) want = kMinKP;
    return next_pow2(want);
}

// d_q64: [B][dim] fp64 on the device.  Outputs [B][kstride].
int search_core(evdb_store *s, const double *d_q64, int B, int k, int kstride, int metric,
                       int kp_min, int plan, uint64_t slot_base, uint64_t *d_ids, double *d_dists,
                       int32_t *d_counts, int32_t *d_flags, cudaStream_t st) {
    const int kk = (uint64_t)k < s->count ? k : (int)s->count;
    if (plan == EVDB_PLAN_AUTO) plan = s->plan;
    int KP = choose_kp(kk, kp_min);
    bool fast_ok = KP <= kMaxKP && !(is_quant(s) && metric != EVDB_COSINE);
    if (plan == EVDB_PLAN_EXACT || !fast_ok) {
        s->last_plan = EVDB_PLAN_EXACT;
        return exact_plan_search(s, d_q64, B, kk, kstride, metric, slot_base, d_ids, d_dists,
                                 d_counts, d_flags, st);
    }
    // (a window wider than the store is fine: KP stays a power of two, lists pad with kKeyMax)
    const double u = 5.9604644775390625e-08;  // 2^-24
    float eps_abs = 0.f, eps_rel = 0.f;
    const float *eps_q = nullptr;
    RawCands raw;
    bool have_raw = false;
    int squared = 0;
    int lists = 0;
    bool use_gemm = false;
    // Auto: the tcgen05 plan reads the 2-byte operand column ONCE for the whole batch, the scan plan
    // reads the 4-byte rows once PER QUERY -- so the GEMM plan also wins for a handful of queries,
    // and even for a single one, as soon as the store is large enough to hide its extra launches
    // (measured on B200, tools/sweep.py: 1 M x 768 B = 1: 0.30 vs 0.52 ms; B = 8: 0.36 vs 3.7 ms;
    // 100 k x 128 B = 1: 0.068 vs 0.057 ms).  EVDB_GEMM_MIN_BYTES overrides the crossover.
    static double gemm_min_bytes = -1.0;
    if (gemm_min_bytes < 0.0) { const char *e = getenv("EVDB_GEMM_MIN_BYTES"); gemm_min_bytes = e ? atof(e) : 256e6; }
    if (plan == EVDB_PLAN_GEMM ||
        (plan == EVDB_PLAN_AUTO && !s->gemm_oom &&
         (B >= 16 || (double)B * (double)s->count * (double)s->row_bytes >= gemm_min_bytes)))
        use_gemm = gemm_plan_supported(s, metric, B, gemm_kp(KP));
    // quantization_8bit stores: batches run as a kind::i8 GEMM over the codes themselves (gemm_i8.cu).  A scan pass
    // serves one query; one GEMM sweep serves up to 128 per CTA at a fixed cost of several scan passes (measured on
    // B200, tools/sweep.py, step time gemm / scan: 12.5 M x 96: B = 4 0.89 / 0.81 ms, B = 6 0.91 / 1.17 ms;
    // 4 M x 256: B = 4 0.41 / 0.66 ms; 1 M x 768: B = 3 0.41 / 0.40 ms, B = 4 0.44 / 0.50 ms).
    // EVDB_QGEMM_MIN_BATCH overrides the crossover.
    bool use_qgemm = false;
    static int qgemm_min_env = -2;
    if (qgemm_min_env == -2) { const char *e = getenv("EVDB_QGEMM_MIN_BATCH"); qgemm_min_env = e ? atoi(e) : -1; }
    const int qgemm_min_batch = qgemm_min_env >= 0 ? qgemm_min_env : (s->dpad <= 128 ? 5 : 4);
    if (s->dtype == EVDB_U8 && metric == EVDB_COSINE &&
        (plan == EVDB_PLAN_GEMM || (plan == EVDB_PLAN_AUTO && !s->gemm_oom && B >= qgemm_min_batch)))
        use_qgemm = qgemm_plan_supported(s, metric, B, gemm_kp(KP));
    if (plan == EVDB_PLAN_GEMM && !use_gemm && !use_qgemm) return EVDB_E_UNSUPPORTED;

    if ((use_gemm || use_qgemm) && B > gemm_max_batch()) {
        // the candidate buffers are sized per sweep: larger batches go through in slices
        const int slice = gemm_max_batch();
        for (int b0 = 0; b0 < B; b0 += slice) {
            const int nb = B - b0 < slice ? B - b0 : slice;
            EVDB_TRY(search_core(s, d_q64 + (size_t)b0 * s->dim, nb, k, kstride, metric, kp_min, plan, slot_base,
                                 d_ids + (size_t)b0 * kstride, d_dists + (size_t)b0 * kstride, d_counts + b0,
                                 d_flags ? d_flags + b0 : nullptr, st));
        }
        return EVDB_OK;
    }
    constexpr int kScanMaxBatch = 32768;   // the scan kernels put the query index in gridDim.y (<= 65535)
    if (!use_gemm && B > kScanMaxBatch) {
        for (int b0 = 0; b0 < B; b0 += kScanMaxBatch) {
            const int nb = B - b0 < kScanMaxBatch ? B - b0 : kScanMaxBatch;
            EVDB_TRY(search_core(s, d_q64 + (size_t)b0 * s->dim, nb, k, kstride, metric, kp_min, plan, slot_base,
                                 d_ids + (size_t)b0 * kstride, d_dists + (size_t)b0 * kstride, d_counts + b0,
                                 d_flags ? d_flags + b0 : nullptr, st));
        }
        return EVDB_OK;
    }
    if (use_gemm) {
        // the per-query error bound of the fp16 operands comes back in eps_q (device)
        const int rc = launch_gemm_topk(s, d_q64, B, gemm_kp(KP), metric, &lists, &eps_q, &raw, st);
        if (rc == EVDB_E_OOM && plan == EVDB_PLAN_AUTO) {
            // no room for the euclidean operand column / candidate buffers: the scan plan needs neither
            (void)cudaGetLastError();
            s->gemm_oom = 1;
            use_gemm = false;
            eps_q = nullptr;
        } else {
            EVDB_TRY(rc);
            s->last_plan = EVDB_PLAN_GEMM;
            KP = gemm_kp(KP);
            have_raw = true;
            squared = metric == EVDB_EUCLIDEAN;
        }
    }
    if (use_qgemm) {
        const int rc = launch_qgemm_topk(s, d_q64, B, gemm_kp(KP), &lists, &eps_q, &raw, st);
        if (rc == EVDB_E_OOM && plan == EVDB_PLAN_AUTO) {
            (void)cudaGetLastError();
            s->gemm_oom = 1;
            use_qgemm = false;
            eps_q = nullptr;
        } else {
            EVDB_TRY(rc);
            s->last_plan = EVDB_PLAN_GEMM;
            KP = gemm_kp(KP);
            have_raw = true;
            // exact integer digit sums; fp32 from there on: two int->float, the combine, cy*sum(Q), two fmas
            // (48 u leaves the same head-room over the count as the scan's 24 u); + the query-grid bound in eps_q
            eps_abs = (float)(48.0 * u);
        }
    }
    if (!use_gemm && !use_qgemm) {
        // Small float stores, a lone query: prep + scan + selection in ONE launch (select.cu small_fused_kernel)
        // -- three dependent kernels cost more than their work.
        static int fused_on = -1;
        if (fused_on < 0) { const char *e = getenv("EVDB_FUSED_SMALL"); fused_on = e ? atoi(e) : 1; }
        // (measured on B200, tools/sweep.py, device-timed step: 10 k x 128 B = 1: 29.9 us fused against 36.1; 100 k x 128:
        //  48.3 against 48.4; 20 k x 768 B = 4: 131 against 75 -- with long rows or several queries the one CTA that
        //  selects is slower than the 1024-thread select kernel, so the fused launch is kept for short rows and lone queries)
        if (fused_on && (s->dtype == EVDB_F32 || s->dtype == EVDB_BF16) && B <= 2 && KP <= 128 && s->dim <= 256 &&
            (double)s->count * (double)s->row_bytes <= 32e6) {
            int G = 0, tpr = 0;
            if (scan_small_plan(s, metric, KP, &G, &tpr) == EVDB_OK) {
                EVDB_TRY(ensure_bytes((void **)&s->w_partial, &s->w_partial_cap, sizeof(uint64_t) * (size_t)B * G * KP));
                EVDB_TRY(ensure_bytes((void **)&s->w_qeps, &s->w_qeps_cap, sizeof(float) * (size_t)B));
                if (!s->d_arrive) {
                    EVDB_CUDA(cudaMalloc((void **)&s->d_arrive, 8 * sizeof(unsigned int)));
                    EVDB_CUDA(cudaMemsetAsync(s->d_arrive, 0, 8 * sizeof(unsigned int), st));
                }
                const double depth = (double)s->dim / 64.0 + 24.0;
                const float ea = metric == EVDB_COSINE ? (float)(depth * u) : 1e-37f;
                const float er = metric == EVDB_COSINE ? 0.f : (float)(depth * u);
                prof_begin(s, st);
                const int rc = launch_small_fused(s, d_q64, B, KP, kk, kstride, metric, G, tpr, s->w_partial, s->w_qeps,
                                                  s->d_arrive, ea, er, slot_base, d_ids, d_dists, d_counts, d_flags, st);
                prof_end(s, st);
                if (rc == EVDB_OK) {
                    s->last_plan = EVDB_PLAN_SCAN;
                    s->n_rows_scanned += (uint64_t)B * s->count;
                    return EVDB_OK;
                }
                if (rc != EVDB_E_UNSUPPORTED) return rc;
                if (s->prof_on && s->prof_n > 0) s->prof_n--;   // the bracket measured nothing
            }
        }
        EVDB_TRY(launch_prep_queries(s, d_q64, B, metric, st));
        s->last_plan = EVDB_PLAN_SCAN;
        int G = 0;
        int rc = scan_grid_size(s, metric, KP, B, &G);
        if (rc == EVDB_E_UNSUPPORTED) {
            s->last_plan = EVDB_PLAN_EXACT;
            return exact_plan_search(s, d_q64, B, kk, kstride, metric, slot_base, d_ids, d_dists,
                                     d_counts, d_flags, st);
        }
        EVDB_TRY(rc);
        EVDB_TRY(ensure_bytes((void **)&s->w_partial, &s->w_partial_cap,
                              sizeof(uint64_t) * (size_t)B * G * KP));
        ScanArgs a;
        a.rows = s->rows; a.row_bytes = s->row_bytes; a.nch = s->nch; a.n = s->count;
        a.inv_norm = s->inv_norm; a.qcoef = s->qcoef;
        a.q32 = s->w_q32; a.q32_stride = s->dpad;
        a.qdig = s->w_qdig; a.qdig_stride = s->dpad;
        a.qstat = s->w_qstat; a.partial = s->w_partial; a.KP = KP; a.G = G; a.B = B;
        prof_begin(s, st);
        EVDB_TRY(launch_scan(s, metric, a, st));
        prof_end(s, st);
        lists = G;
        double depth = (double)s->dim / 64.0 + 24.0;
        if (is_quant(s)) { eps_abs = (float)(24.0 * u); eps_q = s->w_qeps; }   // + the query-grid bound, per query
        else if (metric == EVDB_COSINE) eps_abs = (float)(depth * u);
        else { eps_rel = (float)(depth * u); eps_abs = 1e-37f; eps_q = s->w_qeps; }   // + the fp32 narrowing residual of the query
    }
    // EVDB_PROF_SELECT=1 (tuning aid): the profiling bracket times the select kernel instead of the scan/GEMM
    static int prof_sel = -1;
    if (prof_sel < 0) { const char *e = getenv("EVDB_PROF_SELECT"); prof_sel = e && atoi(e) ? 1 : 0; }
    if (prof_sel) { s->prof_n = s->prof_n > 0 ? s->prof_n - 1 : 0; prof_begin(s, st); }
    EVDB_TRY(launch_select(s, d_q64, have_raw ? nullptr : s->w_partial, have_raw ? &raw : nullptr, lists, KP, B, kk, kstride, metric, eps_abs,
                           eps_rel, eps_q, squared, slot_base, d_ids, d_dists, d_counts, d_flags, st));
    if (prof_sel) prof_end(s, st);
    s->n_rows_scanned += (uint64_t)B * s->count;
    return EVDB_OK;
}

__global__ void widen_kernel(const float *__restrict__ in, double *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (double)in[i];
}

// exponent-field test on the raw words, accumulated without an early exit: the loop vectorises
static bool all_finite(const void *v, bool is_f64, size_t n) {
    uint32_t bad = 0;
    if (is_f64) {
        const uint64_t *w = (const uint64_t *)v;
        for (size_t i = 0; i < n; ++i) bad |= (uint32_t)(((w[i] >> 52) & 0x7FFu) == 0x7FFu);
    } else {
        const uint32_t *w = (const uint32_t *)v;
        for (size_t i = 0; i < n; ++i) bad |= (uint32_t)(((w[i] >> 23) & 0xFFu) == 0xFFu);
    }
    return bad == 0;
}

// Host-facing search: H2D queries, search, D2H results, escalate flagged queries.
static int search_host(evdb_store *s, const void *queries, bool is_f64, int B, int d, int k,
                       int metric, uint32_t *out_slots, double *out_dists, int32_t *out_counts) {
    if (!s || B < 0 || k < 0 || (B > 0 && !queries) || metric < 0 || metric > 2) return EVDB_E_BAD_ARG;
    if (B == 0) return EVDB_OK;
    if (!out_counts || (k > 0 && (!out_slots || !out_dists))) return EVDB_E_BAD_ARG;
    // validate_vector(Q, undefined) accepts any length on an empty store -> {ok, []}
    if (s->dim == 0 || s->count == 0) {
        if (s->dim != 0 && d != s->dim) return EVDB_E_DIM_MISMATCH;
        for (int b = 0; b < B; ++b) out_counts[b] = 0;
        return EVDB_OK;
    }
    if (d != s->dim) return EVDB_E_DIM_MISMATCH;
    if (s->multi) {
        if (k == 0) { for (int b = 0; b < B; ++b) out_counts[b] = 0; return all_finite(queries, is_f64, (size_t)B * d) ? EVDB_OK : EVDB_E_BAD_VECTOR; }
        return m_search_host(s->multi, queries, is_f64, B, d, k, metric, out_slots, out_dists, out_counts);
    }
    size_t nq = (size_t)B * d;
    // validate_vector/2 (lists:all(is_number)): a small query is checked before anything is
    // enqueued; a large batch is checked on the host WHILE the device works on it (the result is
    // discarded if the check fails), so the check costs no latency.
    const bool check_late = nq >= 65536;
    if (!check_late && !all_finite(queries, is_f64, nq)) return EVDB_E_BAD_VECTOR;
    EVDB_TRY(set_device(s));
    if (k == 0) {
        for (int b = 0; b < B; ++b) out_counts[b] = 0;
        return EVDB_OK;
    }
    cudaStream_t st = s->stream;
    const int kk = (uint64_t)k < s->count ? k : (int)s->count;
    const int kstride = kk;
    const size_t esz = is_f64 ? sizeof(double) : sizeof(float);
    const size_t nk = (size_t)B * kstride;
    // one result blob on the device, one copy back: [ids u64][distances f64][counts i32][flags i32]
    const size_t pin_need = nk * (sizeof(uint64_t) + sizeof(double)) + sizeof(int32_t) * 2 * (size_t)B;
    EVDB_TRY(ensure_bytes((void **)&s->w_q64, &s->w_q64_cap, nq * sizeof(double)));
    EVDB_TRY(ensure_bytes((void **)&s->w_ids, &s->w_ids_cap, pin_need));
    EVDB_TRY(ensure_bytes(&s->h_pin, &s->h_pin_cap, pin_need, true));
    if (!is_f64) EVDB_TRY(ensure_bytes(&s->w_tmp, &s->w_tmp_cap, nq * sizeof(float)));
    uint64_t *d_ids = s->w_ids;
    double *d_dists = (double *)(d_ids + nk);
    int32_t *d_counts = (int32_t *)(d_dists + nk), *d_flags = d_counts + B;
    uint64_t *h_ids = (uint64_t *)s->h_pin;
    double *h_d = (double *)(h_ids + nk);
    int32_t *h_c = (int32_t *)(h_d + nk);

    // ---- small batches: the whole call as ONE replayed CUDA graph ----------------------------------
    // A lone query on a small store is bound by the host: a dozen driver calls (copies, launches, event
    // records) cost more than the device work.  The second consecutive call of the same shape is
    // captured -- query copy from a pinned slot, every kernel of the plan, result copy -- and later calls
    // replay it with one cudaGraphLaunch.  Any ingest, delete, plan change or reallocation bumps
    // graph_epoch and drops the graph; profiling (evdb_store_profile) and debug knobs bypass it.
    GraphKey key;
    key.B = B; key.k = k; key.metric = metric; key.f64 = is_f64 ? 1 : 0; key.count = s->count; key.epoch = s->graph_epoch;
    key.q = s->w_q64; key.out = s->w_ids; key.pin = s->h_pin;
    {   // a call of another shape may have regrown (= moved) a workspace since the capture: the graph must not outlive it
        const void *ws[] = {s->rows, s->norm64, s->inv_norm, s->qcoef, s->qms64, s->shadow, s->shadow_l2, s->l2_tail, s->w_q32,
                            s->w_qdig, s->w_seed, s->w_qh, s->w_qstat, s->w_qeps, s->w_partial, s->w_tmp, s->d_arrive, s->d_scalar};
        uint64_t h = 1469598103934665603ull;
        for (const void *p : ws) h = (h ^ (uint64_t)(uintptr_t)p) * 1099511628211ull;
        key.ws = h;
    }
    static int graphs_on = -1;
    if (graphs_on < 0) { const char *e = getenv("EVDB_GRAPHS"); graphs_on = e ? atoi(e) : 1; }
    const bool small = graphs_on && B <= 16 && !s->prof_on && !s->graph_broken && nq * esz <= (64u << 10);
    bool replayed = false, capturing = false;
    if (small && s->gexec && key == s->gkey) {
        memcpy(s->h_gq, queries, nq * esz);
        EVDB_CUDA(cudaEventRecord(s->ev0, st));
        EVDB_CUDA(cudaGraphLaunch(s->gexec, st));
        EVDB_CUDA(cudaEventRecord(s->ev1, st));
        s->n_launches += s->graph_launches;
        s->last_plan = s->graph_plan;
        s->n_rows_scanned += (uint64_t)B * s->count;
        replayed = true;
    } else if (small && key == s->last_key) {
        if (s->gexec) { cudaGraphExecDestroy(s->gexec); s->gexec = nullptr; }
        EVDB_TRY(ensure_bytes(&s->h_gq, &s->h_gq_cap, 64u << 10, true));
        key.pin = s->h_pin;
        memcpy(s->h_gq, queries, nq * esz);
        capturing = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (!capturing) { cudaGetLastError(); s->graph_broken = 1; }
    }
    s->last_key = key;
    if (!replayed) {
        const uint64_t launches0 = s->n_launches;
        const void *src = capturing ? s->h_gq : queries;
        int rc = EVDB_OK;
        do {
#define SH(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_cuda_error(_e, __FILE__, __LINE__); rc = _e == cudaErrorMemoryAllocation ? EVDB_E_OOM : EVDB_E_CUDA; } } while (0)
            if (!capturing) SH(cudaEventRecord(s->ev0, st));
            if (rc != EVDB_OK) break;
            if (is_f64) {
                SH(cudaMemcpyAsync(s->w_q64, src, nq * sizeof(double), cudaMemcpyHostToDevice, st));
            } else {
                SH(cudaMemcpyAsync(s->w_tmp, src, nq * sizeof(float), cudaMemcpyHostToDevice, st));
                widen_kernel<<<(int)((nq + 255) / 256 < 1024 ? (nq + 255) / 256 : 1024), 256, 0, st>>>(
                    (const float *)s->w_tmp, s->w_q64, nq);
                s->n_launches++;
                SH(cudaGetLastError());
            }
            if (rc != EVDB_OK) break;
            if (!capturing) SH(cudaEventRecord(s->ev2, st));
            rc = search_core(s, s->w_q64, B, k, kstride, metric, 0, EVDB_PLAN_AUTO, 0, d_ids, d_dists, d_counts, d_flags, st);
            if (rc != EVDB_OK) break;
            if (!capturing) SH(cudaEventRecord(s->ev3, st));
            SH(cudaMemcpyAsync(s->h_pin, s->w_ids, pin_need, cudaMemcpyDeviceToHost, st));
            if (!capturing) SH(cudaEventRecord(s->ev1, st));
#undef SH
        } while (0);
        if (capturing) {
            cudaGraph_t g = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(st, &g);
            if (rc == EVDB_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&s->gexec, g, 0) == cudaSuccess) {
                s->gkey = key;
                s->graph_launches = s->n_launches - launches0;
                s->graph_plan = s->last_plan;
                EVDB_CUDA(cudaEventRecord(s->ev0, st));
                EVDB_CUDA(cudaGraphLaunch(s->gexec, st));
                EVDB_CUDA(cudaEventRecord(s->ev1, st));
                replayed = true;
            } else {
                // something in this plan cannot be captured (it allocates, synchronises or failed): never try again
                cudaGetLastError();
                s->gexec = nullptr;
                s->graph_broken = 1;
                if (g) cudaGraphDestroy(g);
                return search_host(s, queries, is_f64, B, d, k, metric, out_slots, out_dists, out_counts);
            }
            if (g) cudaGraphDestroy(g);
        } else if (rc != EVDB_OK) {
            return rc;
        }
    }
    const bool bad_query = check_late && !all_finite(queries, is_f64, nq);
    EVDB_CUDA(cudaStreamSynchronize(st));
    if (bad_query) return EVDB_E_BAD_VECTOR;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    s->last_search_ms = ms;
    if (!replayed) {
        cudaEventElapsedTime(&ms, s->ev0, s->ev2); s->last_h2d_ms = ms;
        cudaEventElapsedTime(&ms, s->ev2, s->ev3); s->last_device_ms = ms;
        cudaEventElapsedTime(&ms, s->ev3, s->ev1); s->last_d2h_ms = ms;
    } else {
        s->last_h2d_ms = 0.0; s->last_device_ms = ms; s->last_d2h_ms = 0.0;   // one graph: copies included
    }
    s->n_searches += (uint64_t)B;

    // escalate queries whose candidate window could not be proven complete
    int first_plan = s->last_plan;
    for (int b = 0; b < B; ++b) {
        if (!h_c[B + b]) continue;
        s->n_escalations++;
        int kp_min = 256;
        for (int attempt = 0; attempt < 2; ++attempt) {
            int plan = attempt == 0 ? EVDB_PLAN_SCAN : EVDB_PLAN_EXACT;
            if (attempt == 0 && (choose_kp(kk, kp_min) > kMaxKP)) continue;
            const double *dq = s->w_q64 + (size_t)b * d;
            EVDB_TRY(search_core(s, dq, 1, k, kstride, metric, kp_min, plan, 0,
                                 d_ids + (size_t)b * kstride, d_dists + (size_t)b * kstride,
                                 d_counts + b, d_flags + b, st));
            EVDB_CUDA(cudaMemcpyAsync(h_ids + (size_t)b * kstride, d_ids + (size_t)b * kstride,
                                      kstride * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            EVDB_CUDA(cudaMemcpyAsync(h_d + (size_t)b * kstride, d_dists + (size_t)b * kstride,
                                      kstride * sizeof(double), cudaMemcpyDeviceToHost, st));
            EVDB_CUDA(cudaMemcpyAsync(h_c + b, d_counts + b, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            EVDB_CUDA(cudaMemcpyAsync(h_c + B + b, d_flags + b, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            EVDB_CUDA(cudaStreamSynchronize(st));
            if (!h_c[B + b]) break;
        }
    }
    s->last_plan = first_plan;
    for (int b = 0; b < B; ++b) {
        out_counts[b] = h_c[b];
        for (int j = 0; j < k; ++j) {
            size_t o = (size_t)b * k + j;
            if (j < h_c[b]) {
                out_slots[o] = (uint32_t)h_ids[(size_t)b * kstride + j];
                out_dists[o] = h_d[(size_t)b * kstride + j];
            } else {
                out_slots[o] = 0xFFFFFFFFu;
                out_dists[o] = 0.0;
            }
        }
    }
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// ingest helpers
// ----------------------------------------------------------------------------
// validate_vector/2 (reference src/vector_store.erl:213-225): the length is compared first, the
// dimension of an empty store is fixed only by a vector that passed every check (:128-131)
static int check_dim(const evdb_store *s, int d) {
    if (d <= 0) return EVDB_E_BAD_VECTOR;
    if (s->dim == 0) return EVDB_OK;
    return d == s->dim ? EVDB_OK : EVDB_E_DIM_MISMATCH;
}
static void fix_dim(evdb_store *s, int d) { if (s->dim == 0) set_dim(s, d); }

template <typename T>
static bool rows_finite(const T *rows, uint64_t n) { return all_finite(rows, sizeof(T) == 8, (size_t)n); }

int launch_narrow_rows(evdb_store *s, uint64_t dst0, const void *src, bool is_f64, uint64_t n,
                       cudaStream_t st);

// rows: n host rows of d values (fp64 or fp32), `pitch` elements apart (pitch == d: dense; a
// multi-device store hands every shard each S-th row of the caller's array) -> slots [slot0, slot0+n)
static int ingest_rows_big(evdb_store *s, uint64_t slot0, const void *rows, bool is_f64, uint64_t n, size_t pitch) {
    const int d = s->dim;
    cudaStream_t st = s->stream;
    const size_t esz = is_f64 ? sizeof(double) : sizeof(float);
    // chunk so that staging stays <= 256 MiB
    uint64_t chunk = (256ull << 20) / ((size_t)d * esz);
    if (chunk < 1) chunk = 1;
    for (uint64_t r0 = 0; r0 < n; r0 += chunk) {
        uint64_t cnt = n - r0 < chunk ? n - r0 : chunk;
        const uint8_t *src = (const uint8_t *)rows + r0 * pitch * esz;
        uint64_t dst0 = slot0 + r0;
        if (s->dtype == EVDB_F32 && !is_f64) {
            EVDB_CUDA(cudaMemcpy2DAsync(s->rows + dst0 * s->row_bytes, s->row_bytes, src, pitch * 4,
                                        (size_t)d * 4, cnt, cudaMemcpyHostToDevice, st));
            if (s->dpad != d)
                EVDB_CUDA(cudaMemset2DAsync(s->rows + dst0 * s->row_bytes + (size_t)d * 4, s->row_bytes, 0,
                                            (size_t)(s->dpad - d) * 4, cnt, st));
        } else {
            size_t bytes = cnt * (size_t)d * esz;
            EVDB_TRY(ensure_bytes(&s->w_tmp, &s->w_tmp_cap, bytes));
            if (pitch == (size_t)d) EVDB_CUDA(cudaMemcpyAsync(s->w_tmp, src, bytes, cudaMemcpyHostToDevice, st));
            else EVDB_CUDA(cudaMemcpy2DAsync(s->w_tmp, (size_t)d * esz, src, pitch * esz, (size_t)d * esz, cnt,
                                             cudaMemcpyHostToDevice, st));
            if (is_quant(s)) {
                EVDB_TRY(launch_quantize_rows(s->dtype, is_f64 ? (const double *)s->w_tmp : nullptr,
                                              is_f64 ? nullptr : (const float *)s->w_tmp, cnt, d,
                                              s->rows + dst0 * s->row_bytes, s->row_bytes,
                                              s->qms64 + dst0, nullptr, nullptr, st));
            } else {
                EVDB_TRY(launch_narrow_rows(s, dst0, s->w_tmp, is_f64, cnt, st));
            }
            s->n_launches++;
        }
        EVDB_TRY(launch_finalize_rows(s, dst0, cnt, st));
        EVDB_TRY(launch_l2_shadow_rows(s, dst0, cnt, st));
        EVDB_CUDA(cudaStreamSynchronize(st));
    }
    s->max_norm_dirty = 1;
    s->graph_epoch++;
    return EVDB_OK;
}

// Inserts at rate (reference handle_call({insert,..}), src/vector_store.erl:113-141: one call per vector).
// The rows are copied into a PINNED staging ring (two halves, an event each), so the caller's buffer
// is free when the call returns, the host->device copy and the row kernels are only ENQUEUED on the
// store's stream, and nothing waits: a later search on that stream is ordered behind them, a half
// is waited for only when it comes round again while still in flight.  evdb_store_flush drains.
constexpr size_t kRingHalf = 4u << 20;
static int ensure_ring(evdb_store *s) {
    if (s->h_ring) return EVDB_OK;
    EVDB_CUDA(cudaMallocHost((void **)&s->h_ring, 2 * kRingHalf));
    EVDB_CUDA(cudaMalloc((void **)&s->d_ring, 2 * kRingHalf));
    for (int i = 0; i < 2; ++i) EVDB_CUDA(cudaEventCreateWithFlags(&s->ring_ev[i], cudaEventDisableTiming));
    EVDB_CUDA(cudaEventCreateWithFlags(&s->ev_ing, cudaEventDisableTiming));
    return EVDB_OK;
}

static int ingest_rows(evdb_store *s, uint64_t slot0, const void *rows, bool is_f64, uint64_t n, size_t pitch = 0) {
    const int d = s->dim;
    if (pitch == 0) pitch = (size_t)d;
    const size_t esz = is_f64 ? sizeof(double) : sizeof(float), rowb = (size_t)d * esz;
    // loads of many megabytes keep the chunked synchronous path (the wait is amortised, no second host copy)
    if (rowb > kRingHalf || n * rowb >= ((size_t)64 << 20)) return ingest_rows_big(s, slot0, rows, is_f64, n, pitch);
    EVDB_TRY(ensure_ring(s));
    cudaStream_t st = s->stream;
    for (uint64_t r0 = 0; r0 < n;) {
        // consecutive small upserts share a half (at growing offsets); a half that is full is handed over
        // with an event and waited for only when the ring comes round to it again
        int h = (int)(s->ring_pos & 1);
        if (s->ring_off + rowb > kRingHalf) {
            EVDB_CUDA(cudaEventRecord(s->ring_ev[h], st));
            s->ring_busy[h] = 1;
            s->ring_pos++;
            s->ring_off = 0;
            h ^= 1;
            if (s->ring_busy[h]) { EVDB_CUDA(cudaEventSynchronize(s->ring_ev[h])); s->ring_busy[h] = 0; }
        }
        const uint64_t room = (kRingHalf - s->ring_off) / rowb;
        const uint64_t cnt = n - r0 < room ? n - r0 : room;
        const size_t off = (size_t)h * kRingHalf + s->ring_off;
        uint8_t *hp = s->h_ring + off, *dp = s->d_ring + off;
        const uint8_t *src = (const uint8_t *)rows + r0 * pitch * esz;
        if (pitch == (size_t)d) memcpy(hp, src, cnt * rowb);
        else for (uint64_t i = 0; i < cnt; ++i) memcpy(hp + i * rowb, src + i * pitch * esz, rowb);
        EVDB_CUDA(cudaMemcpyAsync(dp, hp, cnt * rowb, cudaMemcpyHostToDevice, st));
        const uint64_t dst0 = slot0 + r0;
        if (is_quant(s)) {
            EVDB_TRY(launch_quantize_rows(s->dtype, is_f64 ? (const double *)dp : nullptr, is_f64 ? nullptr : (const float *)dp,
                                          cnt, d, s->rows + dst0 * s->row_bytes, s->row_bytes, s->qms64 + dst0, nullptr, nullptr, st));
            s->n_launches++;
            EVDB_TRY(launch_finalize_rows(s, dst0, cnt, st));
        } else {
            EVDB_TRY(launch_finalize_rows(s, dst0, cnt, st, dp, is_f64));   // narrows the staged rows into place first
        }
        EVDB_TRY(launch_l2_shadow_rows(s, dst0, cnt, st));
        s->ring_off += (cnt * rowb + 255) & ~(size_t)255;
        r0 += cnt;
    }
    s->ingest_pending = 1;
    s->max_norm_dirty = 1;
    s->graph_epoch++;
    return EVDB_OK;
}

// device work enqueued on the store's own stream (ingest, delete) must precede a search on a caller's stream
static int order_after_ingest(evdb_store *s, cudaStream_t st) {
    if (!s->ingest_pending || st == s->stream || !s->ev_ing) return EVDB_OK;
    EVDB_CUDA(cudaEventRecord(s->ev_ing, s->stream));
    EVDB_CUDA(cudaStreamWaitEvent(st, s->ev_ing, 0));
    return EVDB_OK;
}

// swap-with-last delete in ONE launch: every column of row `last` -> row `slot`
__global__ void __launch_bounds__(256) delete_swap_kernel(uint8_t *rows, size_t row_bytes, double *norm64, float *inv_norm,
                                                          float *norm_sq, float2 *qcoef, double2 *qms64, __half *shadow,
                                                          int spitch, __half *shadow_l2, __half *l2_tail, int l2_pitch,
                                                          uint64_t slot, uint64_t last) {
    const uint4 *src = reinterpret_cast<const uint4 *>(rows + last * row_bytes);
    uint4 *dst = reinterpret_cast<uint4 *>(rows + slot * row_bytes);
    for (size_t i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) dst[i] = src[i];
    if (shadow) {
        const uint4 *s2 = reinterpret_cast<const uint4 *>(shadow + last * (size_t)spitch);
        uint4 *d2 = reinterpret_cast<uint4 *>(shadow + slot * (size_t)spitch);
        for (int i = threadIdx.x; i < spitch / 8; i += blockDim.x) d2[i] = s2[i];
    }
    if (shadow_l2) {
        const uint4 *s3 = reinterpret_cast<const uint4 *>(shadow_l2 + last * (size_t)l2_pitch);
        uint4 *d3 = reinterpret_cast<uint4 *>(shadow_l2 + slot * (size_t)l2_pitch);
        for (int i = threadIdx.x; i < l2_pitch / 8; i += blockDim.x) d3[i] = s3[i];
        if (threadIdx.x < 2) reinterpret_cast<uint4 *>(l2_tail + slot * 16)[threadIdx.x] = reinterpret_cast<const uint4 *>(l2_tail + last * 16)[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        norm64[slot] = norm64[last];
        inv_norm[slot] = inv_norm[last];
        norm_sq[slot] = norm_sq[last];
        if (qcoef) { qcoef[slot] = qcoef[last]; qms64[slot] = qms64[last]; }
    }
}

template <typename SRC, int DTYPE>
__global__ void narrow_rows_kernel(const SRC *__restrict__ src, uint8_t *__restrict__ rows,
                                   size_t row_bytes, int d, int dpad, uint64_t n) {
    uint64_t total = n * (uint64_t)dpad;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / dpad;
        int c = (int)(i % dpad);
        float v = c < d ? (float)src[r * (uint64_t)d + c] : 0.0f;
        if (DTYPE == EVDB_F32) reinterpret_cast<float *>(rows + r * row_bytes)[c] = v;
        else reinterpret_cast<__nv_bfloat16 *>(rows + r * row_bytes)[c] = __float2bfloat16_rn(v);
    }
}

int launch_narrow_rows(evdb_store *s, uint64_t dst0, const void *src, bool is_f64, uint64_t n,
                       cudaStream_t st) {
    uint64_t blocks = (n * (uint64_t)s->dpad + 255) / 256;
    int grid = (int)(blocks < (uint64_t)s->sm_count * 16 ? blocks : (uint64_t)s->sm_count * 16);
    uint8_t *dst = s->rows + dst0 * s->row_bytes;
    if (s->dtype == EVDB_F32) {
        if (is_f64) narrow_rows_kernel<double, EVDB_F32><<<grid, 256, 0, st>>>((const double *)src, dst, s->row_bytes, s->dim, s->dpad, n);
        else narrow_rows_kernel<float, EVDB_F32><<<grid, 256, 0, st>>>((const float *)src, dst, s->row_bytes, s->dim, s->dpad, n);
    } else {
        if (is_f64) narrow_rows_kernel<double, EVDB_BF16><<<grid, 256, 0, st>>>((const double *)src, dst, s->row_bytes, s->dim, s->dpad, n);
        else narrow_rows_kernel<float, EVDB_BF16><<<grid, 256, 0, st>>>((const float *)src, dst, s->row_bytes, s->dim, s->dpad, n);
    }
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

template <typename T>
static int upsert_any(evdb_store *s, uint32_t slot, const T *vec, int d, bool is_f64) {
    if (!s || !vec) return EVDB_E_BAD_ARG;
    EVDB_TRY(check_dim(s, d));
    if (!rows_finite(vec, (uint64_t)d)) return EVDB_E_BAD_VECTOR;
    if ((uint64_t)slot > s->count) return EVDB_E_BAD_ARG;
    if (s->multi) return m_put(s->multi, slot, vec, is_f64, 1, d, false);
    fix_dim(s, d);
    EVDB_TRY(set_device(s));
    EVDB_TRY(ensure_capacity(s, (uint64_t)slot + 1));
    EVDB_TRY(ingest_rows(s, slot, vec, is_f64, 1));
    if ((uint64_t)slot == s->count) s->count++;
    s->n_upserts++;
    return EVDB_OK;
}

template <typename T>
static int bulk_any(evdb_store *s, const T *rows, uint64_t n, int d, bool is_f64) {
    if (!s || (n > 0 && !rows)) return EVDB_E_BAD_ARG;
    if (n == 0) {
        if (s->multi) return m_put(s->multi, 0, nullptr, is_f64, 0, s->dim, true);
        s->count = 0; s->shadow_valid = 0; s->l2_valid = 0;
        return EVDB_OK;
    }
    EVDB_TRY(check_dim(s, d));
    if (n > 0xFFFFFFF0ull) return EVDB_E_BAD_ARG;
    if (!rows_finite(rows, n * (uint64_t)d)) return EVDB_E_BAD_VECTOR;
    if (s->multi) return m_put(s->multi, 0, rows, is_f64, n, d, true);
    fix_dim(s, d);
    EVDB_TRY(set_device(s));
    EVDB_TRY(ensure_capacity(s, n));
    s->count = 0;
    s->shadow_valid = 0;
    s->l2_valid = 0;
    EVDB_TRY(ingest_rows(s, 0, rows, is_f64, n));
    s->count = n;
    s->n_upserts += n;
    return EVDB_OK;
}

// n rows appended at slots [count, count + n): one H2D and one finalize pass for the lot
template <typename T>
static int append_any(evdb_store *s, const T *rows, uint64_t n, int d, bool is_f64, uint64_t *first_slot) {
    if (!s || (n > 0 && !rows)) return EVDB_E_BAD_ARG;
    if (first_slot) *first_slot = s->count;
    if (n == 0) return EVDB_OK;
    EVDB_TRY(check_dim(s, d));
    if (s->count + n > 0xFFFFFFF0ull) return EVDB_E_BAD_ARG;
    if (!rows_finite(rows, n * (uint64_t)d)) return EVDB_E_BAD_VECTOR;
    if (s->multi) return m_put(s->multi, s->count, rows, is_f64, n, d, false);
    fix_dim(s, d);
    EVDB_TRY(set_device(s));
    EVDB_TRY(ensure_capacity(s, s->count + n));
    EVDB_TRY(ingest_rows(s, s->count, rows, is_f64, n));
    s->count += n;
    s->n_upserts += n;
    return EVDB_OK;
}

// validated rows -> slots [slot0, slot0 + n) of one device store (a shard of a multi-device store
// gets every S-th row of the caller's array: `pitch` elements between consecutive source rows)
int store_put_rows(evdb_store *s, uint64_t slot0, const void *rows, bool is_f64, uint64_t n, size_t pitch, int d) {
    if (n == 0) return EVDB_OK;
    if (slot0 > s->count || slot0 + n > 0xFFFFFFF0ull) return EVDB_E_BAD_ARG;
    EVDB_TRY(check_dim(s, d));
    fix_dim(s, d);
    EVDB_TRY(set_device(s));
    EVDB_TRY(ensure_capacity(s, slot0 + n));
    EVDB_TRY(ingest_rows(s, slot0, rows, is_f64, n, pitch));
    if (slot0 + n > s->count) s->count = slot0 + n;
    s->n_upserts += n;
    return EVDB_OK;
}

// compressed records -> slots [0, n): codes `code_pitch` bytes apart, mins / scales `ms_stride` doubles apart
int store_load_codes(evdb_store *s, const uint8_t *codes, size_t code_pitch, const double *mins, const double *scales,
                     size_t ms_stride, uint64_t n, int d) {
    if (n == 0) { s->count = 0; return EVDB_OK; }
    EVDB_TRY(check_dim(s, d));
    fix_dim(s, d);
    EVDB_TRY(set_device(s));
    EVDB_TRY(ensure_capacity(s, n));
    s->count = 0;
    cudaStream_t st = s->stream;
    const size_t src_row = s->dtype == EVDB_U8 ? (size_t)d : (size_t)(d + 1) / 2;
    EVDB_CUDA(cudaMemset2DAsync(s->rows, s->row_bytes, 0, s->row_bytes, n, st));
    EVDB_CUDA(cudaMemcpy2DAsync(s->rows, s->row_bytes, codes, code_pitch, src_row, n, cudaMemcpyHostToDevice, st));
    // interleave {min, scale} on the host (n pairs), one H2D
    EVDB_TRY(ensure_bytes(&s->h_pin, &s->h_pin_cap, n * sizeof(double2), true));
    double2 *hp = (double2 *)s->h_pin;
    for (uint64_t i = 0; i < n; ++i) { hp[i].x = mins[i * ms_stride]; hp[i].y = scales[i * ms_stride]; }
    EVDB_CUDA(cudaMemcpyAsync(s->qms64, hp, n * sizeof(double2), cudaMemcpyHostToDevice, st));
    EVDB_TRY(launch_finalize_rows(s, 0, n, st));
    EVDB_CUDA(cudaStreamSynchronize(st));
    s->count = n;
    s->max_norm_dirty = 1;
    s->graph_epoch++;
    return EVDB_OK;
}

// drop the last row (the tail of a swap-with-last delete whose hole was filled from elsewhere)
void store_drop_last(evdb_store *s) {
    if (s->count == 0) return;
    s->count--;
    s->graph_epoch++;
    if (s->shadow_valid > s->count) s->shadow_valid = s->count;
    if (s->l2_valid > s->count) s->l2_valid = s->count;
}

// refresh the cached per-row values of slots [slot0, slot0 + n) after their raw columns were written
int store_refinalize(evdb_store *s, uint64_t slot0, uint64_t n, cudaStream_t st) {
    EVDB_TRY(launch_finalize_rows(s, slot0, n, st));
    EVDB_TRY(launch_l2_shadow_rows(s, slot0, n, st));
    s->max_norm_dirty = 1;
    s->graph_epoch++;
    return EVDB_OK;
}

int store_ensure_capacity(evdb_store *s, uint64_t need) { EVDB_TRY(set_device(s)); return ensure_capacity(s, need); }
uint64_t store_device_bytes(const evdb_store *s) { return device_bytes(s); }

}  // namespace evdb

using namespace evdb;

// ============================================================================
// C ABI
// ============================================================================
extern "C" {

int evdb_abi_version(void) { return EVDB_ABI_VERSION; }

const char *evdb_strerror(int code) {
    switch (code) {
        case EVDB_OK: return "ok";
        case EVDB_E_DIM_MISMATCH: return "dimension_mismatch";
        case EVDB_E_BAD_VECTOR: return "invalid_vector_format";
        case EVDB_E_OOM: return "out_of_device_memory";
        case EVDB_E_CUDA: return "cuda_error";
        case EVDB_E_NCCL: return "nccl_error";
        case EVDB_E_BAD_ARG: return "badarg";
        case EVDB_E_NO_DEVICE: return "no_sm100_device";
        case EVDB_E_UNSUPPORTED: return "unsupported";
        case EVDB_E_BADARITH: return "badarith";
        default: return "unknown_error";
    }
}

const char *evdb_last_cuda_error(void) { return g_cuda_err; }

static int check_device(int dev) {
    // cudaGetDeviceProperties costs milliseconds: probe each ordinal once per process
    static int verdict[64];
    static bool probed[64];
    if (dev >= 0 && dev < 64 && probed[dev]) return verdict[dev];
    int rc = EVDB_OK;
    int n = 0;
    cudaDeviceProp p;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); rc = EVDB_E_NO_DEVICE; }
    else if (dev < 0 || dev >= n) rc = EVDB_E_NO_DEVICE;
    else if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); rc = EVDB_E_NO_DEVICE; }
    else if (p.major != 10) rc = EVDB_E_NO_DEVICE;  // sm_100a SASS only: no fallback
    if (dev >= 0 && dev < 64 && rc == EVDB_OK) { verdict[dev] = rc; probed[dev] = true; }
    return rc;
}

}  // extern "C"
namespace evdb { int check_device_public(int dev) { return check_device(dev); } }
extern "C" {

int evdb_init(const int *devices, int n_dev) {
    if (!devices || n_dev <= 0) return check_device(0);
    for (int i = 0; i < n_dev; ++i) EVDB_TRY(check_device(devices[i]));
    return EVDB_OK;
}

int evdb_store_create(const evdb_opts *opts, evdb_store **out) {
    if (!opts || !out) return EVDB_E_BAD_ARG;
    *out = nullptr;
    if (opts->dtype < EVDB_F32 || opts->dtype > EVDB_U4 || opts->dim < 0) return EVDB_E_BAD_ARG;
    if (opts->n_shards < 0 || opts->n_shards > EVDB_MAX_SHARDS) return EVDB_E_BAD_ARG;
    if (opts->n_shards > 1) {
        // ONE handle, n_shards devices: the owner carries count / dimension / dtype for validation,
        // the rows live in the shard stores behind s->multi (mstore.cu)
        for (int i = 0; i < opts->n_shards; ++i) EVDB_TRY(check_device(opts->devices[i]));
        evdb_store *s = new (std::nothrow) evdb_store();
        if (!s) return EVDB_E_OOM;
        s->device = opts->devices[0];
        s->dtype = opts->dtype;
        s->gemm_shadow = opts->gemm_shadow;
        const int rc = mstore_create(s, opts);
        if (rc != EVDB_OK) { delete s; return rc; }
        *out = s;
        return EVDB_OK;
    }
    EVDB_TRY(check_device(opts->device));
    evdb_store *s = new (std::nothrow) evdb_store();
    if (!s) return EVDB_E_OOM;
    s->device = opts->device;
    s->dtype = opts->dtype;
    s->gemm_shadow = opts->gemm_shadow;
    s->capacity_hint = opts->capacity_hint;
    int rc = EVDB_OK;
    do {
        if (cudaSetDevice(s->device) != cudaSuccess) { rc = EVDB_E_CUDA; break; }
        cudaDeviceProp p;
        cudaGetDeviceProperties(&p, s->device);
        s->sm_count = p.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = EVDB_E_CUDA; break; }
        if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess ||
            cudaEventCreate(&s->ev2) != cudaSuccess || cudaEventCreate(&s->ev3) != cudaSuccess) { rc = EVDB_E_CUDA; break; }
        if (opts->dim > 0) {
            set_dim(s, opts->dim);
            if (opts->capacity_hint) rc = ensure_capacity(s, opts->capacity_hint);
        }
    } while (0);
    if (rc != EVDB_OK) { evdb_store_destroy(s); return rc; }
    *out = s;
    return EVDB_OK;
}

void evdb_store_destroy(evdb_store *s) {
    if (!s) return;
    if (s->multi) { mstore_destroy(s->multi); s->multi = nullptr; delete s; return; }
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->rows); cudaFree(s->norm64); cudaFree(s->inv_norm); cudaFree(s->norm_sq);
    cudaFree(s->d_arrive);
    cudaFree(s->qcoef); cudaFree(s->qms64); cudaFree(s->shadow); cudaFree(s->shadow_l2); cudaFree(s->l2_tail); cudaFree(s->d_scalar);
    cudaFree(s->w_q64); cudaFree(s->w_q32); cudaFree(s->w_qdig); cudaFree(s->w_qh); cudaFree(s->w_seed); cudaFree(s->w_qstat); cudaFree(s->w_qeps);
    cudaFree(s->w_partial); cudaFree(s->w_ids); cudaFree(s->w_dists); cudaFree(s->w_counts);
    cudaFree(s->w_tmp); cudaFree(s->w_shard);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    if (s->gexec) cudaGraphExecDestroy(s->gexec);
    if (s->h_gq) cudaFreeHost(s->h_gq);
    if (s->h_ring) cudaFreeHost(s->h_ring);
    cudaFree(s->d_ring);
    for (int i = 0; i < 2; ++i) if (s->ring_ev[i]) cudaEventDestroy(s->ring_ev[i]);
    if (s->ev_ing) cudaEventDestroy(s->ev_ing);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev2) cudaEventDestroy(s->ev2);
    if (s->ev3) cudaEventDestroy(s->ev3);
    if (s->prof_ev) {
        for (int i = 0; i < 2 * kProfMax; ++i) if (s->prof_ev[i]) cudaEventDestroy(s->prof_ev[i]);
        free(s->prof_ev);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
}

int evdb_store_stats(evdb_store *s, evdb_stats *out) {
    if (!s || !out) return EVDB_E_BAD_ARG;
    memset(out, 0, sizeof(*out));
    if (s->multi) return m_stats(s->multi, out);
    out->count = s->count;
    out->dimension = s->dim;
    out->dtype = s->dtype;
    out->device = s->device;
    out->last_plan = s->last_plan;
    out->capacity = s->capacity;
    out->device_bytes = device_bytes(s);
    out->searches = s->n_searches;
    out->rows_scanned = s->n_rows_scanned;
    out->escalations = s->n_escalations;
    out->kernel_launches = s->n_launches;
    out->last_search_ms = s->last_search_ms;
    out->n_shards = 1;
    out->gemm_disabled = s->gemm_oom;
    out->shadow_bytes = (s->shadow ? s->capacity * (uint64_t)s->spitch * 2 : 0) +
                        (s->shadow_l2 ? s->l2_cap * (uint64_t)(s->l2_pitch + 16) * 2 : 0);
    out->upserts = s->n_upserts;
    out->deletes = s->n_deletes;
    out->last_h2d_ms = s->last_h2d_ms;
    out->last_device_ms = s->last_device_ms;
    out->last_d2h_ms = s->last_d2h_ms;
    return EVDB_OK;
}

int evdb_store_flush(evdb_store *s) {
    if (!s) return EVDB_E_BAD_ARG;
    if (s->multi) return m_flush(s->multi);
    EVDB_TRY(set_device(s));
    EVDB_CUDA(cudaStreamSynchronize(s->stream));
    EVDB_CUDA(cudaGetLastError());
    s->ingest_pending = 0;
    s->ring_busy[0] = s->ring_busy[1] = 0;
    s->ring_off = 0;
    return EVDB_OK;
}

int evdb_store_set_plan(evdb_store *s, int plan) {
    if (!s || plan < EVDB_PLAN_AUTO || plan > EVDB_PLAN_EXACT) return EVDB_E_BAD_ARG;
    if (s->multi) return m_set_plan(s->multi, plan);
    s->plan = plan;
    s->graph_epoch++;
    return EVDB_OK;
}

int evdb_store_profile(evdb_store *s, int enable) {
    if (!s) return EVDB_E_BAD_ARG;
    if (s->multi) return m_profile(s->multi, enable);
    EVDB_TRY(set_device(s));
    if (enable && !s->prof_ev) {
        s->prof_ev = (cudaEvent_t *)calloc(2 * kProfMax, sizeof(cudaEvent_t));
        if (!s->prof_ev) return EVDB_E_OOM;
        for (int i = 0; i < 2 * kProfMax; ++i) EVDB_CUDA(cudaEventCreate(&s->prof_ev[i]));
    }
    s->prof_on = enable ? 1 : 0;
    s->prof_n = 0;
    return EVDB_OK;
}

int evdb_store_profile_read(evdb_store *s, int32_t *n_samples, double *total_ms) {
    if (!s || !n_samples || !total_ms) return EVDB_E_BAD_ARG;
    *n_samples = 0;
    *total_ms = 0.0;
    if (s->multi) return m_profile_read(s->multi, n_samples, total_ms);
    if (!s->prof_ev || s->prof_n == 0) return EVDB_OK;
    EVDB_TRY(set_device(s));
    EVDB_CUDA(cudaEventSynchronize(s->prof_ev[2 * s->prof_n - 1]));
    double tot = 0.0;
    for (int i = 0; i < s->prof_n; ++i) {
        float ms = 0.f;
        EVDB_CUDA(cudaEventElapsedTime(&ms, s->prof_ev[2 * i], s->prof_ev[2 * i + 1]));
        tot += ms;
    }
    *n_samples = s->prof_n;
    *total_ms = tot;
    s->prof_n = 0;
    return EVDB_OK;
}

int evdb_store_upsert_f64(evdb_store *s, uint32_t slot, const double *vec, int d) {
    return upsert_any<double>(s, slot, vec, d, true);
}
int evdb_store_upsert_f32(evdb_store *s, uint32_t slot, const float *vec, int d) {
    return upsert_any<float>(s, slot, vec, d, false);
}
int evdb_store_bulk_load_f32(evdb_store *s, const float *rows, uint64_t n, int d) {
    return bulk_any<float>(s, rows, n, d, false);
}
int evdb_store_bulk_load_f64(evdb_store *s, const double *rows, uint64_t n, int d) {
    return bulk_any<double>(s, rows, n, d, true);
}

int evdb_store_append_f64(evdb_store *s, const double *rows, uint64_t n, int d, uint64_t *first_slot) {
    return append_any<double>(s, rows, n, d, true, first_slot);
}
int evdb_store_append_f32(evdb_store *s, const float *rows, uint64_t n, int d, uint64_t *first_slot) {
    return append_any<float>(s, rows, n, d, false, first_slot);
}

int evdb_store_bulk_load_codes(evdb_store *s, const uint8_t *codes, const double *mins,
                               const double *scales, uint64_t n, int d) {
    if (!s || !is_quant(s)) return EVDB_E_BAD_ARG;
    if (n > 0 && (!codes || !mins || !scales)) return EVDB_E_BAD_ARG;
    if (n > 0xFFFFFFF0ull) return EVDB_E_BAD_ARG;
    if (n > 0) {
        EVDB_TRY(check_dim(s, d));
        if (!rows_finite(mins, n) || !rows_finite(scales, n)) return EVDB_E_BAD_VECTOR;
    }
    if (s->multi) return m_bulk_codes(s->multi, codes, mins, scales, n, d);
    return store_load_codes(s, codes, s->dtype == EVDB_U8 ? (size_t)d : (size_t)(d + 1) / 2, mins, scales, 1, n, d);
}

int evdb_store_delete(evdb_store *s, uint32_t slot, int64_t *moved_from) {
    if (!s) return EVDB_E_BAD_ARG;
    if (moved_from) *moved_from = -1;
    if ((uint64_t)slot >= s->count) return EVDB_E_BAD_ARG;
    if (s->multi) return m_delete(s->multi, slot, moved_from);
    EVDB_TRY(set_device(s));
    uint64_t last = s->count - 1;
    cudaStream_t st = s->stream;
    if ((uint64_t)slot != last) {
        // the euclidean operand column follows only where both rows are current; otherwise the hole is rebuilt at the next search
        const bool l2 = s->shadow_l2 && (uint64_t)slot < s->l2_valid && last < s->l2_valid;
        if (s->shadow_l2 && (uint64_t)slot < s->l2_valid && !l2) s->l2_valid = slot;
        delete_swap_kernel<<<1, 256, 0, st>>>(s->rows, s->row_bytes, s->norm64, s->inv_norm, s->norm_sq,
                                             is_quant(s) ? s->qcoef : nullptr, s->qms64, s->shadow, s->spitch,
                                             l2 ? s->shadow_l2 : nullptr, s->l2_tail, s->l2_pitch, (uint64_t)slot, last);
        EVDB_CUDA(cudaGetLastError());
        s->n_launches++;
        if (!s->ev_ing) EVDB_CUDA(cudaEventCreateWithFlags(&s->ev_ing, cudaEventDisableTiming));
        s->ingest_pending = 1;     // enqueued, not waited for: ordered before any later work on the store's stream
        if (moved_from) *moved_from = (int64_t)last;
    }
    s->count = last;
    s->graph_epoch++;
    s->n_deletes++;
    if (s->shadow_valid > s->count) s->shadow_valid = s->count;
    if (s->l2_valid > s->count) s->l2_valid = s->count;
    return EVDB_OK;
}

int evdb_store_get_f64(evdb_store *s, uint32_t slot, double *out, int d) {
    if (!s || !out) return EVDB_E_BAD_ARG;
    if ((uint64_t)slot >= s->count) return EVDB_E_BAD_ARG;
    if (d != s->dim) return EVDB_E_DIM_MISMATCH;
    if (s->multi) return m_get_f64(s->multi, slot, out, d);
    EVDB_TRY(set_device(s));
    cudaStream_t st = s->stream;
    EVDB_TRY(ensure_bytes(&s->h_pin, &s->h_pin_cap, s->row_bytes + sizeof(double2) + (size_t)d * sizeof(double), true));
    uint8_t *h = (uint8_t *)s->h_pin;
    if (is_quant(s)) {
        EVDB_TRY(ensure_bytes(&s->w_tmp, &s->w_tmp_cap, (size_t)d * sizeof(double)));
        EVDB_TRY(launch_dequantize_rows(s->dtype, s->rows + (size_t)slot * s->row_bytes, s->row_bytes,
                                        s->qms64 + slot, 1, d, (double *)s->w_tmp, st));
        EVDB_CUDA(cudaMemcpyAsync(out, s->w_tmp, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, st));
        EVDB_CUDA(cudaStreamSynchronize(st));
        return EVDB_OK;
    }
    EVDB_CUDA(cudaMemcpyAsync(h, s->rows + (size_t)slot * s->row_bytes, s->row_bytes, cudaMemcpyDeviceToHost, st));
    EVDB_CUDA(cudaStreamSynchronize(st));
    if (s->dtype == EVDB_F32) {
        const float *f = (const float *)h;
        for (int i = 0; i < d; ++i) out[i] = (double)f[i];
    } else {
        const uint16_t *u = (const uint16_t *)h;
        for (int i = 0; i < d; ++i) {
            union { uint32_t u; float f; } cv;
            cv.u = (uint32_t)u[i] << 16;
            out[i] = (double)cv.f;
        }
    }
    return EVDB_OK;
}

int evdb_store_get_codes(evdb_store *s, uint32_t slot, uint8_t *codes, double *mn, double *scale) {
    if (!s || !codes || !is_quant(s)) return EVDB_E_BAD_ARG;
    if ((uint64_t)slot >= s->count) return EVDB_E_BAD_ARG;
    if (s->multi) return m_get_codes(s->multi, slot, codes, mn, scale);
    EVDB_TRY(set_device(s));
    size_t nb = s->dtype == EVDB_U8 ? (size_t)s->dim : (size_t)(s->dim + 1) / 2;
    double2 ms;
    EVDB_CUDA(cudaMemcpyAsync(codes, s->rows + (size_t)slot * s->row_bytes, nb, cudaMemcpyDeviceToHost, s->stream));
    EVDB_CUDA(cudaMemcpyAsync(&ms, s->qms64 + slot, sizeof(ms), cudaMemcpyDeviceToHost, s->stream));
    EVDB_CUDA(cudaStreamSynchronize(s->stream));
    if (mn) *mn = ms.x;
    if (scale) *scale = ms.y;
    return EVDB_OK;
}

int evdb_store_fill_synthetic(evdb_store *s, uint64_t seed, uint64_t row0, uint64_t n, int d) {
    if (!s) return EVDB_E_BAD_ARG;
    EVDB_TRY(check_dim(s, d));
    if (n > 0xFFFFFFF0ull) return EVDB_E_BAD_ARG;
    if (s->multi) return m_fill_synthetic(s->multi, seed, row0, n, d);
    return store_fill_synthetic(s, seed, row0, 1, n, d);
}
}  // extern "C"
namespace evdb {
// local row i = global row row0 + i * row_stride of the synthetic corpus
int store_fill_synthetic(evdb_store *s, uint64_t seed, uint64_t row0, uint64_t row_stride, uint64_t n, int d) {
    fix_dim(s, d);
    EVDB_TRY(set_device(s));
    EVDB_TRY(ensure_capacity(s, n));
    s->count = 0;
    s->shadow_valid = 0;
    s->l2_valid = 0;
    s->max_norm_dirty = 1;
    s->graph_epoch++;
    EVDB_TRY(launch_fill_synthetic(s, seed, row0, row_stride, n, s->stream));
    EVDB_TRY(launch_finalize_rows(s, 0, n, s->stream));
    EVDB_CUDA(cudaStreamSynchronize(s->stream));
    s->count = n;
    return EVDB_OK;
}
}  // namespace evdb
extern "C" {

int evdb_store_search_f64(evdb_store *s, const double *queries, int B, int d, int k, int metric,
                          uint32_t *out_slots, double *out_dists, int32_t *out_counts) {
    return search_host(s, queries, true, B, d, k, metric, out_slots, out_dists, out_counts);
}
int evdb_store_search_f32(evdb_store *s, const float *queries, int B, int d, int k, int metric,
                          uint32_t *out_slots, double *out_dists, int32_t *out_counts) {
    return search_host(s, queries, false, B, d, k, metric, out_slots, out_dists, out_counts);
}

int evdb_store_search_dev(evdb_store *s, const void *d_queries_f64, int B, int d, int k, int metric,
                          uint64_t slot_base, void *d_out_ids_u64, void *d_out_dists_f64,
                          void *d_out_counts_i32, void *d_out_flags_i32, void *stream) {
    if (s && s->multi) return EVDB_E_UNSUPPORTED;   // one handle, N devices: host entry points only
    if (!s || !d_queries_f64 || B <= 0 || k <= 0 || metric < 0 || metric > 2) return EVDB_E_BAD_ARG;
    if (!d_out_ids_u64 || !d_out_dists_f64 || !d_out_counts_i32) return EVDB_E_BAD_ARG;
    if (s->dim == 0 || s->count == 0) return EVDB_E_BAD_ARG;
    if (d != s->dim) return EVDB_E_DIM_MISMATCH;
    EVDB_TRY(set_device(s));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    EVDB_TRY(order_after_ingest(s, st));
    EVDB_TRY(search_core(s, (const double *)d_queries_f64, B, k, k, metric, 0, EVDB_PLAN_AUTO,
                         slot_base, (uint64_t *)d_out_ids_u64, (double *)d_out_dists_f64,
                         (int32_t *)d_out_counts_i32, (int32_t *)d_out_flags_i32, st));
    s->n_searches += (uint64_t)B;
    return EVDB_OK;
}

int evdb_store_search_dev_ex(evdb_store *s, const void *d_queries_f64, int B, int d, int k, int metric,
                             const evdb_search_opts *o, void *d_out_ids_u64, void *d_out_dists_f64,
                             void *d_out_counts_i32, void *d_out_flags_i32, void *stream) {
    if (s && s->multi) return EVDB_E_UNSUPPORTED;   // one handle, N devices: host entry points only
    if (!s || !o || !d_queries_f64 || B <= 0 || k <= 0 || metric < 0 || metric > 2) return EVDB_E_BAD_ARG;
    if (!d_out_ids_u64 || !d_out_dists_f64 || !d_out_counts_i32) return EVDB_E_BAD_ARG;
    if (o->plan < EVDB_PLAN_AUTO || o->plan > EVDB_PLAN_EXACT || o->kp_min < 0) return EVDB_E_BAD_ARG;
    if (s->dim == 0 || s->count == 0) return EVDB_E_BAD_ARG;
    if (d != s->dim) return EVDB_E_DIM_MISMATCH;
    EVDB_TRY(set_device(s));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    EVDB_TRY(order_after_ingest(s, st));
    const uint64_t keep_mul = s->slot_mul;
    s->slot_mul = o->slot_stride > 1 ? o->slot_stride : 1;
    const int rc = search_core(s, (const double *)d_queries_f64, B, k, k, metric, o->kp_min, o->plan, o->slot_base,
                               (uint64_t *)d_out_ids_u64, (double *)d_out_dists_f64, (int32_t *)d_out_counts_i32,
                               (int32_t *)d_out_flags_i32, st);
    s->slot_mul = keep_mul;
    EVDB_TRY(rc);
    s->n_searches += (uint64_t)B;
    return EVDB_OK;
}

int evdb_merge_topk_dev(int device, const void *d_ids_u64, const void *d_dists_f64,
                        const void *d_counts_i32, int G, int B, int k, void *d_out_ids_u64,
                        void *d_out_dists_f64, void *d_out_counts_i32, void *stream) {
    if (!d_ids_u64 || !d_dists_f64 || !d_counts_i32 || !d_out_ids_u64 || !d_out_dists_f64 || !d_out_counts_i32)
        return EVDB_E_BAD_ARG;
    EVDB_TRY(check_device(device));
    EVDB_CUDA(cudaSetDevice(device));
    return launch_merge_topk((const uint64_t *)d_ids_u64, (const double *)d_dists_f64,
                             (const int32_t *)d_counts_i32, G, B, k, (uint64_t *)d_out_ids_u64,
                             (double *)d_out_dists_f64, (int32_t *)d_out_counts_i32, (cudaStream_t)stream);
}

int evdb_merge_topk_packed_dev(int device, const void *d_blobs, int G, int B, int k, void *d_out_blob,
                               void *stream) {
    if (!d_blobs || !d_out_blob) return EVDB_E_BAD_ARG;
    EVDB_TRY(check_device(device));
    EVDB_CUDA(cudaSetDevice(device));
    return launch_merge_topk_packed((const uint64_t *)d_blobs, 2 * (size_t)B * k + (size_t)B, G, B, k,
                                    (uint64_t *)d_out_blob, (cudaStream_t)stream);
}

// ---- row-sharded GEMM batches in two phases (select.cu: windows travel, owners re-rank) ----
struct ShardLayout { size_t win_words, e_off, g_off, m_off, total; int KP, kk; };
static ShardLayout shard_layout(int B, int k, uint64_t n_total) {
    ShardLayout l;
    l.kk = (uint64_t)k < n_total ? k : (int)n_total;
    l.KP = gemm_kp(choose_kp(l.kk, 0));
    l.win_words = (size_t)B * l.KP + (size_t)B;
    l.e_off = l.win_words * 8;
    l.g_off = l.e_off + (size_t)B * l.KP * 8;
    l.m_off = l.g_off + (size_t)B * l.KP * 8;
    l.total = l.m_off + shard_gmeta_bytes(B);
    return l;
}

int evdb_store_search_sharded_phase1(evdb_store *s, evdb_exchange *xw, const void *d_queries_f64, int B, int d, int k,
                                     int metric, uint64_t slot_base, uint64_t n_total, void *stream) {
    if (s && s->multi) return EVDB_E_UNSUPPORTED;
    if (!s || !xw || !d_queries_f64 || B <= 0 || k <= 0) return EVDB_E_BAD_ARG;
    if (s->dim == 0 || s->count == 0) return EVDB_E_UNSUPPORTED;
    if (d != s->dim) return EVDB_E_DIM_MISMATCH;
    const ShardLayout l = shard_layout(B, k, n_total);
    if (B > gemm_max_batch() || !gemm_plan_supported(s, metric, B, l.KP) || (size_t)xw->world * l.KP > 2048 || xw->world > 32 ||   // the phase kernels map one lane per rank
       
        l.win_words > xw->max_words || n_total > 0xFFFFFFF0ull)
        return EVDB_E_UNSUPPORTED;
    EVDB_TRY(set_device(s));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    EVDB_TRY(order_after_ingest(s, st));
    EVDB_TRY(ensure_bytes(&s->w_shard, &s->w_shard_cap, l.total));
    int lists = 0;
    const float *eps_q = nullptr;
    RawCands raw;
    s->last_plan = EVDB_PLAN_GEMM;
    EVDB_TRY(launch_gemm_topk(s, (const double *)d_queries_f64, B, l.KP, metric, &lists, &eps_q, &raw, st));
    // the window kernel stores straight into every rank's mailbox and publishes the epoch itself
    EVDB_TRY(launch_shard_window(s, (const double *)d_queries_f64, &raw, lists, l.KP, B, l.kk, metric, eps_q, slot_base,
                                 exchange_begin_push(xw), st));
    s->n_launches++;
    s->n_searches += (uint64_t)B;
    s->n_rows_scanned += (uint64_t)B * s->count;
    return EVDB_OK;
}

int evdb_store_search_sharded_phase2(evdb_store *s, evdb_exchange *xw, evdb_exchange *xe, const void *d_queries_f64,
                                     int B, int k, int metric, uint64_t n_total, void *stream) {
    if (s && s->multi) return EVDB_E_UNSUPPORTED;
    if (!s || !xw || !xe || !d_queries_f64) return EVDB_E_BAD_ARG;
    const ShardLayout l = shard_layout(B, k, n_total);
    if ((size_t)B * l.KP > xe->max_words) return EVDB_E_UNSUPPORTED;
    EVDB_TRY(set_device(s));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    uint8_t *w = (uint8_t *)s->w_shard;
    const ExchangeView win = exchange_view(xw);
    EVDB_TRY(launch_shard_rerank(s, (const double *)d_queries_f64, B, l.KP, k, l.kk, metric, xw->rank, xw->world, n_total,
                                 win, exchange_begin_push(xe), (uint64_t *)(w + l.g_off), w + l.m_off, st));
    s->n_launches++;
    return EVDB_OK;
}

int evdb_store_search_sharded_phase3(evdb_store *s, evdb_exchange *xe, int B, int k, int metric, uint64_t n_total,
                                     void *d_out_blob, void *stream) {
    if (s && s->multi) return EVDB_E_UNSUPPORTED;
    if (!s || !xe || !d_out_blob) return EVDB_E_BAD_ARG;
    const ShardLayout l = shard_layout(B, k, n_total);
    EVDB_TRY(set_device(s));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    uint8_t *w = (uint8_t *)s->w_shard;
    return launch_shard_final(s, B, l.KP, k, l.kk, metric, xe->rank, xe->world, n_total, exchange_view(xe),
                              (const uint64_t *)(w + l.g_off), w + l.m_off, (uint64_t *)d_out_blob, st);
}

// ---- vector_utils pairwise functions ----------------------------------------------------------
int evdb_vector_utils_f64(int device, int op, const double *a, const double *b, uint64_t n, int d, double *out) {
    if (!a || !out || d <= 0 || op < EVDB_VU_COSINE_SIMILARITY || op > EVDB_VU_NORM) return EVDB_E_BAD_ARG;
    if (!b && op != EVDB_VU_NORM) return EVDB_E_BAD_ARG;
    if (n == 0) return EVDB_OK;
    if (!rows_finite(a, n * (uint64_t)d) || (b && !rows_finite(b, n * (uint64_t)d))) return EVDB_E_BAD_VECTOR;
    EVDB_TRY(check_device(device));
    EVDB_CUDA(cudaSetDevice(device));
    const size_t bytes = n * (size_t)d * sizeof(double);
    double *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
    int rc = EVDB_OK;
#define VU(expr) do { if ((expr) != cudaSuccess) { set_cuda_error(cudaGetLastError(), __FILE__, __LINE__); rc = EVDB_E_CUDA; goto done; } } while (0)
    VU(cudaMalloc((void **)&d_a, bytes));
    if (b) VU(cudaMalloc((void **)&d_b, bytes));
    VU(cudaMalloc((void **)&d_o, n * sizeof(double)));
    VU(cudaMemcpy(d_a, a, bytes, cudaMemcpyHostToDevice));
    if (b) VU(cudaMemcpy(d_b, b, bytes, cudaMemcpyHostToDevice));
    rc = launch_vector_utils(op, d_a, d_b, n, d, d_o, 0);
    if (rc != EVDB_OK) goto done;
    VU(cudaMemcpy(out, d_o, n * sizeof(double), cudaMemcpyDeviceToHost));
done:
#undef VU
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_o);
    return rc;
}

// ---- diagnostics: integer digit-plane sums of the quantized scan -------------------------------
int evdb_debug_quant_dots(evdb_store *s, const double *query, int d, const uint32_t *slots, int n, int64_t *out_sum,
                          int32_t *out_planes, int32_t *out_code_sum, int32_t *out_shift) {
    if (!s || s->multi || !is_quant(s) || !query || !slots || n <= 0 || !out_sum || !out_planes || !out_code_sum || !out_shift)
        return EVDB_E_BAD_ARG;
    if (d != s->dim) return EVDB_E_DIM_MISMATCH;
    for (int i = 0; i < n; ++i) if ((uint64_t)slots[i] >= s->count) return EVDB_E_BAD_ARG;
    EVDB_TRY(set_device(s));
    cudaStream_t st = s->stream;
    const size_t qb = round_up64((size_t)d * 8, 256), sb = round_up64((size_t)n * 4, 256), Sb = round_up64((size_t)n * 8, 256),
                 pb = round_up64((size_t)n * 12, 256);
    EVDB_TRY(ensure_bytes(&s->w_tmp, &s->w_tmp_cap, qb + 2 * sb + Sb + pb));
    uint8_t *w = (uint8_t *)s->w_tmp;
    double *d_q = (double *)w; uint32_t *d_slots = (uint32_t *)(w + qb); long long *d_S = (long long *)(w + qb + sb);
    int *d_pl = (int *)(w + qb + sb + Sb), *d_cs = (int *)(w + qb + sb + Sb + pb);
    EVDB_CUDA(cudaMemcpyAsync(d_q, query, (size_t)d * 8, cudaMemcpyHostToDevice, st));
    EVDB_CUDA(cudaMemcpyAsync(d_slots, slots, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    float fx = 0.f;
    EVDB_TRY(debug_quant_dots(s, d_q, d_slots, n, d_S, d_pl, d_cs, &fx, st));
    EVDB_CUDA(cudaMemcpy(out_sum, d_S, (size_t)n * 8, cudaMemcpyDeviceToHost));
    EVDB_CUDA(cudaMemcpy(out_planes, d_pl, (size_t)n * 12, cudaMemcpyDeviceToHost));
    EVDB_CUDA(cudaMemcpy(out_code_sum, d_cs, (size_t)n * 4, cudaMemcpyDeviceToHost));
    int e = 0;
    frexp((double)fx, &e);     // fx = 2^-shift = 0.5 * 2^(1-shift)
    *out_shift = 1 - e;
    return EVDB_OK;
}

// ---- standalone codecs ------------------------------------------------------
static int codec_quantize(int device, int dtype, const double *rows, uint64_t n, int d, uint8_t *codes,
                          double *mins, double *maxs, double *scales, uint8_t *ok) {
    if (!rows || !codes || !mins || !scales || d <= 0) return EVDB_E_BAD_ARG;
    if (n == 0) return EVDB_OK;
    EVDB_TRY(check_device(device));
    EVDB_CUDA(cudaSetDevice(device));
    size_t crb = dtype == EVDB_U8 ? (size_t)d : (size_t)(d + 1) / 2;
    double *d_rows = nullptr, *d_max = nullptr;
    uint8_t *d_codes = nullptr, *d_ok = nullptr;
    double2 *d_ms = nullptr;
    int rc = EVDB_OK;
    double2 *h_ms = (double2 *)malloc(n * sizeof(double2));
    if (!h_ms) return EVDB_E_OOM;
#define CQ(expr) do { if ((expr) != cudaSuccess) { set_cuda_error(cudaGetLastError(), __FILE__, __LINE__); rc = EVDB_E_CUDA; goto done; } } while (0)
    CQ(cudaMalloc((void **)&d_rows, n * (size_t)d * sizeof(double)));
    CQ(cudaMalloc((void **)&d_codes, n * crb));
    CQ(cudaMalloc((void **)&d_ms, n * sizeof(double2)));
    CQ(cudaMalloc((void **)&d_max, n * sizeof(double)));
    CQ(cudaMalloc((void **)&d_ok, n));
    CQ(cudaMemcpy(d_rows, rows, n * (size_t)d * sizeof(double), cudaMemcpyHostToDevice));
    rc = launch_quantize_rows(dtype, d_rows, nullptr, n, d, d_codes, crb, d_ms, d_max, d_ok, 0);
    if (rc != EVDB_OK) goto done;
    CQ(cudaMemcpy(codes, d_codes, n * crb, cudaMemcpyDeviceToHost));
    CQ(cudaMemcpy(h_ms, d_ms, n * sizeof(double2), cudaMemcpyDeviceToHost));
    if (maxs) CQ(cudaMemcpy(maxs, d_max, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (ok) CQ(cudaMemcpy(ok, d_ok, n, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < n; ++i) { mins[i] = h_ms[i].x; scales[i] = h_ms[i].y; }
done:
#undef CQ
    cudaFree(d_rows); cudaFree(d_codes); cudaFree(d_ms); cudaFree(d_max); cudaFree(d_ok);
    free(h_ms);
    return rc;
}

static int codec_dequantize(int device, int dtype, const uint8_t *codes, const double *mins,
                            const double *scales, uint64_t n, int d, double *out) {
    if (!codes || !mins || !scales || !out || d <= 0) return EVDB_E_BAD_ARG;
    if (n == 0) return EVDB_OK;
    EVDB_TRY(check_device(device));
    EVDB_CUDA(cudaSetDevice(device));
    size_t crb = dtype == EVDB_U8 ? (size_t)d : (size_t)(d + 1) / 2;
    uint8_t *d_codes = nullptr;
    double2 *d_ms = nullptr;
    double *d_out = nullptr;
    int rc = EVDB_OK;
    double2 *h_ms = (double2 *)malloc(n * sizeof(double2));
    if (!h_ms) return EVDB_E_OOM;
    for (uint64_t i = 0; i < n; ++i) { h_ms[i].x = mins[i]; h_ms[i].y = scales[i]; }
#define CQ(expr) do { if ((expr) != cudaSuccess) { set_cuda_error(cudaGetLastError(), __FILE__, __LINE__); rc = EVDB_E_CUDA; goto done; } } while (0)
    CQ(cudaMalloc((void **)&d_codes, n * crb));
    CQ(cudaMalloc((void **)&d_ms, n * sizeof(double2)));
    CQ(cudaMalloc((void **)&d_out, n * (size_t)d * sizeof(double)));
    CQ(cudaMemcpy(d_codes, codes, n * crb, cudaMemcpyHostToDevice));
    CQ(cudaMemcpy(d_ms, h_ms, n * sizeof(double2), cudaMemcpyHostToDevice));
    rc = launch_dequantize_rows(dtype, d_codes, crb, d_ms, n, d, d_out, 0);
    if (rc != EVDB_OK) goto done;
    CQ(cudaMemcpy(out, d_out, n * (size_t)d * sizeof(double), cudaMemcpyDeviceToHost));
done:
#undef CQ
    cudaFree(d_codes); cudaFree(d_ms); cudaFree(d_out);
    free(h_ms);
    return rc;
}

int evdb_quantize_8bit(int device, const double *rows, uint64_t n, int d, uint8_t *codes,
                       double *mins, double *maxs, double *scales, uint8_t *ok) {
    return codec_quantize(device, EVDB_U8, rows, n, d, codes, mins, maxs, scales, ok);
}
int evdb_quantize_4bit(int device, const double *rows, uint64_t n, int d, uint8_t *packed,
                       double *mins, double *maxs, double *scales, uint8_t *ok) {
    return codec_quantize(device, EVDB_U4, rows, n, d, packed, mins, maxs, scales, ok);
}
int evdb_dequantize_8bit(int device, const uint8_t *codes, const double *mins, const double *scales,
                         uint64_t n, int d, double *out) {
    return codec_dequantize(device, EVDB_U8, codes, mins, scales, n, d, out);
}
int evdb_dequantize_4bit(int device, const uint8_t *packed, const double *mins, const double *scales,
                         uint64_t n, int d, double *out) {
    return codec_dequantize(device, EVDB_U4, packed, mins, scales, n, d, out);
}

}  // extern "C"
