// mstore.cu -- ONE store behind ONE handle on several GPUs of one process.
//
// The reference runs one gen_server per store inside one OS process (the BEAM): reference
// src/vector_store.erl:38-57 (start_link / init), :113-164 (insert, search, delete as synchronous
// calls), src/vector_store_sup.erl:24-31.  A NIF can therefore only drive several GPUs if a single
// C-ABI handle does; this file is that handle (evdb_opts.n_shards > 1).
//
// Layout: global slot g lives on shard g % S at local slot g / S.  An append (slot == count) lands
// on the shortest shard, a swap-with-last delete removes the last row of the longest one, so every
// shard stays dense and within one row of the others without any rebalancing.
//
// Search: the caller's query batch is cut into S slices; shard j validates and copies ITS slice over
// its own PCIe link, then stores it into every peer's query buffer over NVLink and raises a flag
// (bcast_slice_kernel) -- the batch crosses each host link once, 1/S of it.  Every shard then runs
// the same device path a single store runs (search_core) or, for tcgen05 batches, the two-phase
// sharded search (approximate windows travel, owners re-rank in fp64: store.cu
// evdb_store_search_sharded_phase1/2/3), candidates meet in shard 0 through the peer-memory
// exchange (exchange.cu), and only shard 0 copies the merged result back.  One host thread per
// device enqueues its shard's kernels; nothing but shard 0's final stream wait blocks.
// Queries whose window could not be proven complete are re-issued on every shard (256-key scan
// window, then the exhaustive fp64 plan): no unproven result leaves the handle.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>

#include "internal.h"

namespace evdb {

constexpr int kMS = EVDB_MAX_SHARDS;

struct BcastArgs {
    double *dst[kMS];                  // every shard's query buffer
    unsigned long long *flags[kMS];    // every shard's flag array [S]
    unsigned int *counter;             // local block counter (zero between uses)
    const double *src;                 // my slice (already in my own query buffer)
    size_t off, words;                 // slice position / length in doubles
    unsigned long long epoch;
    int S, rank;
};

// my slice -> the same offset of every peer's query buffer; the last block raises my flag everywhere
__global__ void __launch_bounds__(256) bcast_slice_kernel(const BcastArgs a) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.words; i += (size_t)gridDim.x * blockDim.x) {
        const double v = a.src[i];
        for (int p = 0; p < a.S; ++p)
            if (p != a.rank) a.dst[p][a.off + i] = v;
    }
    __threadfence_system();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < a.S) *reinterpret_cast<volatile unsigned long long *>(a.flags[threadIdx.x] + a.rank) = a.epoch;
        if (threadIdx.x == 0) *a.counter = 0;
    }
}

// one warp: all S slices of this epoch have landed in my query buffer (bounded: a dead peer traps)
__global__ void wait_slices_kernel(const unsigned long long *flags, unsigned long long epoch, int S) {
    if ((int)threadIdx.x < S) {
        const volatile unsigned long long *f = flags + threadIdx.x;
        const long long t0 = clock64();
        while (*f < epoch)
            if (clock64() - t0 > 8000000000ll) __trap();
    }
    __syncwarp();
    __threadfence_system();
}

__global__ void widen_slice_kernel(const float *__restrict__ in, double *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (double)in[i];
}

struct MJob {
    int kind = 0;                 // 1 = search, 2 = put rows
    // search
    const void *queries = nullptr;
    bool is_f64 = true;
    int B = 0, d = 0, k = 0, metric = 0, plan = EVDB_PLAN_AUTO, kp_min = 0;
    bool two_phase = false;
    uint64_t n_total = 0;
    // put
    const void *rows = nullptr;
    uint64_t first = 0, n = 0;
};

struct MShard {
    evdb_store *st = nullptr;
    int dev = 0;
    double *qbuf = nullptr;  size_t qbuf_cap = 0;      // [B][dim] fp64: the whole batch
    void *stage_dev = nullptr; size_t stage_dev_cap = 0; // fp32 slice before widening
    void *h_stage = nullptr; size_t h_stage_cap = 0;   // pinned: my slice of the caller's queries
    uint64_t *blob = nullptr; size_t blob_cap = 0;     // local packed result
    uint64_t *merged = nullptr; size_t merged_cap = 0; // merged packed result (shard 0 reads it back)
    unsigned long long *qflags = nullptr;              // [kMS] epochs of the query broadcast + a block counter behind them
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::thread th;
    int rc = EVDB_OK;
    bool bad_query = false;
};

struct MStore {
    evdb_store *owner = nullptr;
    int S = 0;
    bool serial = false;          // several shards share a device: one stream, phases enqueued in lock-step
    MShard sh[kMS];
    evdb_exchange *xb[kMS] = {nullptr};  uint64_t xb_words = 0;   // packed results
    evdb_exchange *xw[kMS] = {nullptr};  uint64_t xw_words = 0;   // two-phase: windows
    evdb_exchange *xe[kMS] = {nullptr};  uint64_t xe_words = 0;   // two-phase: exact distances
    unsigned long long qepoch = 0;
    void *h_res = nullptr; size_t h_res_cap = 0;       // pinned: merged result on the host
    void *h_esc = nullptr; size_t h_esc_cap = 0;       // host: flagged queries packed for the ladder
    // worker threads
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t seq = 0;
    std::atomic<uint64_t> seq_pub{0};   // == seq, readable without the mutex (the workers' spin phase)
    std::atomic<int> ndone_pub{0};
    int ndone = 0;
    bool quit = false;
    MJob job;
    // counters
    uint64_t n_searches = 0, n_escalations = 0, n_upserts = 0, n_deletes = 0;
    double last_ms = 0.0, last_h2d_ms = 0.0, last_dev_ms = 0.0, last_d2h_ms = 0.0;
    int last_plan = 0;
};

static inline uint64_t shard_count(uint64_t n, int S, int j) { return n > (uint64_t)j ? (n - j + S - 1) / S : 0; }

static int set_dev(int dev) {
    EVDB_CUDA(cudaSetDevice(dev));
    return EVDB_OK;
}

// ---- exchanges (re)built by the caller thread when a search needs larger mailboxes ----------------
static void destroy_exchanges(MStore *m, evdb_exchange **xs) {
    for (int j = 0; j < m->S; ++j) { evdb_exchange_destroy(xs[j]); xs[j] = nullptr; }
}
static int ensure_exchanges(MStore *m, evdb_exchange **xs, uint64_t *cur, uint64_t words) {
    if (words <= *cur && xs[0]) return EVDB_OK;
    uint64_t cap = 4096;
    while (cap < words) cap *= 2;
    for (int j = 0; j < m->S; ++j) {       // nothing may still be using the old mailboxes
        EVDB_TRY(set_dev(m->sh[j].dev));
        EVDB_CUDA(cudaStreamSynchronize(m->sh[j].st->stream));
    }
    destroy_exchanges(m, xs);
    *cur = 0;
    const void *boxes[kMS];
    for (int j = 0; j < m->S; ++j) {
        EVDB_TRY(evdb_exchange_create(m->sh[j].dev, j, m->S, cap, &xs[j], nullptr));
        boxes[j] = evdb_exchange_mailbox(xs[j]);
    }
    for (int j = 0; j < m->S; ++j) EVDB_TRY(evdb_exchange_connect_ptrs(xs[j], boxes));
    *cur = cap;
    return EVDB_OK;
}

static cudaStream_t shard_stream(MStore *m, int j) { return m->serial ? m->sh[0].st->stream : m->sh[j].st->stream; }

// ---- the steps of one search on shard j (threaded mode: its worker runs them back to back; serial
// mode: the caller runs step by step over all shards, so no kernel ever waits for one enqueued later)
static int step_stage(MStore *m, int j, const MJob &jb) {
    MShard &h = m->sh[j];
    EVDB_TRY(set_dev(h.dev));
    cudaStream_t st = shard_stream(m, j);
    if (h.st->ingest_pending && st != h.st->stream && h.st->ev_ing) {   // serial mode: the shard's ingest ran on its own stream
        EVDB_CUDA(cudaEventRecord(h.st->ev_ing, h.st->stream));
        EVDB_CUDA(cudaStreamWaitEvent(st, h.st->ev_ing, 0));
    }
    const int per = (jb.B + m->S - 1) / m->S;
    const int b0 = j * per < jb.B ? j * per : jb.B;
    const int nb = jb.B - b0 < per ? jb.B - b0 : per;
    const size_t nq = (size_t)nb * jb.d;
    h.bad_query = false;
    if (j == 0) EVDB_CUDA(cudaEventRecord(h.ev[0], st));
    if (nq > 0) {
        const size_t esz = jb.is_f64 ? 8 : 4;
        // validate_vector/2 (lists:all(is_number)) and the copy into pinned memory in one pass over my slice
        EVDB_TRY(ensure_bytes(&h.h_stage, &h.h_stage_cap, nq * esz, true));
        // (exponent-field tests on the raw words: the loop vectorises, a per-element isfinite() does not)
        const uint8_t *srcb = (const uint8_t *)jb.queries + (size_t)b0 * jb.d * esz;
        memcpy(h.h_stage, srcb, nq * esz);
        uint32_t bad = 0;
        if (jb.is_f64) {
            const uint64_t *w = (const uint64_t *)h.h_stage;
            for (size_t i = 0; i < nq; ++i) bad |= (uint32_t)(((w[i] >> 52) & 0x7FFu) == 0x7FFu);
        } else {
            const uint32_t *w = (const uint32_t *)h.h_stage;
            for (size_t i = 0; i < nq; ++i) bad |= (uint32_t)(((w[i] >> 23) & 0xFFu) == 0xFFu);
        }
        // a bad slice still goes through the protocol (peers wait for it) as zeros; the caller discards the result
        if (bad) memset(h.h_stage, 0, nq * esz);
        h.bad_query = bad != 0;
        double *mine = h.qbuf + (size_t)b0 * jb.d;
        if (jb.is_f64) {
            EVDB_CUDA(cudaMemcpyAsync(mine, h.h_stage, nq * 8, cudaMemcpyHostToDevice, st));
        } else {
            EVDB_TRY(ensure_bytes(&h.stage_dev, &h.stage_dev_cap, nq * 4));
            EVDB_CUDA(cudaMemcpyAsync(h.stage_dev, h.h_stage, nq * 4, cudaMemcpyHostToDevice, st));
            const int grid = (int)((nq + 255) / 256 < 1024 ? (nq + 255) / 256 : 1024);
            widen_slice_kernel<<<grid, 256, 0, st>>>((const float *)h.stage_dev, mine, nq);
            EVDB_CUDA(cudaGetLastError());
        }
    }
    if (m->S > 1) {
        BcastArgs a;
        memset(&a, 0, sizeof(a));
        for (int p = 0; p < m->S; ++p) { a.dst[p] = m->sh[p].qbuf; a.flags[p] = m->sh[p].qflags; }
        a.counter = reinterpret_cast<unsigned int *>(h.qflags + kMS);
        a.src = h.qbuf + (size_t)b0 * jb.d;
        a.off = (size_t)b0 * jb.d;
        a.words = nq;
        a.epoch = m->qepoch;
        a.S = m->S;
        a.rank = j;
        int grid = (int)((nq + 255) / 256);
        if (grid > 64) grid = 64;
        if (grid < 1) grid = 1;
        bcast_slice_kernel<<<grid, 256, 0, st>>>(a);
        EVDB_CUDA(cudaGetLastError());
    }
    return EVDB_OK;
}

static int step_wait(MStore *m, int j, const MJob &) {
    MShard &h = m->sh[j];
    EVDB_TRY(set_dev(h.dev));
    cudaStream_t st = shard_stream(m, j);
    if (m->S > 1) {
        wait_slices_kernel<<<1, 32, 0, st>>>(h.qflags, m->qepoch, m->S);
        EVDB_CUDA(cudaGetLastError());
    }
    if (j == 0) EVDB_CUDA(cudaEventRecord(h.ev[1], st));
    return EVDB_OK;
}

static inline size_t blob_words(int B, int k) { return 2 * (size_t)B * k + (size_t)B; }

static int step_local(MStore *m, int j, const MJob &jb) {
    MShard &h = m->sh[j];
    EVDB_TRY(set_dev(h.dev));
    cudaStream_t st = shard_stream(m, j);
    if (jb.two_phase)
        return evdb_store_search_sharded_phase1(h.st, m->xw[j], h.qbuf, jb.B, jb.d, jb.k, jb.metric, (uint64_t)j,
                                                jb.n_total, st);
    const size_t nk = (size_t)jb.B * jb.k;
    uint64_t *ids = h.blob;
    double *dists = reinterpret_cast<double *>(h.blob + nk);
    int32_t *counts = reinterpret_cast<int32_t *>(h.blob + 2 * nk);
    if (h.st->count == 0) {
        EVDB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 2 * (size_t)jb.B, st));   // nothing here: zero counts and flags
    } else {
        EVDB_TRY(search_core(h.st, h.qbuf, jb.B, jb.k, jb.k, jb.metric, jb.kp_min, jb.plan, (uint64_t)j, ids, dists, counts,
                             counts + jb.B, st));
    }
    if (m->S > 1) EVDB_TRY(exchange_push_words(m->xb[j], h.blob, blob_words(jb.B, jb.k), st));
    return EVDB_OK;
}

static int step_phase2(MStore *m, int j, const MJob &jb) {
    if (!jb.two_phase) return EVDB_OK;
    MShard &h = m->sh[j];
    EVDB_TRY(set_dev(h.dev));
    return evdb_store_search_sharded_phase2(h.st, m->xw[j], m->xe[j], h.qbuf, jb.B, jb.k, jb.metric, jb.n_total,
                                            shard_stream(m, j));
}

// shard 0 only: the merged result, copied to pinned host memory
static int step_final(MStore *m, const MJob &jb) {
    MShard &h = m->sh[0];
    EVDB_TRY(set_dev(h.dev));
    cudaStream_t st = shard_stream(m, 0);
    const uint64_t *res = h.merged;
    if (jb.two_phase) {
        EVDB_TRY(evdb_store_search_sharded_phase3(h.st, m->xe[0], jb.B, jb.k, jb.metric, jb.n_total, h.merged, st));
    } else if (m->S > 1) {
        EVDB_TRY(evdb_exchange_merge(m->xb[0], jb.B, jb.k, h.merged, st));
    } else {
        res = h.blob;
    }
    EVDB_CUDA(cudaEventRecord(h.ev[2], st));
    EVDB_CUDA(cudaMemcpyAsync(m->h_res, res, blob_words(jb.B, jb.k) * 8, cudaMemcpyDeviceToHost, st));
    EVDB_CUDA(cudaEventRecord(h.ev[3], st));
    EVDB_CUDA(cudaStreamSynchronize(st));
    return EVDB_OK;
}

// two-phase bookkeeping on the non-final shards (phase 2 advanced xe's epoch on every shard; shard 0's
// phase 3 reads it).  Nothing to do: kept as a named step for symmetry of the serial schedule.

static int run_put(MStore *m, int j, const MJob &jb) {
    // rows first + i of the caller's array with (first + i) % S == j  ->  local slots from (first + i0) / S
    const int S = m->S;
    const uint64_t i0 = ((uint64_t)j + S - jb.first % S) % S;
    if (i0 >= jb.n) return EVDB_OK;
    const uint64_t nj = (jb.n - i0 + S - 1) / S;
    const size_t esz = jb.is_f64 ? 8 : 4;
    const uint8_t *src = (const uint8_t *)jb.rows + i0 * (size_t)jb.d * esz;
    return store_put_rows(m->sh[j].st, (jb.first + i0) / S, src, jb.is_f64, nj, (size_t)S * jb.d, jb.d);
}

static int run_search_steps(MStore *m, int j, const MJob &jb) {
    EVDB_TRY(step_stage(m, j, jb));
    EVDB_TRY(step_wait(m, j, jb));
    EVDB_TRY(step_local(m, j, jb));
    EVDB_TRY(step_phase2(m, j, jb));
    if (j == 0) EVDB_TRY(step_final(m, jb));
    return EVDB_OK;
}

static void worker_main(MStore *m, int j) {
    uint64_t seen = 0;
    cudaSetDevice(m->sh[j].dev);
    while (true) {
        MJob jb;
        {
            // searches come in bursts: watch the sequence number for a little while before sleeping on the
            // condition variable (a futex wake-up costs more than a lone query's device time allows)
            for (int spin = 0; spin < 20000 && m->seq_pub.load(std::memory_order_acquire) == seen && !m->quit; ++spin) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_go.wait(lk, [&] { return m->quit || m->seq != seen; });
            if (m->quit) return;
            seen = m->seq;
            jb = m->job;
        }
        int rc = EVDB_OK;
        if (jb.kind == 1) rc = run_search_steps(m, j, jb);
        else if (jb.kind == 2) rc = run_put(m, j, jb);
        {
            std::lock_guard<std::mutex> lk(m->mu);
            m->sh[j].rc = rc;
            ++m->ndone;
            m->ndone_pub.store(m->ndone, std::memory_order_release);
            if (m->ndone == m->S) m->cv_done.notify_one();
        }
    }
}

// run `jb` on every shard; returns the first error
static int dispatch(MStore *m, const MJob &jb) {
    if (m->serial) {
        if (jb.kind == 2) {
            for (int j = 0; j < m->S; ++j) EVDB_TRY(run_put(m, j, jb));
            return EVDB_OK;
        }
        for (int j = 0; j < m->S; ++j) EVDB_TRY(step_stage(m, j, jb));
        for (int j = 0; j < m->S; ++j) EVDB_TRY(step_wait(m, j, jb));
        for (int j = 0; j < m->S; ++j) EVDB_TRY(step_local(m, j, jb));
        for (int j = 0; j < m->S; ++j) EVDB_TRY(step_phase2(m, j, jb));
        return step_final(m, jb);
    }
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->job = jb;
        m->ndone = 0;
        m->ndone_pub.store(0, std::memory_order_relaxed);
        m->seq++;
        m->seq_pub.store(m->seq, std::memory_order_release);
    }
    m->cv_go.notify_all();
    for (int spin = 0; spin < 200000 && m->ndone_pub.load(std::memory_order_acquire) != m->S; ++spin) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    {
        std::unique_lock<std::mutex> lk(m->mu);
        m->cv_done.wait(lk, [&] { return m->ndone == m->S; });
    }
    for (int j = 0; j < m->S; ++j)
        if (m->sh[j].rc != EVDB_OK) return m->sh[j].rc;
    return EVDB_OK;
}

// ---- lifecycle ---------------------------------------------------------------------------------------
int mstore_create(evdb_store *owner, const evdb_opts *o) {
    MStore *m = new (std::nothrow) MStore();
    if (!m) return EVDB_E_OOM;
    m->owner = owner;
    m->S = o->n_shards;
    for (int i = 0; i < m->S; ++i)
        for (int j = 0; j < i; ++j)
            if (o->devices[i] == o->devices[j]) m->serial = true;
    int rc = EVDB_OK;
    // every pair of distinct devices must reach each other's memory (NVLink / PCIe peer access)
    for (int i = 0; i < m->S && rc == EVDB_OK; ++i)
        for (int j = 0; j < m->S && rc == EVDB_OK; ++j) {
            const int a = o->devices[i], b = o->devices[j];
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { cudaGetLastError(); rc = EVDB_E_UNSUPPORTED; break; }
            if (cudaSetDevice(a) != cudaSuccess) { rc = EVDB_E_CUDA; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_cuda_error(e, __FILE__, __LINE__); rc = EVDB_E_CUDA; }
            cudaGetLastError();
        }
    for (int j = 0; j < m->S && rc == EVDB_OK; ++j) {
        evdb_opts co = *o;
        co.n_shards = 0;
        co.device = o->devices[j];
        co.capacity_hint = (o->capacity_hint + m->S - 1) / m->S;
        rc = evdb_store_create(&co, &m->sh[j].st);
        if (rc != EVDB_OK) break;
        m->sh[j].dev = co.device;
        m->sh[j].st->slot_mul = (uint64_t)m->S;
        if (cudaSetDevice(co.device) != cudaSuccess ||
            cudaMalloc((void **)&m->sh[j].qflags, sizeof(unsigned long long) * (kMS + 2)) != cudaSuccess ||
            cudaMemset(m->sh[j].qflags, 0, sizeof(unsigned long long) * (kMS + 2)) != cudaSuccess) { rc = EVDB_E_CUDA; break; }
        for (int e = 0; e < 4; ++e)
            if (cudaEventCreate(&m->sh[j].ev[e]) != cudaSuccess) { rc = EVDB_E_CUDA; break; }
    }
    owner->multi = m;
    if (rc != EVDB_OK) { mstore_destroy(m); owner->multi = nullptr; return rc; }
    owner->dim = o->dim > 0 ? o->dim : 0;
    if (!m->serial)
        for (int j = 0; j < m->S; ++j) m->sh[j].th = std::thread(worker_main, m, j);
    return EVDB_OK;
}

void mstore_destroy(MStore *m) {
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->quit = true;
    }
    m->cv_go.notify_all();
    for (int j = 0; j < m->S; ++j)
        if (m->sh[j].th.joinable()) m->sh[j].th.join();
    for (int j = 0; j < m->S; ++j)
        if (m->sh[j].st) { cudaSetDevice(m->sh[j].dev); cudaStreamSynchronize(m->sh[j].st->stream); }
    destroy_exchanges(m, m->xb);
    destroy_exchanges(m, m->xw);
    destroy_exchanges(m, m->xe);
    for (int j = 0; j < m->S; ++j) {
        MShard &h = m->sh[j];
        cudaSetDevice(h.dev);
        cudaFree(h.qbuf); cudaFree(h.stage_dev); cudaFree(h.blob); cudaFree(h.merged); cudaFree(h.qflags);
        if (h.h_stage) cudaFreeHost(h.h_stage);
        for (int e = 0; e < 4; ++e) if (h.ev[e]) cudaEventDestroy(h.ev[e]);
        if (h.st) evdb_store_destroy(h.st);
    }
    if (m->h_res) cudaFreeHost(m->h_res);
    free(m->h_esc);
    cudaGetLastError();
    delete m;
}

// ---- ingest -----------------------------------------------------------------------------------------
int m_put(MStore *m, uint64_t slot0, const void *rows, bool is_f64, uint64_t n, int d, bool replace_all) {
    evdb_store *o = m->owner;
    if (replace_all) {   // vector_store:init/1 bulk load: the previous content goes
        for (int j = 0; j < m->S; ++j) { evdb_store *c = m->sh[j].st; c->count = 0; c->shadow_valid = 0; c->l2_valid = 0; }
        o->count = 0;
        slot0 = 0;
    }
    if (n == 0) return EVDB_OK;
    if (slot0 > o->count || slot0 + n > 0xFFFFFFF0ull) return EVDB_E_BAD_ARG;
    MJob jb;
    jb.kind = 2;
    jb.rows = rows; jb.is_f64 = is_f64; jb.first = slot0; jb.n = n; jb.d = d;
    if (n < (uint64_t)(4 * m->S)) {   // a handful of rows: not worth waking the workers
        for (int j = 0; j < m->S; ++j) EVDB_TRY(run_put(m, j, jb));
    } else {
        EVDB_TRY(dispatch(m, jb));
    }
    if (o->dim == 0) o->dim = d;
    if (slot0 + n > o->count) o->count = slot0 + n;
    m->n_upserts += n;
    return EVDB_OK;
}

int m_bulk_codes(MStore *m, const uint8_t *codes, const double *mins, const double *scales, uint64_t n, int d) {
    evdb_store *o = m->owner;
    const size_t src_row = o->dtype == EVDB_U8 ? (size_t)d : (size_t)(d + 1) / 2;
    for (int j = 0; j < m->S; ++j) {
        const uint64_t nj = shard_count(n, m->S, j);
        EVDB_TRY(store_load_codes(m->sh[j].st, nj ? codes + (size_t)j * src_row : nullptr, src_row * m->S,
                                  nj ? mins + j : nullptr, nj ? scales + j : nullptr, (size_t)m->S, nj, d));
    }
    o->count = n;
    if (n > 0 && o->dim == 0) o->dim = d;
    m->n_upserts += n;
    return EVDB_OK;
}

int m_delete(MStore *m, uint32_t slot, int64_t *moved_from) {
    evdb_store *o = m->owner;
    const int S = m->S;
    const uint64_t last = o->count - 1;
    const int js = (int)(slot % S), jl = (int)(last % S);
    evdb_store *cs = m->sh[js].st, *cl = m->sh[jl].st;
    if ((uint64_t)slot != last) {
        if (js == jl) {
            // the global last row is also the last row of this shard: the shard's own swap-with-last
            int64_t mv = -1;
            EVDB_TRY(evdb_store_delete(cs, slot / S, &mv));
        } else {
            // the hole and the last row live on different devices: the raw row (and its codec metadata)
            // crosses over, the cached values are recomputed where it lands
            const uint64_t ls = slot / S, ll = last / S;
            EVDB_TRY(set_dev(m->sh[jl].dev));
            EVDB_CUDA(cudaStreamSynchronize(cl->stream));
            EVDB_TRY(set_dev(m->sh[js].dev));
            cudaStream_t st = cs->stream;
            EVDB_CUDA(cudaMemcpyPeerAsync(cs->rows + ls * cs->row_bytes, m->sh[js].dev, cl->rows + ll * cl->row_bytes,
                                          m->sh[jl].dev, cs->row_bytes, st));
            if (cs->qms64)
                EVDB_CUDA(cudaMemcpyPeerAsync(cs->qms64 + ls, m->sh[js].dev, cl->qms64 + ll, m->sh[jl].dev, sizeof(double2), st));
            EVDB_TRY(store_refinalize(cs, ls, 1, st));
            EVDB_CUDA(cudaStreamSynchronize(st));
            store_drop_last(cl);
        }
        if (moved_from) *moved_from = (int64_t)last;
    } else {
        store_drop_last(cl);
    }
    o->count = last;
    m->n_deletes++;
    return EVDB_OK;
}

int m_get_f64(MStore *m, uint32_t slot, double *out, int d) {
    return evdb_store_get_f64(m->sh[slot % m->S].st, slot / m->S, out, d);
}
int m_get_codes(MStore *m, uint32_t slot, uint8_t *codes, double *mn, double *scale) {
    return evdb_store_get_codes(m->sh[slot % m->S].st, slot / m->S, codes, mn, scale);
}

int m_fill_synthetic(MStore *m, uint64_t seed, uint64_t row0, uint64_t n, int d) {
    for (int j = 0; j < m->S; ++j)
        EVDB_TRY(store_fill_synthetic(m->sh[j].st, seed, row0 + j, (uint64_t)m->S, shard_count(n, m->S, j), d));
    m->owner->count = n;
    if (m->owner->dim == 0) m->owner->dim = d;
    return EVDB_OK;
}

int m_stats(MStore *m, evdb_stats *out) {
    evdb_store *o = m->owner;
    out->count = o->count;
    out->dimension = o->dim;
    out->dtype = o->dtype;
    out->device = m->sh[0].dev;
    out->last_plan = m->last_plan;
    out->n_shards = m->S;
    for (int j = 0; j < m->S; ++j) {
        evdb_stats cs;
        EVDB_TRY(evdb_store_stats(m->sh[j].st, &cs));
        out->capacity += cs.capacity;
        out->device_bytes += cs.device_bytes + m->sh[j].qbuf_cap + m->sh[j].blob_cap + m->sh[j].merged_cap + m->sh[j].stage_dev_cap;
        out->shadow_bytes += cs.shadow_bytes;
        out->rows_scanned += cs.rows_scanned;
        out->kernel_launches += cs.kernel_launches;
        out->gemm_disabled |= cs.gemm_disabled;
    }
    out->searches = m->n_searches;
    out->escalations = m->n_escalations;
    out->upserts = m->n_upserts;
    out->deletes = m->n_deletes;
    out->last_search_ms = m->last_ms;
    out->last_h2d_ms = m->last_h2d_ms;
    out->last_device_ms = m->last_dev_ms;
    out->last_d2h_ms = m->last_d2h_ms;
    return EVDB_OK;
}

int m_set_plan(MStore *m, int plan) {
    for (int j = 0; j < m->S; ++j) EVDB_TRY(evdb_store_set_plan(m->sh[j].st, plan));
    m->owner->plan = plan;
    return EVDB_OK;
}
int m_flush(MStore *m) {
    for (int j = 0; j < m->S; ++j) EVDB_TRY(evdb_store_flush(m->sh[j].st));
    return EVDB_OK;
}
int m_profile(MStore *m, int enable) { return evdb_store_profile(m->sh[0].st, enable); }
int m_profile_read(MStore *m, int32_t *n, double *ms) { return evdb_store_profile_read(m->sh[0].st, n, ms); }

// ---- search ----------------------------------------------------------------------------------------
static int gemm_window_of(int k, uint64_t n_total) {
    const int kk = (uint64_t)k < n_total ? k : (int)n_total;
    return gemm_kp(choose_kp(kk, 0));
}

// one pass over all shards for B queries (host pointers); the merged packed result lands in m->h_res
static int search_pass(MStore *m, const void *queries, bool is_f64, int B, int d, int k, int metric, int plan, int kp_min) {
    evdb_store *o = m->owner;
    const int S = m->S;
    MJob jb;
    jb.kind = 1;
    jb.queries = queries; jb.is_f64 = is_f64; jb.B = B; jb.d = d; jb.k = k; jb.metric = metric;
    jb.plan = plan; jb.kp_min = kp_min; jb.n_total = o->count;
    // The tcgen05 two-phase search must be chosen from GLOBAL facts so that every shard takes the same
    // path: an F32 store with the operand column, cosine / euclidean, a real batch, a window of at
    // most 128 keys, every shard at least one corpus tile.
    const int KP = gemm_window_of(k, o->count);
    const uint64_t smallest = shard_count(o->count, S, S - 1);
    // (one exchange of finished per-shard results is the default since the group-of-warps fold: re-ranking a
    //  whole local window costs a shard little more than its share of the global one, and it saves a cross-GPU
    //  wait; EVDB_SHARD_TWO_PHASE=1 selects the window / owner scheme, which stays tested)
    const char *tp_env = getenv("EVDB_SHARD_TWO_PHASE");   // read per call: tests flip it
    const bool two_phase_on = tp_env && atoi(tp_env) == 1;
    jb.two_phase = two_phase_on && S > 1 && S <= 32 && plan == EVDB_PLAN_AUTO && kp_min == 0 && o->plan == EVDB_PLAN_AUTO &&
                   o->dtype == EVDB_F32 && o->gemm_shadow && (metric == EVDB_COSINE || metric == EVDB_EUCLIDEAN) &&
                   B >= 16 && B <= gemm_max_batch() && KP <= 128 && (size_t)S * KP <= 2048 && smallest >= 256;
    if (jb.two_phase)
        for (int j = 0; j < S; ++j)
            if (m->sh[j].st->gemm_oom || !gemm_plan_supported(m->sh[j].st, metric, B, KP)) jb.two_phase = false;
    if (jb.two_phase) {
        EVDB_TRY(ensure_exchanges(m, m->xw, &m->xw_words, (uint64_t)B * KP + B));
        EVDB_TRY(ensure_exchanges(m, m->xe, &m->xe_words, (uint64_t)B * KP));
    } else if (S > 1) {
        EVDB_TRY(ensure_exchanges(m, m->xb, &m->xb_words, blob_words(B, k)));
    }
    const size_t qbytes = (size_t)B * d * sizeof(double), bbytes = blob_words(B, k) * 8;
    for (int j = 0; j < S; ++j) {
        MShard &h = m->sh[j];
        if (h.qbuf_cap < qbytes || h.blob_cap < bbytes || (j == 0 && h.merged_cap < bbytes)) {
            EVDB_TRY(set_dev(h.dev));
            EVDB_CUDA(cudaStreamSynchronize(shard_stream(m, j)));
            EVDB_TRY(ensure_bytes((void **)&h.qbuf, &h.qbuf_cap, qbytes));
            EVDB_TRY(ensure_bytes((void **)&h.blob, &h.blob_cap, bbytes));
            if (j == 0) EVDB_TRY(ensure_bytes((void **)&h.merged, &h.merged_cap, bbytes));
        }
    }
    EVDB_TRY(ensure_bytes(&m->h_res, &m->h_res_cap, bbytes, true));
    m->qepoch++;
    EVDB_TRY(dispatch(m, jb));
    for (int j = 0; j < S; ++j)
        if (m->sh[j].bad_query) return EVDB_E_BAD_VECTOR;
    m->last_plan = jb.two_phase ? EVDB_PLAN_GEMM : m->sh[0].st->last_plan;
    return EVDB_OK;
}

int m_search_host(MStore *m, const void *queries, bool is_f64, int B, int d, int k, int metric, uint32_t *out_slots,
                  double *out_dists, int32_t *out_counts) {
    evdb_store *o = m->owner;
    const int kcaller = k;
    if ((uint64_t)k > o->count) k = (int)o->count;    // lists:sublist(Sorted, K) with K > N: all N rows
    EVDB_TRY(search_pass(m, queries, is_f64, B, d, k, metric, EVDB_PLAN_AUTO, 0));
    {   // per-phase device times of this call, from shard 0's events
        float a = 0.f, b = 0.f, c = 0.f;
        MShard &h = m->sh[0];
        cudaSetDevice(h.dev);
        cudaEventElapsedTime(&a, h.ev[0], h.ev[1]);
        cudaEventElapsedTime(&b, h.ev[1], h.ev[2]);
        cudaEventElapsedTime(&c, h.ev[2], h.ev[3]);
        m->last_h2d_ms = a; m->last_dev_ms = b; m->last_d2h_ms = c; m->last_ms = a + b + c;
    }
    m->n_searches += (uint64_t)B;
    const size_t nk = (size_t)B * k;
    const uint64_t *h_ids = (const uint64_t *)m->h_res;
    const double *h_d = (const double *)(h_ids + nk);
    const int32_t *h_c = (const int32_t *)(h_d + nk);
    auto emit = [&](int b, const uint64_t *ids, const double *dd, int cnt) {
        out_counts[b] = cnt;
        for (int j = 0; j < kcaller; ++j) {
            const size_t oo = (size_t)b * kcaller + j;
            if (j < cnt) { out_slots[oo] = (uint32_t)ids[j]; out_dists[oo] = dd[j]; }
            else { out_slots[oo] = 0xFFFFFFFFu; out_dists[oo] = 0.0; }
        }
    };
    int nflag = 0;
    for (int b = 0; b < B; ++b) {
        emit(b, h_ids + (size_t)b * k, h_d + (size_t)b * k, h_c[b]);
        nflag += h_c[B + b] ? 1 : 0;
    }
    if (nflag == 0) return EVDB_OK;
    // ---- the escalation ladder of store.cu search_host, on every shard at once ----
    m->n_escalations += (uint64_t)nflag;
    const int keep_plan = m->last_plan;
    int *idx = (int *)malloc(sizeof(int) * (size_t)nflag);
    if (!idx) return EVDB_E_OOM;
    int nf = 0;
    for (int b = 0; b < B; ++b) if (h_c[B + b]) idx[nf++] = b;
    const size_t esz = is_f64 ? 8 : 4;
    int rc = EVDB_OK;
    for (int attempt = 0; attempt < 2 && nf > 0 && rc == EVDB_OK; ++attempt) {
        const int plan = attempt == 0 ? EVDB_PLAN_SCAN : EVDB_PLAN_EXACT;
        if (attempt == 0 && choose_kp(k, 256) > kMaxKP) continue;
        if (m->h_esc_cap < (size_t)nf * d * esz) {
            free(m->h_esc);
            m->h_esc = malloc((size_t)nf * d * esz);
            m->h_esc_cap = m->h_esc ? (size_t)nf * d * esz : 0;
            if (!m->h_esc) { rc = EVDB_E_OOM; break; }
        }
        for (int i = 0; i < nf; ++i)
            memcpy((uint8_t *)m->h_esc + (size_t)i * d * esz, (const uint8_t *)queries + (size_t)idx[i] * d * esz, (size_t)d * esz);
        rc = search_pass(m, m->h_esc, is_f64, nf, d, k, metric, plan, attempt == 0 ? 256 : 0);
        if (rc != EVDB_OK) break;
        const size_t nk2 = (size_t)nf * k;
        const uint64_t *e_ids = (const uint64_t *)m->h_res;
        const double *e_d = (const double *)(e_ids + nk2);
        const int32_t *e_c = (const int32_t *)(e_d + nk2);
        int keep = 0;
        for (int i = 0; i < nf; ++i) {
            emit(idx[i], e_ids + (size_t)i * k, e_d + (size_t)i * k, e_c[i]);
            if (e_c[nf + i]) idx[keep++] = idx[i];
        }
        nf = keep;
    }
    free(idx);
    m->last_plan = keep_plan;
    return rc;
}

}  // namespace evdb
