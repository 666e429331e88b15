// select.cu -- per-query candidate merge, exact fp64 re-rank and final top-k.
//
// Replaces lists:sort/1 + lists:sublist/2 of perform_search/3 (reference
// src/vector_store.erl:233-236).  One CTA per query:
//   1. bitonic-merge the per-CTA candidate lists of the scan / GEMM stage into
//      the KP best approximate (score, slot) keys;
//   2. one warp per candidate recomputes the distance in fp64 in the
//      reference's exact operation order (exact.cuh);
//   3. sort by (exact distance, slot), emit the first k, and PROVE the window
//      complete: every row outside it has approximate score >= the window's
//      last key, hence exact distance >= bound - eps; if the k-th exact
//      distance is not strictly below that, flag the query for escalation.
// Also: the exhaustive fp64 plan (every row, then a stable radix sort) that
// backs k > kMaxKP, metrics without a fast scan, and failed escalations; and
// the G-way merge that follows the cross-GPU allgather.
#include <string.h>

#include <cub/device/device_radix_sort.cuh>

#include "exact.cuh"
#include "internal.h"
#include "topk.cuh"
#include "scan_common.cuh"

namespace evdb {

constexpr int kSelSort = 4096;  // keys sorted / selected per merge round

// ---- fused push (exchange.cu): store one word into my slot of EVERY rank's mailbox ----
__device__ __forceinline__ void push_store(const PushTarget &t, size_t word, uint64_t v) {
    for (int p = 0; p < t.world; ++p) t.peer_box[p][t.slot_off + word] = v;
}
// Every warp of the kernel calls this once after its last push_store (all 32 lanes): the warp that
// arrives last publishes the epoch in every rank's mailbox.
__device__ __forceinline__ void push_arrive(const PushTarget &t, unsigned int total_warps, int lane) {
    __threadfence_system();
    __syncwarp();
    unsigned int prev = 0;
    if (lane == 0) prev = atomicAdd(t.counter, 1u);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == total_warps - 1) {
        __threadfence_system();
        if (lane < t.world)
            *reinterpret_cast<volatile unsigned long long *>(t.peer_box[lane] + t.flag_off) = t.epoch;
        if (lane == 0) *t.counter = 0;
    }
}

struct SelectArgs {
    const uint8_t *rows;
    size_t row_bytes;
    const double *norm64;
    const double2 *qms64;
    uint64_t n;
    int d;
    const double *q64;        // [B][d]
    const uint64_t *partial;  // [B][L][KP] ascending lists (scan plans), or NULL with `raw` set
    RawCands raw;             // unsorted per-(CTA, part) candidate buffers of the GEMM plan
    int L, KP, kk, kstride, metric;
    float eps_abs, eps_rel;
    const float *eps_q;  // optional [B]: per-query absolute bound (GEMM plans)
    int squared;         // key scores are squared distances (euclidean GEMM plan)
    int variant;         // EVDB_SEL_VARIANT (measurement only): 16 = per-phase cycle counters, 32 = CTA-per-query kernel for GEMM batches too
    int win_mode;        // sharded search, phase 1: push the ascending window (keys with GLOBAL rows) + meta to every rank and stop
    PushTarget push;     //   blob words: [B*KP keys][B meta = (eps bits << 32) | ncand]
    uint64_t slot_base;  // returned id = slot_base + slot * slot_mul (slot_mul > 1: round-robin shard of a multi-device store)
    uint64_t slot_mul;
    uint64_t *out_ids;
    double *out_dists;
    int32_t *out_counts;
    int32_t *out_flags;
};

// Block-wide selection: keep the `need` smallest of buf[0, n) (need < n <= kSelSort), compacted
// to buf[0, need) in arbitrary order.  Bisection on the 32-bit score with block-wide counts
// (4-way, one barrier pair per round), then one compaction pass; ties at the cut are taken
// first come first served (any `need` keys not above the cut are a valid window).
// tmp: >= need u64 of scratch.  Every thread of the block must call.
__device__ void block_select_smallest(uint64_t *buf, int n, int need, uint64_t *tmp) {
    __shared__ uint32_t s_lo, s_hi;
    __shared__ int s_c[8];
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t sc = (uint32_t)(buf[i] >> 32);
        mn = min(mn, sc);
        mx = max(mx, sc);
    }
    if (threadIdx.x == 0) { s_lo = 0xFFFFFFFFu; s_hi = 0u; }
    if (threadIdx.x < 8) s_c[threadIdx.x] = 0;
    __syncthreads();
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_lo, mn); atomicMax(&s_hi, mx); }
    __syncthreads();
    uint32_t lo = s_lo, hi = s_hi;  // invariant: count(score <= hi) >= need, count(score < lo) < need
    int it = 0;
    while (lo < hi) {
        const uint32_t span = hi - lo;
        const uint32_t qd = span >> 2;
        const uint32_t p2 = lo + (span >> 1);
        const uint32_t p1 = qd ? lo + qd : p2;
        const uint32_t p3 = qd ? p2 + qd : p2;
        int c1 = 0, c2 = 0, c3 = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t sc = (uint32_t)(buf[i] >> 32);
            c1 += sc <= p1 ? 1 : 0;
            c2 += sc <= p2 ? 1 : 0;
            c3 += sc <= p3 ? 1 : 0;
        }
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        int *cc = s_c + (it & 1) * 3;
        if ((threadIdx.x & 31) == 0) { atomicAdd(cc, c1); atomicAdd(cc + 1, c2); atomicAdd(cc + 2, c3); }
        __syncthreads();
        const int t1 = cc[0], t2 = cc[1], t3 = cc[2];
        if (threadIdx.x < 3) s_c[((it + 1) & 1) * 3 + threadIdx.x] = 0;
        if (t1 >= need) hi = p1;
        else if (t2 >= need) { lo = p1 + 1; hi = p2; }
        else if (t3 >= need) { lo = p2 + 1; hi = p3; }
        else lo = p3 + 1;
        ++it;
        __syncthreads();
    }
    // lo == the need-th smallest score.  Compact: everything below it, then ties up to `need`.
    if (threadIdx.x < 2) s_c[6 + threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = buf[i];
        if ((uint32_t)(key >> 32) < lo) tmp[atomicAdd(&s_c[6], 1)] = key;
    }
    __syncthreads();
    const int n_less = s_c[6];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = buf[i];
        if ((uint32_t)(key >> 32) == lo) {
            const int t = atomicAdd(&s_c[7], 1);
            if (n_less + t < need) tmp[n_less + t] = key;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < need; i += blockDim.x) buf[i] = tmp[i];
    __syncthreads();
}

// Lone-query re-rank (1024 threads, <= 30 candidates to re-rank): warp r forms the products of
// candidate r chunk by chunk (one more warp the query's squares for cosine) into a double-buffered
// stage, and ONE warp folds all rows in lock-step -- lane r adds row r's terms left to right, in
// the reference order.  One fp64 chain instruction stream for all candidates: the chains no longer
// contend for the SM's fp64 pipe, and staging chunk c+1 overlaps folding chunk c.
constexpr int kFoldStride = kExactChunk + 1;   // doubles per staged row: lanes r, r+16 share a bank, no others
constexpr int kFoldRows = 31;                  // producer warps 0..30, folder = warp 31
template <int DTYPE>
__device__ __forceinline__ void lone_rerank(const SelectArgs &a, const uint64_t *buf, int nrer, int nsort,
                                            const double *__restrict__ q, double *stage, uint64_t *dkey,
                                            uint64_t *dslot, int warp, int lane) {
    const int d = a.d, metric = a.metric;
    const bool cosine = metric == EVDB_COSINE;
    const int nrows = nrer + (cosine ? 1 : 0);
    const int nchunks = (d + kExactChunk - 1) / kExactChunk;
    constexpr int kPer = kExactChunk / kWarp;
    const bool producer = warp < nrows;
    const bool qrow = cosine && warp == nrer;
    const bool folder = warp == kFoldRows;
    const uint8_t *row = a.rows;
    double mn = 0.0, sc = 0.0;
    if (producer && !qrow) {
        const uint32_t slot = key_slot(buf[warp]);
        row = a.rows + (size_t)slot * a.row_bytes;
        if (DTYPE == EVDB_U8 || DTYPE == EVDB_U4) {
            const double2 ms = a.qms64[slot];
            mn = ms.x;
            sc = ms.y;
        }
        // the whole row into L2 now: a DRAM round trip (~2 k cycles) is longer than folding one
        // chunk, so loading chunk c+1 from DRAM while chunk c folds would stall every chunk
        const int lines = (int)((a.row_bytes + 254) / 128);
        for (int l = lane; l < lines; l += 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + (size_t)l * 128));
    }
    uint32_t fslot = 0;
    double vnorm = 0.0;
    if (folder && lane < nrer) {
        fslot = key_slot(buf[lane]);
        if (cosine) vnorm = a.norm64[fslot];
    }
    double nv[kPer], nq[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int t = i * kWarp + lane;
        nq[i] = (producer && t < d) ? q[t] : 0.0;
        nv[i] = (producer && !qrow && t < d) ? row_elem<DTYPE>(row, t, mn, sc) : nq[i];
    }
    double acc = 0.0;
    for (int c = 0; c <= nchunks; ++c) {
        if (producer && c < nchunks) {
            double *dst = stage + (size_t)((c & 1) * kFoldRows + warp) * kFoldStride;
            const int base = c * kExactChunk;
            const int cnt = min(kExactChunk, d - base);
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int t = i * kWarp + lane;
                const double v = nv[i], qq = nq[i];
                if (t < cnt) {
                    if (cosine) {
                        dst[t] = __dmul_rn(qq, v);           // dot_product X*Y; the q row: v == qq, X*X
                    } else if (metric == EVDB_EUCLIDEAN) {
                        const double t0 = __dsub_rn(qq, v);  // vector_subtract
                        dst[t] = __dmul_rn(t0, t0);
                    } else {
                        dst[t] = fabs(__dsub_rn(qq, v));     // abs(X - Y)
                    }
                }
            }
            const int nb = base + kExactChunk;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int t = nb + i * kWarp + lane;
                nq[i] = t < d ? q[t] : 0.0;
                nv[i] = (!qrow && t < d) ? row_elem<DTYPE>(row, t, mn, sc) : nq[i];
            }
        }
        if (folder && c > 0 && lane < nrows) {
            const int base = (c - 1) * kExactChunk;
            acc = fold_staged(acc, stage + (size_t)(((c - 1) & 1) * kFoldRows + lane) * kFoldStride,
                              min(kExactChunk, d - base));
        }
        __syncthreads();
    }
    if (folder) {
        double dist;
        if (cosine) {
            const double sq = __shfl_sync(0xffffffffu, acc, nrer);
            const double n1 = __dsqrt_rn(sq), n2 = vnorm;
            dist = (n1 == 0.0 || n2 == 0.0) ? 1.0 : __dsub_rn(1.0, __ddiv_rn(acc, __dmul_rn(n1, n2)));
        } else if (metric == EVDB_EUCLIDEAN) {
            dist = __dsqrt_rn(acc);
        } else {
            dist = acc;
        }
        if (lane < nrer) {
            dkey[lane] = f64_orderable(dist);
            dslot[lane] = fslot;
        } else if (lane < nsort) {
            dkey[lane] = kKeyMax;
            dslot[lane] = kKeyMax;
        }
    }
}

// EVDB_SEL_VARIANT bit 4 (tuning aid): cycles per phase, summed over CTAs
__device__ unsigned long long g_sel_dbg[10];
#define SEL_MARK(i) do { if ((a.variant & 16) && threadIdx.x == 0) { long long _t = clock64(); atomicAdd(&g_sel_dbg[i], (unsigned long long)(_t - t_mark)); t_mark = _t; } } while (0)

// THREADS = 1024 for a lone query (latency), 256 for batches (more CTAs per SM, cheaper barriers).
template <int DTYPE, int THREADS>
__device__ __forceinline__ void select_body(const SelectArgs &a, const int b, uint8_t *smem) {
    constexpr int kSelWarps = THREADS / 32;
    uint64_t *buf = reinterpret_cast<uint64_t *>(smem);                       // [kSelSort]
    uint64_t *dkey = buf + kSelSort;                                          // [kMaxKP]
    uint64_t *dslot = dkey + kMaxKP;                                          // [kMaxKP]
    double *sp_all = reinterpret_cast<double *>(dslot + kMaxKP);              // lone-query variant: [warps][2*chunk]
    __shared__ int s_ncand;
    __shared__ float s_bound;

    const int KP = a.KP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float eps_abs = a.eps_abs + (a.eps_q ? a.eps_q[b] : 0.f);
    long long t_mark = clock64();

    if (a.raw.cand) {
        // ---- 1a. GEMM plan: gather this query's unsorted candidate buffers, keep the KP best ----
        __shared__ int s_off[kRawMaxLists + 1];  // exclusive prefix of the buffer fill counts
        __shared__ int s_cut;
        const RawCands &rw = a.raw;
        const int L = a.L;                       // NG * parts buffers hold candidates of this query
        const int blk = b / rw.gm, et = b % rw.gm;
        const int c = blk / rw.MB, mb_local = blk % rw.MB;
        auto list_index = [&](int l) -> size_t {  // buffer set of (row group ng, column part)
            const int ng = l / rw.parts, part = l % rw.parts;
            const int cta = ng * rw.MB + mb_local;
            return ((size_t)c * rw.nCTA + cta) * rw.parts + part;
        };
        for (int l = threadIdx.x; l < L; l += blockDim.x) s_off[l + 1] = rw.cnt[list_index(l) * rw.gm + et];
        if (threadIdx.x == 0) s_off[0] = 0;
        __syncthreads();
        if (warp == 0) {  // inclusive scan of s_off[1..L], 32 at a time
            int carry = 0;
            for (int base = 1; base <= L; base += 32) {
                const int i = base + lane;
                int v = i <= L ? s_off[i] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += u;
                }
                if (i <= L) s_off[i] = v + carry;
                carry += __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncthreads();
        SEL_MARK(0);
        int carried = 0, l0 = 0;
        bool sorted = false;
        while (l0 < L) {
            // as many whole buffers as fit next to the carried keys (a buffer holds <= 256 keys)
            int lo = l0 + 1, hi = L;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (carried + s_off[mid] - s_off[l0] <= kSelSort) lo = mid; else hi = mid - 1;
            }
            const int l1 = lo;
            // one candidate per thread: the buffers hold a key or two each, walking them one after
            // the other would put a memory latency on every buffer
            const int e0 = s_off[l0], e1 = s_off[l1];
            for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
                int a0 = l0, a1 = l1 - 1;           // largest l with s_off[l] <= e
                while (a0 < a1) {
                    const int mid = (a0 + a1 + 1) >> 1;
                    if (s_off[mid] <= e) a0 = mid; else a1 = mid - 1;
                }
                const uint64_t *src = rw.cand + list_index(a0) * rw.cap * rw.gm + et;  // entry i at [i*gm]
                buf[carried + (e - e0)] = __ldcg(src + (size_t)(e - s_off[a0]) * rw.gm);
            }
            const int filled = carried + e1 - e0;
            __syncthreads();
            SEL_MARK(1);
            if ((a.variant & 16) && threadIdx.x == 0) atomicAdd(&g_sel_dbg[8], (unsigned long long)(e1 - e0));
            if (filled > KP) {
                if (KP <= 128 && filled >= 2 * (KP <= 64 ? 256 : 512)) {
                    // Hundreds to thousands of keys of which KP are wanted (a full 1024-key sort costs
                    // ~17 k cycles): the KP-th smallest of a strided subset bounds the KP-th smallest
                    // overall, so only keys <= it can matter -- typically ~ filled * KP / subset of
                    // them; one small sort finishes the round.
                    uint64_t *sub = dslot;                 // 1024 keys of scratch (results come later)
                    const int nsub = KP <= 64 ? 256 : 512;
                    const int stride = filled / nsub;
                    for (int i = threadIdx.x; i < nsub; i += blockDim.x) sub[i] = buf[i * stride];
                    if (threadIdx.x == 0) s_cut = 0;
                    __syncthreads();
                    block_bitonic_sort_fast(sub, nsub, dkey);
                    const uint64_t T = sub[KP - 1];
                    __syncthreads();
                    for (int i = threadIdx.x; i < filled; i += blockDim.x) {
                        const uint64_t key = buf[i];
                        if (key <= T) {
                            const int p = atomicAdd(&s_cut, 1);
                            if (p < 1024) sub[p] = key;
                        }
                    }
                    __syncthreads();
                    const int cnt = s_cut;
                    if (cnt <= 1024) {
                        int np = KP;
                        while (np < cnt) np <<= 1;
                        for (int i = cnt + threadIdx.x; i < np; i += blockDim.x) sub[i] = kKeyMax;
                        __syncthreads();
                        block_bitonic_sort_fast(sub, np, dkey);
                        for (int i = threadIdx.x; i < KP; i += blockDim.x) buf[i] = sub[i];
                        __syncthreads();
                        sorted = l1 == L;
                    } else {
                        block_select_smallest(buf, filled, KP, dkey);
                    }
                } else if (l1 == L && filled <= (int)blockDim.x) {
                    // last round, one key per thread: a single sort selects and orders
                    int np = KP;
                    while (np < filled) np <<= 1;
                    for (int i = filled + threadIdx.x; i < np; i += blockDim.x) buf[i] = kKeyMax;
                    __syncthreads();
                    block_bitonic_sort_fast(buf, np, dkey);
                    sorted = true;
                } else {
                    block_select_smallest(buf, filled, KP, dkey);
                }
                carried = KP;
            } else {
                carried = filled;
            }
            l0 = l1;
            SEL_MARK(2);
        }
        if (!sorted) {
            for (int i = carried + threadIdx.x; i < KP; i += blockDim.x) buf[i] = kKeyMax;
            __syncthreads();
            block_bitonic_sort_fast(buf, KP, dkey);
        }
    } else {
    // ---- 1b. scan plans: merge L ascending lists of KP keys into the KP smallest ----
    // Fast path: pool the first P = ceil(KP/L) keys of every list; the KP-th smallest of the pool
    // is an upper bound T on the global KP-th smallest key (KP distinct keys are <= it), so only
    // keys <= T can matter -- typically not many more than KP.
    const uint64_t *src = a.partial + (size_t)b * a.L * KP;
    const size_t total = (size_t)a.L * KP;
    __shared__ int s_cnt;
    bool merged = false;
    // heads per list so that the pool holds at least KP keys (lists are ascending)
    const int P = (KP + a.L - 1) / a.L;
    const bool pool_ok = (size_t)a.L * P <= (size_t)kSelSort && P <= KP;
    if (pool_ok) {
        const int pool = a.L * P;
        int Lp = 2;
        while (Lp < pool) Lp <<= 1;
        // any KP keys bound the KP-th smallest: with many lists the heads of the first 256 do, and a
        // 256-key sort is several times cheaper than a 1024-key one
        const int pool_used = (Lp > 256 && P == 1 && KP <= 128) ? 256 : pool;
        if (pool_used < pool) Lp = 256;
        for (int i = threadIdx.x; i < Lp; i += blockDim.x)
            buf[i] = i < pool_used ? src[(size_t)(i / P) * KP + (i % P)] : kKeyMax;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        block_bitonic_sort_fast(buf, Lp, dkey);
        const uint64_t T = buf[KP - 1];  // >= KP keys are <= T (kKeyMax if the lists hold fewer)
        __syncthreads();
        for (size_t i0 = threadIdx.x; i0 < total; i0 += (size_t)blockDim.x * 8) {
            uint64_t kk[8];   // eight loads in flight per thread, then the filter
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const size_t i = i0 + (size_t)u * blockDim.x;
                kk[u] = i < total ? __ldcg(src + i) : kKeyMax;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (kk[u] <= T && kk[u] != kKeyMax) {
                    int p = atomicAdd(&s_cnt, 1);
                    if (p < kSelSort) buf[p] = kk[u];
                }
        }
        __syncthreads();
        const int cnt = s_cnt;
        if (cnt <= kSelSort) {
            int nsort = KP;
            while (nsort < cnt) nsort <<= 1;
            for (int i = cnt + threadIdx.x; i < nsort; i += blockDim.x) buf[i] = kKeyMax;
            __syncthreads();
            block_bitonic_sort_fast(buf, nsort, dkey);
            merged = true;
        }
        __syncthreads();
    }
    if (!merged) {  // general path: chunked sort, carrying the KP best
        size_t pos = 0;
        int carried = 0;
        while (true) {
            size_t room = (size_t)kSelSort - carried;
            size_t take = total - pos < room ? total - pos : room;
            int filled = carried + (int)take;
            int nsort = KP;
            while (nsort < filled) nsort <<= 1;
            for (int i = threadIdx.x; i < nsort - carried; i += blockDim.x)
                buf[carried + i] = (size_t)i < take ? src[pos + i] : kKeyMax;
            __syncthreads();
            block_bitonic_sort_fast(buf, nsort, dkey);
            pos += take;
            carried = KP;
            if (pos >= total) break;
        }
    }
    }
    // ---- count valid candidates (keys are ascending; kKeyMax pads) ----
    SEL_MARK(3);
    if (threadIdx.x == 0) s_ncand = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < KP; i += blockDim.x)
        if (buf[i] != kKeyMax && (i + 1 == KP || buf[i + 1] == kKeyMax)) {
            s_ncand = i + 1;
            s_bound = key_score(buf[i]);
        }
    __syncthreads();
    const int ncand = s_ncand;
    const float bound = s_bound;

    // ---- 2. exact fp64 distance per candidate, one LANE each ----
    // Only candidates that can still reach the top k are re-ranked: a true top-k row r has
    // exact D_r <= D_(k) <= s_k + eps (the k best approximate scores bound the k-th exact one),
    // hence approximate score <= s_k + 2*eps.  Scores are ascending, so that set is a prefix.
    const int kout = a.kk < ncand ? a.kk : ncand;
    __shared__ int s_nrer;
    if (threadIdx.x == 0) s_nrer = kout;
    __syncthreads();
    if (kout > 0) {
        const float sk = key_score(buf[kout - 1]);
        const float lim = sk + 2.0f * (eps_abs + a.eps_rel * fabsf(sk)) * 1.0001f;
        for (int i = kout + threadIdx.x; i < ncand; i += blockDim.x)
            if (key_score(buf[i]) <= lim && (i + 1 == ncand || !(key_score(buf[i + 1]) <= lim))) s_nrer = i + 1;
    }
    __syncthreads();
    const int nrer = s_nrer;
    SEL_MARK(4);
    if ((a.variant & 16) && threadIdx.x == 0) atomicAdd(&g_sel_dbg[7], (unsigned long long)nrer);
    int nsort = 2;
    while (nsort < nrer) nsort <<= 1;
    const double *q = a.q64 + (size_t)b * a.d;
    if (THREADS == 1024 && nrer + (a.metric == EVDB_COSINE ? 1 : 0) <= kFoldRows && nrer > 0 && !(a.variant & 64)) {
        lone_rerank<DTYPE>(a, buf, nrer, nsort, q, sp_all, dkey, dslot, warp, lane);
    } else if (THREADS == 1024) {
        // Many candidates to re-rank (or EVDB_SEL_VARIANT bit 6): one warp per candidate, 32 lanes
        // form the independent products, one lane folds them in the reference order.
        double *sp = sp_all + warp * 2 * kExactChunk;
        for (int j = warp; j < nsort; j += kSelWarps) {
            if (j < nrer) {
                const uint32_t slot = key_slot(buf[j]);
                const uint8_t *row = a.rows + (size_t)slot * a.row_bytes;
                double mn = 0.0, sc = 0.0;
                if (DTYPE == EVDB_U8 || DTYPE == EVDB_U4) {
                    const double2 ms = a.qms64[slot];
                    mn = ms.x;
                    sc = ms.y;
                }
                const double dist = exact_distance_warp<DTYPE>(row, mn, sc, q, a.d, a.metric, a.norm64[slot], sp, lane);
                if (lane == 0) {
                    dkey[j] = f64_orderable(dist);
                    dslot[j] = slot;
                }
            } else if (lane == 0) {
                dkey[j] = kKeyMax;
                dslot[j] = kKeyMax;
            }
        }
    } else {
        // Batches: the fold of one row is a strictly sequential fp64 chain (reference order), so candidates
        // are spread one per lane: cpw per warp, plus -- for cosine -- lane 31 of every working
        // warp folding the query's own squares.
        const bool cosine = a.metric == EVDB_COSINE;
        const int max_cpw = cosine ? 31 : 32;
        // as few warps as possible: the fp64 unit issues per warp instruction whatever the number of
        // active lanes, and the chain is no shorter with fewer rows per warp
        const int cpw = max_cpw;
        for (int base = warp * cpw; base < nrer; base += kSelWarps * cpw) {
            const int j = base + lane;
            const bool mine = lane < cpw && j < nrer;
            const bool qlane = cosine && lane == 31;
            const uint32_t slot = mine ? key_slot(buf[j]) : key_slot(buf[base]);
            const uint8_t *row = a.rows + (size_t)slot * a.row_bytes;
            double mn = 0.0, sc = 0.0;
            if (DTYPE == EVDB_U8 || DTYPE == EVDB_U4) {
                const double2 ms = a.qms64[slot];
                mn = ms.x;
                sc = ms.y;
            }
            const double s = exact_fold_lane<DTYPE>(row, mn, sc, q, a.d, a.metric, qlane);
            double dist;
            if (cosine) {
                const double sq = __shfl_sync(0xffffffffu, s, 31);
                const double n1 = __dsqrt_rn(sq), n2 = a.norm64[slot];
                dist = (n1 == 0.0 || n2 == 0.0) ? 1.0 : __dsub_rn(1.0, __ddiv_rn(s, __dmul_rn(n1, n2)));
            } else if (a.metric == EVDB_EUCLIDEAN) {
                dist = __dsqrt_rn(s);
            } else {
                dist = s;
            }
            if (mine) {
                dkey[j] = f64_orderable(dist);
                dslot[j] = slot;
            }
        }
        for (int j = nrer + threadIdx.x; j < nsort; j += blockDim.x) {
            dkey[j] = kKeyMax;
            dslot[j] = kKeyMax;
        }
    }
    __syncthreads();

    // ---- 3. final order by (exact distance, slot); emit k; completeness proof ----
    SEL_MARK(5);
    block_bitonic_sort_pairs_fast(dkey, dslot, nsort, buf);
    for (int i = threadIdx.x; i < a.kstride; i += blockDim.x) {
        size_t o = (size_t)b * a.kstride + i;
        if (i < kout) {
            uint64_t ob = dkey[i];
            uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            a.out_dists[o] = __longlong_as_double((long long)bits);
            a.out_ids[o] = a.slot_base + dslot[i] * a.slot_mul;
        } else {
            a.out_dists[o] = 0.0;
            a.out_ids[o] = kKeyMax;
        }
    }
    if (threadIdx.x == 0) {
        a.out_counts[b] = kout;
        int flag = 0;
        if (kout > 0 && (uint64_t)ncand < a.n) {  // rows exist outside the window
            uint64_t ob = dkey[kout - 1];
            uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            double dk = __longlong_as_double((long long)bits);
            double lim = (double)bound - (double)eps_abs - (double)a.eps_rel * fabs((double)bound);
            // squared keys: compare the exact distance in the same domain (rounded up a hair)
            if (a.squared) dk = __dmul_rn(__dmul_rn(dk, dk), 1.0 + 1e-15);
            flag = !(dk < lim);
            // threshold-admitted candidates can be fewer than k (e.g. massive exact ties): escalate
            if ((uint64_t)a.kk <= a.n ? ncand < a.kk : false) flag = 1;
        }
        if (kout == 0 && a.kk > 0 && a.n > 0) flag = 1;
        if (a.out_flags) a.out_flags[b] = flag;
    }
    SEL_MARK(6);
}

template <int DTYPE, int THREADS>
__global__ void __launch_bounds__(THREADS) select_kernel(const SelectArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();   // last link of a search's kernel chain (common.cuh): the candidates come from the scan / GEMM before it
    select_body<DTYPE, THREADS>(a, blockIdx.x, smem);
}

// ============================================================================
// Small stores, a handful of queries: query preparation + scan + selection in ONE launch.
// The reference's everyday case (erlvectordb:search on ~10 k rows, BASELINE configs[0]) is bound by
// launch latency, not by bytes: three dependent kernels cost more than their work.  Every CTA narrows
// the query itself (d numbers: trivial), scans its rows exactly as scan_float_kernel does, publishes
// its KP-key list, and the CTA that arrives LAST for a query (one atomic counter per query) runs the
// whole selection -- merge, exact fp64 re-rank, order, proof (select_body) -- on the lists of all CTAs.
// ============================================================================
struct FusedArgs {
    SelectArgs sel;            // partial = the per-CTA lists this kernel writes; eps_q = where the per-query bound goes
    const float *inv_norm;     // row 1/norm (cosine)
    uint64_t *partial;         // [B][G][KP]
    float *eps_q;              // [B]
    unsigned int *arrive;      // [B] zero between launches
    int nch, tpr, G;
};

template <int METRIC, int DTYPE>
__global__ void __launch_bounds__(256) small_fused_kernel(const FusedArgs f) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ double s_red[8];
    __shared__ bool s_last;
    constexpr int QPC = (DTYPE == EVDB_F32) ? 1 : 2;
    constexpr int R = 4;
    const SelectArgs &a = f.sel;
    const int b = blockIdx.y, d = a.d, nch = f.nch, KP = a.KP, TPR = f.tpr, GPW = 32 / f.tpr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sqf = reinterpret_cast<float *>(smem);                                   // [nch*QPC*4] the query, fp32, zero padded
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + (size_t)nch * QPC * 16);
    // ---- 1. the query: fp32 copy, norm, and the narrowing residual that bounds the scan's score error ----
    const double *q = a.q64 + (size_t)b * d;
    double ss = 0.0, r2 = 0.0, r1 = 0.0;
    for (int i = threadIdx.x; i < nch * QPC * 4; i += blockDim.x) {
        const double v = i < d ? q[i] : 0.0;
        const float fv = (float)v;
        sqf[i] = fv;
        ss += v * v;
        const double df = fabs(v - (double)fv);
        r2 += df * df;
        r1 += df;
    }
    auto block_sum = [&](double v) -> double {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s_red[i];
        return t;
    };
    ss = block_sum(ss);
    float eps_q = 0.f;
    if (METRIC != EVDB_COSINE) {
        r2 = block_sum(r2);
        r1 = block_sum(r1);
        eps_q = METRIC == EVDB_EUCLIDEAN ? (float)(sqrt(r2) * 1.0000002) : (float)(r1 * 1.0000002);
        if (eps_q > 0.f) eps_q = nextafterf(eps_q, __int_as_float(0x7f800000));
    }
    if (threadIdx.x == 0) f.eps_q[b] = eps_q;      // every CTA writes the same value
    const float q_inv = ss > 0.0 ? (float)(1.0 / sqrt(ss)) : 0.0f;
    __syncthreads();
    // ---- 2. the scan (scan_float_kernel with a run-time lane group) ----
    const float4 *sq = reinterpret_cast<const float4 *>(sqf);
    WarpCands wc;
    wc.init(lists, KP, warp);
    const int g = lane / TPR, gl = lane % TPR;
    const uint64_t rows_per_wi = (uint64_t)GPW * R;
    const uint64_t total_wi = (a.n + rows_per_wi - 1) / rows_per_wi;
    for (uint64_t wi = (uint64_t)blockIdx.x * kScanWarps + warp; wi < total_wi; wi += (uint64_t)gridDim.x * kScanWarps) {
        const uint64_t base = wi * rows_per_wi;
        const uint4 *rp[R];
        uint64_t rix[R];
        bool valid[R];
        float inv[R];
        float4 acc[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint64_t r = base + (uint64_t)j * GPW + g;
            valid[j] = r < a.n;
            rix[j] = r;
            rp[j] = reinterpret_cast<const uint4 *>(a.rows + (valid[j] ? r : a.n - 1) * a.row_bytes);
            inv[j] = (METRIC == EVDB_COSINE && valid[j] && gl == 0) ? __ldg(f.inv_norm + rix[j]) : 0.f;
            acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 2
        for (int c = gl; c < nch; c += TPR) {
            uint4 v[R];
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = ldg_stream_u4(rp[j] + c);
            if (DTYPE == EVDB_F32) {
                const float4 q0 = sq[c];
#pragma unroll
                for (int j = 0; j < R; ++j)
                    acc_f4<METRIC>(acc[j], make_float4(__uint_as_float(v[j].x), __uint_as_float(v[j].y),
                                                       __uint_as_float(v[j].z), __uint_as_float(v[j].w)), q0);
            } else {
                const float4 q0 = sq[2 * c], q1 = sq[2 * c + 1];
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    acc_f4<METRIC>(acc[j], bf16x4_lo(v[j]), q0);
                    acc_f4<METRIC>(acc[j], bf16x4_hi(v[j]), q1);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            float sacc = (acc[j].x + acc[j].y) + (acc[j].z + acc[j].w);
            for (int o = TPR >> 1; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            float score;
            if (METRIC == EVDB_COSINE) score = (inv[j] == 0.f || q_inv == 0.f) ? 1.0f : 1.0f - sacc * inv[j] * q_inv;
            else if (METRIC == EVDB_EUCLIDEAN) score = sqrtf(sacc);
            else score = sacc;
            const uint64_t key = (valid[j] && gl == 0) ? make_key(score, (uint32_t)rix[j]) : kKeyMax;
            wc.offer(key, lane);
        }
    }
    wc.finish(lists, warp, lane);
    cta_merge_and_store(lists, KP, f.partial + ((size_t)b * f.G + blockIdx.x) * KP);
    // ---- 3. the last CTA of this query selects ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(f.arrive + b, 1u) == (unsigned)f.G - 1u;
        if (s_last) f.arrive[b] = 0;       // ready for the next launch (every CTA of this query has arrived)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    select_body<DTYPE, 256>(a, b, smem);
}

// ============================================================================
// Warp-per-query select for the GEMM plan's query batches: no block barriers anywhere.
//   gather the query's candidate buffers into shared memory (passes of up to 2048 keys),
//   keep the KP best by bisection with warp population counts, order them, then re-rank:
//   the warp stages 128-element chunks of up to 32 candidate rows in shared memory with
//   coalesced loads, and every lane folds ONE row from there in the reference's strict
//   left-to-right fp64 order (lane 31 folds the query's own squares for cosine).
// The key buffer and the row staging area share storage: they are never live together.
// ============================================================================
constexpr int kSwWarps = 8;        // queries per CTA: 8 consecutive queries' keys share 64 contiguous bytes of every buffer row
constexpr int kSwKeys = 2048;      // keys per selection pass
constexpr int kSwKC = 64;          // elements per staged row chunk
constexpr int kSwMaxKP = 128;

constexpr int kSwRows = 32;                               // staged product rows per round (candidates + the query's own squares): one per lane
constexpr int kSwProdStride = kSwKC * 8 + 16;             // bytes between staged product rows (bank spread)
struct SwLayout {
    static constexpr int kUnion = (kSwRows * kSwProdStride > kSwKeys * 8) ? kSwRows * kSwProdStride : kSwKeys * 8;
    static constexpr int kPerWarp = kUnion + kSwMaxKP * 8 * 3;  // + ckeys/dkey/dslot
};

// ascending bitonic sort of n (power of two) u64 keys in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_smem(uint64_t *k, int n, int lane) {
    for (int k2 = 2; k2 <= n; k2 <<= 1)
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (int i = lane; i < n; i += 32) {
                const int ixj = i ^ j2;
                if (ixj > i) {
                    const uint64_t x = k[i], y = k[ixj];
                    if ((x > y) == ((i & k2) == 0)) { k[i] = y; k[ixj] = x; }
                }
            }
            __syncwarp();
        }
}
__device__ __forceinline__ void warp_bitonic_smem_pairs(uint64_t *k, uint64_t *v, int n, int lane) {
    for (int k2 = 2; k2 <= n; k2 <<= 1)
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (int i = lane; i < n; i += 32) {
                const int ixj = i ^ j2;
                if (ixj > i) {
                    const uint64_t x = k[i], y = k[ixj], vx = v[i], vy = v[ixj];
                    const bool gt = x > y || (x == y && vx > vy);
                    if (gt == ((i & k2) == 0)) { k[i] = y; k[ixj] = x; v[i] = vy; v[ixj] = vx; }
                }
            }
            __syncwarp();
        }
}

// keep the `need` smallest of keys[0, T) (need < T), compacted in place to keys[0, need)
// Returns the need-th smallest of the kSub = 32 * VPL scores a warp holds in registers (VPL per lane).
template <int VPL>
__device__ __forceinline__ uint32_t warp_nth_of_regs(const uint32_t (&v)[VPL], int need) {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { mn = min(mn, v[i]); mx = max(mx, v[i]); }
    uint32_t lo = __reduce_min_sync(0xffffffffu, mn), hi = __reduce_max_sync(0xffffffffu, mx);
    while (lo < hi) {  // invariant: count(v <= hi) >= need > count(v < lo)
        const uint32_t span = hi - lo, qd = span >> 2;
        const uint32_t p2 = lo + (span >> 1), p1 = qd ? lo + qd : p2, p3 = qd ? p2 + qd : p2;
        int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            c1 += v[i] <= p1 ? 1 : 0;
            c2 += v[i] <= p2 ? 1 : 0;
            c3 += v[i] <= p3 ? 1 : 0;
        }
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        if (c1 >= need) hi = p1;
        else if (c2 >= need) { lo = p1 + 1; hi = p2; }
        else if (c3 >= need) { lo = p2 + 1; hi = p3; }
        else lo = p3 + 1;
    }
    return lo;
}

// Many more keys than wanted (the GEMM plan leaves ~30 x KP per query): the need-th smallest score of a
// strided SUBSET bounds the need-th smallest overall, so only keys at or below it can matter --
// about T * need / subset of them.  The subset lives in registers (no shared-memory passes per
// bisection round); one compaction pass, then the exact selection runs on what is left.
template <int VPL>                               // 32 * VPL-key subset
__device__ __forceinline__ int warp_precut_inplace(uint64_t *keys, int T, int need, int lane) {
    const uint32_t *hi32 = reinterpret_cast<const uint32_t *>(keys) + 1;
    const int stride = T / (32 * VPL);
    uint32_t v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = hi32[2 * ((i * 32 + lane) * stride)];
    const uint32_t cut = warp_nth_of_regs<VPL>(v, need);
    int out = 0;
    const unsigned below = (1u << lane) - 1u;
    for (int base = 0; base < T; base += 32) {    // in order: the write cursor never passes the read cursor
        const int i = base + lane;
        const uint64_t key = i < T ? keys[i] : kKeyMax;
        const bool keep = i < T && (uint32_t)(key >> 32) <= cut;
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) keys[out + __popc(mk & below)] = key;
        out += __popc(mk);
        __syncwarp();
    }
    return out;   // >= need
}

__device__ __forceinline__ void warp_select_inplace(uint64_t *keys, int T, int need, int lane) {
    if (need <= 64 && T >= 8 * need && T >= 512) {
        T = warp_precut_inplace<8>(keys, T, need, lane);
        if (T == need) return;
    } else if (need <= 128 && T >= 8 * need && T >= 1024) {      // k = 100 windows: a 512-key subset (stride >= 2)
        T = warp_precut_inplace<16>(keys, T, need, lane);
        if (T == need) return;
    }
    const uint32_t *hi32 = reinterpret_cast<const uint32_t *>(keys) + 1;  // score word of key i at hi32[2*i]
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (int i = lane; i < T; i += 32) {
        const uint32_t sc = hi32[2 * i];
        mn = min(mn, sc);
        mx = max(mx, sc);
    }
    uint32_t lo = __reduce_min_sync(0xffffffffu, mn), hi = __reduce_max_sync(0xffffffffu, mx);
    while (lo < hi) {  // invariant: count(score <= hi) >= need > count(score < lo)
        const uint32_t span = hi - lo, qd = span >> 2;
        const uint32_t p2 = lo + (span >> 1), p1 = qd ? lo + qd : p2, p3 = qd ? p2 + qd : p2;
        int c1 = 0, c2 = 0, c3 = 0;
        for (int i = lane; i < T; i += 32) {
            const uint32_t sc = hi32[2 * i];
            c1 += sc <= p1 ? 1 : 0;
            c2 += sc <= p2 ? 1 : 0;
            c3 += sc <= p3 ? 1 : 0;
        }
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        if (c1 >= need) hi = p1;
        else if (c2 >= need) { lo = p1 + 1; hi = p2; }
        else if (c3 >= need) { lo = p2 + 1; hi = p3; }
        else lo = p3 + 1;
    }
    int nl = 0;
    for (int i = lane; i < T; i += 32) nl += hi32[2 * i] < lo ? 1 : 0;
    int ties_left = need - __reduce_add_sync(0xffffffffu, nl);
    // one in-order pass: everything below the cut, and ties at the cut while the budget lasts;
    // the write cursor never passes the read cursor
    int out = 0;
    const unsigned below = (1u << lane) - 1u;
    for (int base = 0; base < T; base += 32) {
        const int i = base + lane;
        const uint64_t key = i < T ? keys[i] : kKeyMax;
        const uint32_t sc = (uint32_t)(key >> 32);
        const bool tie = i < T && sc == lo;
        const unsigned mt = __ballot_sync(0xffffffffu, tie);
        const bool keep = i < T && (sc < lo || (tie && __popc(mt & below) < ties_left));
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        ties_left -= min(__popc(mt), max(ties_left, 0));
        __syncwarp();
        if (keep) keys[out + __popc(mk & below)] = key;
        out += __popc(mk);
        __syncwarp();
    }
}

// ---- the exact fp64 fold by a GROUP of warps (one query): a leader and kMwProducers producer warps ----
// One warp per query (round 1) is bound by its own instruction latency: ~500 instructions per 64-element chunk,
// each waiting ~5 cycles for the previous one, 12 chunks at d = 768 -- 89 k cycles for a chain whose
// dependent DADDs need 6.6 k (r02 EVDB_SEL_VARIANT=16 counters).  Here the producer warps form the
// products of chunk c (two rows per warp instruction, row pairs dealt round-robin) into one half of
// a double-buffered stage WHILE the leader's lanes fold chunk c-1 out of the other half, one group
// barrier (bar.sync on the group's own id) per chunk.  At most kMwRows rows are staged per round.
constexpr int kMwProducers = 3;
constexpr int kMwGroup = 1 + kMwProducers;         // warps per query
constexpr int kMwRows = 16;                        // staged rows per round: 2 buffers x 16 rows == the single-warp stage
constexpr int kMwIt = (kMwRows / 2 + kMwProducers - 1) / kMwProducers;   // row pairs per producer warp and chunk
static_assert(2 * kMwRows * kSwProdStride <= SwLayout::kUnion, "the double-buffered stage must fit the key buffer it reuses");

__device__ __forceinline__ void group_bar(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kMwGroup * 32) : "memory");
}

// role 0: the leader (lane r < nrows folds row r and gets its sum back); role 1..kMwProducers: producers.
// slots: shared memory, the nr row slots of this round (written by the leader before the group's barrier).
template <int METRIC>
__device__ __forceinline__ double mw_fold_m(const uint8_t *__restrict__ rows, size_t row_bytes, const double *__restrict__ q,
                                            int d, uint8_t *stage, const uint32_t *slots, int nr, bool qrow,
                                            int role, int lane, int bar_id) {
    constexpr int metric = METRIC;
    const int nrows = nr + (qrow ? 1 : 0);
    const int npairs = (nrows + 1) >> 1;
    const int unit = lane & 15, rsub = lane >> 4;   // a row chunk is 16 units of 4 elements: half a warp per row
    const int nchunks = (d + kSwKC - 1) / kSwKC;
    const int nu = (int)(row_bytes >> 4);          // 16-byte units per stored row
    // producers: the raw rows of one chunk into registers
    auto load_raw = [&](int kb, uint4 (&raw)[kMwIt]) {
        const int u = (kb >> 2) + unit;
#pragma unroll
        for (int j = 0; j < kMwIt; ++j) {
            const int it = (role - 1) + kMwProducers * j, r = 2 * it + rsub;
            raw[j] = make_uint4(0u, 0u, 0u, 0u);
            if (it < npairs && r < nr && u < nu)
                raw[j] = __ldg(reinterpret_cast<const uint4 *>(rows + (size_t)slots[r] * row_bytes) + u);
        }
    };
    auto produce = [&](int c, const uint4 (&raw)[kMwIt]) {
        uint8_t *st = stage + (size_t)(c & 1) * kMwRows * kSwProdStride;
        const int e0 = c * kSwKC + 4 * unit;
        double qd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) qd[i] = e0 + i < d ? __ldg(q + e0 + i) : 0.0;   // L1 hits: the group shares q
#pragma unroll
        for (int j = 0; j < kMwIt; ++j) {
            const int it = (role - 1) + kMwProducers * j, r = 2 * it + rsub;
            if (it >= npairs || r >= nrows) continue;
            const uint32_t w[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
            double t[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // F2F.F64.F32 (exact, subnormals included): one issue slot per element; the conversion unit's
                // 16 lanes/clk/SM (tools/micro/cvt_bench.cu: 2.0 cycles per warp conversion, the same as the
                // ~8-instruction integer widening) sit idle otherwise
                const double x = r < nr ? (double)__uint_as_float(w[i]) : qd[i];   // the last row: the query's own squares
                if (metric == EVDB_COSINE) t[i] = __dmul_rn(qd[i], x);
                else {
                    const double df = __dsub_rn(qd[i], x);
                    t[i] = metric == EVDB_EUCLIDEAN ? __dmul_rn(df, df) : fabs(df);
                }
            }
            double2 *dst = reinterpret_cast<double2 *>(st + (size_t)r * kSwProdStride) + 2 * unit;
            dst[0] = make_double2(t[0], t[1]);
            dst[1] = make_double2(t[2], t[3]);
        }
    };
    auto fold = [&](int c, double s) -> double {   // leader: lane r adds the staged terms of chunk c to row r's sum, left to right
        const int kb = c * kSwKC;
        const int cnt = d - kb < kSwKC ? d - kb : kSwKC;
        const double2 *p = reinterpret_cast<const double2 *>(stage + (size_t)(c & 1) * kMwRows * kSwProdStride +
                                                             (size_t)lane * kSwProdStride);
        const int pairs = cnt >> 1;
        // the chain waits 8.6 cycles per DADD whatever else happens: the staged terms of the NEXT four
        // pairs are requested before the current four are added, so no shared-memory latency joins it
        double2 v0 = make_double2(0.0, 0.0), v1 = v0, v2 = v0, v3 = v0;
        if (pairs >= 4) { v0 = p[0]; v1 = p[1]; v2 = p[2]; v3 = p[3]; }
        int i = 0;
        for (; i + 4 <= pairs; i += 4) {
            double2 n0 = v0, n1 = v1, n2 = v2, n3 = v3;
            if (i + 8 <= pairs) { n0 = p[i + 4]; n1 = p[i + 5]; n2 = p[i + 6]; n3 = p[i + 7]; }
            s = __dadd_rn(s, v0.x); s = __dadd_rn(s, v0.y);
            s = __dadd_rn(s, v1.x); s = __dadd_rn(s, v1.y);
            s = __dadd_rn(s, v2.x); s = __dadd_rn(s, v2.y);
            s = __dadd_rn(s, v3.x); s = __dadd_rn(s, v3.y);
            v0 = n0; v1 = n1; v2 = n2; v3 = n3;
        }
        for (; i < pairs; ++i) {
            const double2 v = p[i];
            s = __dadd_rn(s, v.x);
            s = __dadd_rn(s, v.y);
        }
        if (cnt & 1) s = __dadd_rn(s, reinterpret_cast<const double *>(p)[cnt - 1]);
        return s;
    };
    double s = 0.0;
    if (role > 0) {
        // (two register sets -- the loads of chunk c+1 issued before chunk c is worked on -- were tried with
        // 2 producers per query: the kernel then needs > 80 registers at 768 threads and spills; slower)
        uint4 ra[kMwIt];
        load_raw(0, ra);
        for (int c = 0; c <= nchunks; ++c) {
            if (c < nchunks) produce(c, ra);
            if (c + 1 < nchunks) load_raw((c + 1) * kSwKC, ra);   // in flight across the barrier
            group_bar(bar_id);
        }
    } else {
        for (int c = 0; c <= nchunks; ++c) {
            if (c > 0 && lane < nrows) s = fold(c - 1, s);
            group_bar(bar_id);
        }
    }
    return s;
}

__device__ __forceinline__ double mw_fold(const uint8_t *__restrict__ rows, size_t row_bytes, const double *__restrict__ q,
                                          int d, int metric, uint8_t *stage, const uint32_t *slots, int nr, bool qrow,
                                          int role, int lane, int bar_id) {
    if (metric == EVDB_COSINE) return mw_fold_m<EVDB_COSINE>(rows, row_bytes, q, d, stage, slots, nr, qrow, role, lane, bar_id);
    if (metric == EVDB_EUCLIDEAN) return mw_fold_m<EVDB_EUCLIDEAN>(rows, row_bytes, q, d, stage, slots, nr, qrow, role, lane, bar_id);
    return mw_fold_m<EVDB_MANHATTAN>(rows, row_bytes, q, d, stage, slots, nr, qrow, role, lane, bar_id);
}

// Producer side of a query group: wait for the leader's rounds until it posts nr == 0.
__device__ __forceinline__ void mw_produce(const uint8_t *rows, size_t row_bytes, const double *q, int d, int metric,
                                           uint8_t *stage, const uint32_t *slots, const volatile int *ctl, bool qrow,
                                           int role, int lane, int bar_id) {
    while (true) {
        group_bar(bar_id);
        const int nr = ctl[0];
        if (nr <= 0) return;
        mw_fold(rows, row_bytes, q, d, metric, stage, slots, nr, qrow, role, lane, bar_id);
    }
}

// F32 stores only (the GEMM plan's precondition).  A CTA serves kSwWarps consecutive queries with
// kMwGroup warps each: warp w < kSwWarps leads query w (gather, selection, order, fold, proof), warps
// kSwWarps + kMwProducers*w + {0..} are its producers.
__global__ void __launch_bounds__(kSwWarps * kMwGroup * 32) select_warp_kernel(const SelectArgs a, int B) {
    using LY = SwLayout;
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();   // the candidate buffers come from gemm_topk_kernel
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int role = wid < kSwWarps ? 0 : 1 + (wid - kSwWarps) % kMwProducers;
    const int warp = wid < kSwWarps ? wid : (wid - kSwWarps) / kMwProducers;   // the query of the CTA this warp works for
    const int b0 = blockIdx.x * kSwWarps;
    const int b = b0 + warp;
    uint8_t *wbase = smem + (size_t)warp * LY::kPerWarp;
    // CTA-shared gather tables behind the per-warp regions
    const int L = a.L;
    uint16_t *s_cnt = reinterpret_cast<uint16_t *>(smem + (size_t)kSwWarps * LY::kPerWarp);  // [L][8] fill counts
    int *s_off = reinterpret_cast<int *>(s_cnt + (size_t)(L + 1) * kSwWarps);                // [L+1][8] exclusive prefix of each query's counts
    int *s_moff = s_off + (size_t)(L + 1) * kSwWarps;                                         // [L+1] prefix of max-over-queries counts
    __shared__ int s_car[kSwWarps];
    __shared__ uint32_t s_slots[kSwWarps][kMwRows];   // the rows of the current fold round, per query group
    __shared__ int s_nr[kSwWarps];
    // ---- 1. gather: the CTA's 8 consecutive queries together ----
    // The GEMM epilogue leaves thread t's i-th key of a buffer at [i][t]: 8 consecutive queries'
    // keys are 64 contiguous bytes, so one pair of sectors serves all 8 warps (reading each query
    // alone would use 8 of every 32 bytes fetched and open a DRAM page per key).
    const RawCands &rw = a.raw;
    const int blk = b0 / rw.gm, et0 = b0 % rw.gm;   // b0 is a multiple of 8, gm of 8: one block, one sweep
    const int c = blk / rw.MB, mb_local = blk % rw.MB;
    auto list_index = [&](int l) -> size_t {
        const int ng = l / rw.parts, part = l % rw.parts;
        return ((size_t)c * rw.nCTA + (size_t)ng * rw.MB + mb_local) * rw.parts + part;
    };
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const int4 *p = reinterpret_cast<const int4 *>(rw.cnt + list_index(l) * rw.gm + et0);
        const int4 c0 = __ldcg(p), c1 = __ldcg(p + 1);
        const int v[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        int mx = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int n = b0 + w < B ? v[w] : 0;
            s_cnt[l * 8 + w] = (uint16_t)n;
            mx = n > mx ? n : mx;
        }
        s_moff[l + 1] = mx;
    }
    if (threadIdx.x == 0) s_moff[0] = 0;
    __syncthreads();
    if (role == 0) {   // leader w: exclusive prefix of its query's counts; warp 0 also the prefix of the row maxima
        int carry = 0;
        for (int l0 = 0; l0 < L; l0 += 32) {
            const int l = l0 + lane;
            const int n = l < L ? (int)s_cnt[l * 8 + warp] : 0;
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (l < L) s_off[l * 8 + warp] = carry + incl - n;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) { s_off[L * 8 + warp] = carry; s_car[warp] = 0; }
        if (warp == 0) {
            int mc = 0;
            for (int l0 = 1; l0 <= L; l0 += 32) {
                const int l = l0 + lane;
                int incl = l <= L ? s_moff[l] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += u;
                }
                if (l <= L) s_moff[l] = mc + incl;
                mc += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    __syncthreads();
    // Passes over ranges of buffers [l0, l1): as many whole buffers as fit next to every query's
    // carried keys (a buffer holds <= 256 keys, the carry <= 128, a pass 2048: progress is certain).
    uint64_t *keys = reinterpret_cast<uint64_t *>(wbase);              // selection passes
    const int KP = a.KP;
    int carried = 0, pos = 0;
    for (int l0 = 0; l0 < L;) {
        int lo = l0 + 1, hi = L;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            int worst = 0;
#pragma unroll
            for (int w = 0; w < kSwWarps; ++w) {
                const int f = s_car[w] + s_off[mid * 8 + w] - s_off[l0 * 8 + w];
                worst = f > worst ? f : worst;
            }
            if (worst <= kSwKeys) lo = mid; else hi = mid - 1;
        }
        const int l1 = lo;
        // one work item = one buffer row (l, i): 64 bytes = the 8 queries' i-th keys of buffer l
        for (int m = s_moff[l0] + threadIdx.x; m < s_moff[l1]; m += blockDim.x) {
            int a0 = l0, a1 = l1 - 1;  // largest l with s_moff[l] <= m
            while (a0 < a1) {
                const int mid = (a0 + a1 + 1) >> 1;
                if (s_moff[mid] <= m) a0 = mid; else a1 = mid - 1;
            }
            const int l = a0, i = m - s_moff[l];
            const uint4 *src = reinterpret_cast<const uint4 *>(rw.cand + (list_index(l) * rw.cap + i) * rw.gm + et0);
            const uint4 x0 = __ldcg(src), x1 = __ldcg(src + 1), x2 = __ldcg(src + 2), x3 = __ldcg(src + 3);
            const uint64_t kk8[8] = {((uint64_t)x0.y << 32) | x0.x, ((uint64_t)x0.w << 32) | x0.z,
                                     ((uint64_t)x1.y << 32) | x1.x, ((uint64_t)x1.w << 32) | x1.z,
                                     ((uint64_t)x2.y << 32) | x2.x, ((uint64_t)x2.w << 32) | x2.z,
                                     ((uint64_t)x3.y << 32) | x3.x, ((uint64_t)x3.w << 32) | x3.z};
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if (i < (int)s_cnt[l * 8 + w])
                    reinterpret_cast<uint64_t *>(smem + (size_t)w * LY::kPerWarp)[s_car[w] + s_off[l * 8 + w] - s_off[l0 * 8 + w] + i] = kk8[w];
        }
        __syncthreads();
        pos = carried + s_off[l1 * 8 + warp] - s_off[l0 * 8 + warp];
        if (l1 < L) {  // more to come: reduce to the KP best so the next range fits
            if (role == 0) {
                if (pos > KP) { warp_select_inplace(keys, pos, KP, lane); carried = KP; } else carried = pos;
                __syncwarp();
                if (lane == 0) s_car[warp] = carried;
            } else {
                carried = pos > KP ? KP : pos;
            }
        }
        __syncthreads();
        l0 = l1;
    }
    if (role > 0) {      // producers: nothing to do for a padding query or when only the window travels
        if (b >= B || a.win_mode) return;
        mw_produce(a.rows, a.row_bytes, a.q64 + (size_t)b * a.d, a.d, a.metric, wbase, s_slots[warp], &s_nr[warp],
                   a.metric == EVDB_COSINE, role, lane, 1 + warp);
        return;
    }
    if (b >= B) {        // (after the last block barrier)
        if (a.win_mode) push_arrive(a.push, gridDim.x * kSwWarps, lane);
        return;
    }
    uint8_t *stage = wbase;                                            // ... later: staged fp64 product rows
    uint64_t *ckeys = reinterpret_cast<uint64_t *>(wbase + LY::kUnion);  // [kSwMaxKP] the window, ascending
    uint64_t *dkey = ckeys + kSwMaxKP, *dslot = dkey + kSwMaxKP;
    const float eps_abs = a.eps_abs + (a.eps_q ? a.eps_q[b] : 0.f);
    long long t_mark = clock64();
#define SW_MARK(i) do { if ((a.variant & 16) && lane == 0) { long long _t = clock64(); atomicAdd(&g_sel_dbg[i], (unsigned long long)(_t - t_mark)); t_mark = _t; } } while (0)
    __syncwarp();
    SW_MARK(0);
    if ((a.variant & 16) && lane == 0) atomicAdd(&g_sel_dbg[6], (unsigned long long)pos);
    if (pos > KP) { warp_select_inplace(keys, pos, KP, lane); carried = KP; } else carried = pos;
    SW_MARK(1);
    int nk = 2;
    while (nk < carried) nk <<= 1;
    for (int i = carried + lane; i < nk; i += 32) keys[i] = kKeyMax;
    __syncwarp();
    warp_bitonic_smem(keys, nk, lane);
    const int ncand = carried;
    for (int i = lane; i < ncand; i += 32) ckeys[i] = keys[i];
    __syncwarp();
    const float bound = ncand > 0 ? key_score(ckeys[ncand - 1]) : 0.f;
    if (!a.win_mode) {
        // the rows most likely to be re-ranked start their trip from DRAM now (L2 prefetch): the fold
        // below walks them 256 bytes at a time
        const int npf = ncand < a.kk + 6 ? ncand : a.kk + 6;
        const int lines = (int)((a.row_bytes + 127) >> 7);
        for (int i = lane; i < npf * lines; i += 32) {
            const int j = i / lines, l = i - j * lines;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rows + (size_t)key_slot(ckeys[j]) * a.row_bytes + (size_t)l * 128));
        }
    }
    if (a.win_mode) {  // sharded search, phase 1: the window travels, the re-rank happens after the global merge
        for (int i = lane; i < KP; i += 32) push_store(a.push, (size_t)b * KP + i, i < ncand ? (ckeys[i] & 0xFFFFFFFF00000000ull) | (a.slot_base + (ckeys[i] & 0xFFFFFFFFull) * a.slot_mul) : kKeyMax);
        if (lane == 0) push_store(a.push, (size_t)B * KP + b, ((uint64_t)__float_as_uint(eps_abs) << 32) | (uint32_t)ncand);
        push_arrive(a.push, gridDim.x * kSwWarps, lane);
        return;
    }

    // ---- 2. which candidates can still reach the top k (a prefix of the ascending window) ----
    const int kout = a.kk < ncand ? a.kk : ncand;
    int nrer = kout;
    if (kout > 0) {
        const float sk = key_score(ckeys[kout - 1]);
        const float lim = sk + 2.0f * (eps_abs + a.eps_rel * fabsf(sk)) * 1.0001f;
        int cntl = 0;
        for (int i = kout + lane; i < ncand; i += 32) cntl += key_score(ckeys[i]) <= lim ? 1 : 0;
        nrer = kout + __reduce_add_sync(0xffffffffu, cntl);
    }

    // ---- 3. exact distances: rows staged chunk by chunk, one lane folds one row ----
    SW_MARK(2);
    if ((a.variant & 16) && lane == 0) atomicAdd(&g_sel_dbg[7], (unsigned long long)nrer);
    // The independent terms (q*v, (q-v)^2, |q-v|) are formed by all 32 lanes in parallel -- 32 useful
    // fp64 multiplies per instruction -- and staged as fp64 in shared memory; then lane r folds row r
    // strictly left to right.  For cosine one more staged row holds q*q (vector_norm(Query)).
    const double *q = a.q64 + (size_t)b * a.d;
    const bool cosine = a.metric == EVDB_COSINE;
    const int RC = cosine ? kMwRows - 1 : kMwRows;   // candidates per round
    for (int base = 0; base < nrer; base += RC) {
        const int nr = nrer - base < RC ? nrer - base : RC;
        const bool mine = lane < nr;
        const uint32_t slot = mine ? key_slot(ckeys[base + lane]) : 0u;
        if (mine) s_slots[warp][lane] = slot;
        if (lane == 0) s_nr[warp] = nr;
        __syncwarp();
        group_bar(1 + warp);
        const double s = mw_fold(a.rows, a.row_bytes, q, a.d, a.metric, stage, s_slots[warp], nr, cosine, 0, lane, 1 + warp);
        double dist;
        if (cosine) {
            const double sq = __shfl_sync(0xffffffffu, s, nr);
            const double n1 = __dsqrt_rn(sq), n2 = mine ? a.norm64[slot] : 0.0;
            dist = (n1 == 0.0 || n2 == 0.0) ? 1.0 : __dsub_rn(1.0, __ddiv_rn(s, __dmul_rn(n1, n2)));
        } else if (a.metric == EVDB_EUCLIDEAN) {
            dist = __dsqrt_rn(s);
        } else {
            dist = s;
        }
        if (mine) {
            dkey[base + lane] = f64_orderable(dist);
            dslot[base + lane] = slot;
        }
    }
    if (lane == 0) s_nr[warp] = 0;   // no more rounds: the producers leave
    __syncwarp();
    group_bar(1 + warp);
    int nsort = 2;
    while (nsort < nrer) nsort <<= 1;
    for (int j = nrer + lane; j < nsort; j += 32) { dkey[j] = kKeyMax; dslot[j] = kKeyMax; }
    __syncwarp();
    SW_MARK(3);

    // ---- 4. final order by (exact distance, slot); emit k; completeness proof ----
    warp_bitonic_smem_pairs(dkey, dslot, nsort, lane);
    for (int i = lane; i < a.kstride; i += 32) {
        const size_t o = (size_t)b * a.kstride + i;
        if (i < kout) {
            const uint64_t ob = dkey[i];
            const uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            a.out_dists[o] = __longlong_as_double((long long)bits);
            a.out_ids[o] = a.slot_base + dslot[i] * a.slot_mul;
        } else {
            a.out_dists[o] = 0.0;
            a.out_ids[o] = kKeyMax;
        }
    }
    if (lane == 0) {
        a.out_counts[b] = kout;
        int flag = 0;
        if (kout > 0 && (uint64_t)ncand < a.n) {  // rows exist outside the window
            const uint64_t ob = dkey[kout - 1];
            const uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            double dk = __longlong_as_double((long long)bits);
            const double lim = (double)bound - (double)eps_abs - (double)a.eps_rel * fabs((double)bound);
            if (a.squared) dk = __dmul_rn(__dmul_rn(dk, dk), 1.0 + 1e-15);
            flag = !(dk < lim);
            if ((uint64_t)a.kk <= a.n ? ncand < a.kk : false) flag = 1;
        }
        if (kout == 0 && a.kk > 0 && a.n > 0) flag = 1;
        if (a.out_flags) a.out_flags[b] = flag;
    }
    SW_MARK(4);
}

static int launch_select_warp(const SelectArgs &a, int B, cudaStream_t st) {
    const size_t smem = (size_t)kSwWarps * SwLayout::kPerWarp + (size_t)(a.L + 1) * kSwWarps * 6 + (size_t)(a.L + 1) * 4 + 16;
    if (smem > 220 * 1024) return EVDB_E_UNSUPPORTED;
    EVDB_TRY(ensure_func_smem((const void *)select_warp_kernel, smem));
    EVDB_CUDA(launch_chained(select_warp_kernel, dim3((B + kSwWarps - 1) / kSwWarps), dim3(kSwWarps * kMwGroup * 32), smem, st, 1, a, B));
    if (a.variant & 16) {
        cudaStreamSynchronize(st);
        unsigned long long h[10], z[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyFromSymbol(h, g_sel_dbg, sizeof(h));
        cudaMemcpyToSymbol(g_sel_dbg, z, sizeof(z));
        fprintf(stderr, "[select-warp dbg] cycles/query: gather(slow path)=%.0f select=%.0f sort+window=%.0f fold=%.0f final=%.0f  keys=%.0f nrer=%.1f\n",
                (double)h[0] / B, (double)h[1] / B, (double)h[2] / B, (double)h[3] / B, (double)h[4] / B, (double)h[6] / B, (double)h[7] / B);
    }
    return EVDB_OK;
}

// ============================================================================
// Row-sharded GEMM batches, two phases.  Re-ranking every shard's local top-k would
// repeat the exact fp64 work on every rank; instead the APPROXIMATE windows travel first:
//   phase 1  (select_warp_kernel, win_out mode): each rank's ascending window of KP keys
//            (score, global row) + its size and error bound -> pushed to every rank;
//   phase 2  (shard_rerank_kernel): every rank merges the world windows into the same global
//            window, derives the same completeness bound, and re-ranks in exact fp64 ONLY the
//            candidates it owns (1/world of them) -> exact distances pushed to every rank;
//   phase 3  (shard_final_kernel): every rank orders the global candidates by the owners' exact
//            distances, emits the top k and proves the window complete.
// Bound: a row outside the global window was either cut by the merge (score >= the window's last
// score) or never entered its shard's window (score >= that shard's last key, which only matters
// for a shard whose window does not cover all of its rows).
// ============================================================================
struct GMeta { int nrer, kout, flag, has_outside; float bound, eps; };

struct ShardArgs {
    const uint8_t *rows; size_t row_bytes; const double *norm64; int d; const double *q64;
    int B, KP, kk, metric, squared, world, rank;
    int strided;          // 0: rank r owns the contiguous block shard_range(r); 1: rank r owns the rows == r (mod world)
    uint64_t n_total;
    ExchangeView win;     // phase 2 input: per rank [B*KP keys][B meta]
    PushTarget e_push;    // phase 2 output: [B][KP] exact distances of the rows this rank owns (0 elsewhere), pushed to every rank
    uint64_t *g_out;      // [B][KP] the global window (each rank's own copy, identical everywhere)
    GMeta *g_meta;        // [B]
    ExchangeView ex;      // phase 3 input: per rank [B][KP] exact distances
    uint64_t *out_blob;   // packed result: [B*k ids][B*k dists][B counts i32][B flags i32]
    int k;
};

__device__ __forceinline__ void shard_range(uint64_t n, int world, int r, uint64_t *lo, uint64_t *hi) {
    const uint64_t per = (n + world - 1) / world;
    *lo = (uint64_t)r * per < n ? (uint64_t)r * per : n;
    *hi = *lo + per < n ? *lo + per : n;
}
// rows of the whole store that live on rank r
__device__ __forceinline__ uint64_t shard_rows(uint64_t n, int world, int r, int strided) {
    if (strided) return n > (uint64_t)r ? (n - r + world - 1) / world : 0;
    uint64_t lo, hi;
    shard_range(n, world, r, &lo, &hi);
    return hi - lo;
}

__device__ __forceinline__ void wait_flags(const unsigned long long *flags, unsigned long long epoch, int world, int lane) {
    if (lane < world) {
        const volatile unsigned long long *f = flags + lane;
        const long long t0 = clock64();
        while (*f < epoch)
            if (clock64() - t0 > 8000000000ll) __trap();   // a dead peer must not hang the GPU
    }
    __syncwarp();
    __threadfence_system();
}

constexpr int kShWarps = 4;
constexpr int kShPerWarp = SwLayout::kUnion + kSwMaxKP * 8 * 3;   // keys / product staging + window copy + owned list + exact values

// kShWarps queries per CTA, kMwGroup warps each (a leader + the fold's producers, see mw_fold)
__global__ void __launch_bounds__(kShWarps * kMwGroup * 32) shard_rerank_kernel(const ShardArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_slots[kShWarps][kMwRows];
    __shared__ int s_nr[kShWarps];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int role = wid < kShWarps ? 0 : 1 + (wid - kShWarps) % kMwProducers;
    const int warp = wid < kShWarps ? wid : (wid - kShWarps) / kMwProducers;
    const int b = blockIdx.x * kShWarps + warp;
    if (role > 0) {
        if (b < a.B)
            mw_produce(a.rows, a.row_bytes, a.q64 + (size_t)b * a.d, a.d, a.metric, smem + (size_t)warp * kShPerWarp,
                       s_slots[warp], &s_nr[warp], a.metric == EVDB_COSINE, role, lane, 1 + warp);
        return;
    }
    if (b >= a.B) {
        push_arrive(a.e_push, gridDim.x * kShWarps, lane);
        return;
    }
    uint8_t *wbase = smem + (size_t)warp * kShPerWarp;
    uint64_t *keys = reinterpret_cast<uint64_t *>(wbase);
    uint8_t *stage = wbase;
    uint64_t *ckeys = reinterpret_cast<uint64_t *>(wbase + SwLayout::kUnion);   // [KP] global window
    int *olist = reinterpret_cast<int *>(ckeys + kSwMaxKP);                      // [KP] window positions this rank owns
    double *evals = reinterpret_cast<double *>(ckeys + 2 * kSwMaxKP);            // [KP] exact distances of my rows, 0 elsewhere
    const int KP = a.KP;
    wait_flags(a.win.flags, a.win.epoch, a.world, lane);

    // ---- merge the world windows (lane r reads rank r's header; then every key load is independent) ----
    int flag = 0, has_outside = 0;
    float eps = 0.f, bstar = __int_as_float(0x7f800000);
    int nc = 0;
    if (lane < a.world) {
        const uint64_t *wr = a.win.slots + (size_t)lane * a.win.stride;
        const uint64_t meta = __ldcg(wr + (size_t)a.B * KP + b);
        nc = (int)(uint32_t)meta;
        eps = __uint_as_float((uint32_t)(meta >> 32));
        if ((uint64_t)nc < shard_rows(a.n_total, a.world, lane, a.strided)) {   // rows of shard r exist outside its window
            has_outside = 1;
            if (nc > 0) bstar = key_score(__ldcg(wr + (size_t)b * KP + nc - 1));
            else flag = 1;                     // nothing admitted there: no bound to offer
        }
    }
    int incl = nc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    const int T = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - nc;
    for (int o = 16; o > 0; o >>= 1) {
        eps = fmaxf(eps, __shfl_xor_sync(0xffffffffu, eps, o));
        bstar = fminf(bstar, __shfl_xor_sync(0xffffffffu, bstar, o));
        flag |= __shfl_xor_sync(0xffffffffu, flag, o);
        has_outside |= __shfl_xor_sync(0xffffffffu, has_outside, o);
    }
    for (int idx = lane; idx < a.world * KP; idx += 32) {
        const int r = idx / KP, j = idx - r * KP;
        const int nr_ = __shfl_sync(0xffffffffu, nc, r), off = __shfl_sync(0xffffffffu, excl, r);
        if (j < nr_) keys[off + j] = __ldcg(a.win.slots + (size_t)r * a.win.stride + (size_t)b * KP + j);
    }
    __syncwarp();
    int ng = T;
    if (T > KP) { warp_select_inplace(keys, T, KP, lane); ng = KP; }
    int nk = 2;
    while (nk < ng) nk <<= 1;
    for (int i = ng + lane; i < nk; i += 32) keys[i] = kKeyMax;
    __syncwarp();
    warp_bitonic_smem(keys, nk, lane);
    if (T > KP) { has_outside = 1; bstar = fminf(bstar, key_score(keys[KP - 1])); }
    for (int i = lane; i < KP; i += 32) {
        const uint64_t key = i < ng ? keys[i] : kKeyMax;
        ckeys[i] = key;
        a.g_out[(size_t)b * KP + i] = key;
        evals[i] = 0.0;
    }
    __syncwarp();

    // ---- which candidates can still reach the top k; which of those are mine ----
    const int kout = a.kk < ng ? a.kk : ng;
    int nrer = kout;
    if (kout > 0) {
        const float sk = key_score(ckeys[kout - 1]);
        const float lim = sk + 2.0f * eps * 1.0001f;
        int c = 0;
        for (int i = kout + lane; i < ng; i += 32) c += key_score(ckeys[i]) <= lim ? 1 : 0;
        nrer = kout + __reduce_add_sync(0xffffffffu, c);
    }
    if (ng < a.kk && has_outside) flag = 1;
    uint64_t mylo, myhi;
    shard_range(a.n_total, a.world, a.rank, &mylo, &myhi);
    int no = 0;
    for (int base = 0; base < nrer; base += 32) {
        const int j = base + lane;
        const uint64_t row = j < nrer ? (uint64_t)key_slot(ckeys[j]) : ~0ull;
        const bool own = j < nrer && (a.strided ? row % (uint64_t)a.world == (uint64_t)a.rank : row >= mylo && row < myhi);
        const unsigned m = __ballot_sync(0xffffffffu, own);
        if (own) olist[no + __popc(m & ((1u << lane) - 1u))] = j;
        no += __popc(m);
    }
    __syncwarp();
    if (lane == 0) {
        GMeta gm;
        gm.nrer = nrer; gm.kout = kout; gm.flag = flag; gm.has_outside = has_outside; gm.bound = bstar; gm.eps = eps;
        a.g_meta[b] = gm;
    }

    // ---- exact fp64 distances of my rows ----
    const double *q = a.q64 + (size_t)b * a.d;
    const bool cosine = a.metric == EVDB_COSINE;
    const int RC = cosine ? kMwRows - 1 : kMwRows;
    for (int base = 0; base < no; base += RC) {
        const int nr = no - base < RC ? no - base : RC;
        const bool mine = lane < nr;
        const int j = mine ? olist[base + lane] : 0;
        const uint64_t grow = mine ? (uint64_t)key_slot(ckeys[j]) : 0ull;
        const uint32_t slot = mine ? (uint32_t)(a.strided ? grow / (uint64_t)a.world : grow - mylo) : 0u;
        if (mine) s_slots[warp][lane] = slot;
        if (lane == 0) s_nr[warp] = nr;
        __syncwarp();
        group_bar(1 + warp);
        const double s = mw_fold(a.rows, a.row_bytes, q, a.d, a.metric, stage, s_slots[warp], nr, cosine, 0, lane, 1 + warp);
        double dist;
        if (cosine) {
            const double sq = __shfl_sync(0xffffffffu, s, nr);
            const double n1 = __dsqrt_rn(sq), n2 = mine ? a.norm64[slot] : 0.0;
            dist = (n1 == 0.0 || n2 == 0.0) ? 1.0 : __dsub_rn(1.0, __ddiv_rn(s, __dmul_rn(n1, n2)));
        } else if (a.metric == EVDB_EUCLIDEAN) {
            dist = __dsqrt_rn(s);
        } else {
            dist = s;
        }
        if (mine) evals[j] = dist;
    }
    if (lane == 0) s_nr[warp] = 0;   // the producers leave
    __syncwarp();
    group_bar(1 + warp);
    for (int i = lane; i < KP; i += 32) push_store(a.e_push, (size_t)b * KP + i, (uint64_t)__double_as_longlong(evals[i]));
    push_arrive(a.e_push, gridDim.x * kShWarps, lane);
}

__global__ void __launch_bounds__(kShWarps * 32) shard_final_kernel(const ShardArgs a) {
    __shared__ uint64_t s_key[kShWarps][kSwMaxKP], s_id[kShWarps][kSwMaxKP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kShWarps + warp;
    if (b >= a.B) return;
    wait_flags(a.ex.flags, a.ex.epoch, a.world, lane);
    const int KP = a.KP, k = a.k;
    const GMeta gm = a.g_meta[b];
    uint64_t *dk = s_key[warp], *di = s_id[warp];
    const uint64_t per = (a.n_total + a.world - 1) / a.world;
    int nsort = 2;
    while (nsort < gm.nrer) nsort <<= 1;
    for (int j = lane; j < nsort; j += 32) {
        uint64_t key = kKeyMax, id = kKeyMax;
        if (j < gm.nrer) {
            id = (uint64_t)key_slot(a.g_out[(size_t)b * KP + j]);
            const int owner = a.strided ? (int)(id % (uint64_t)a.world) : (int)(id / per);
            key = f64_orderable(__ldcg(reinterpret_cast<const double *>(a.ex.slots + (size_t)owner * a.ex.stride) +
                                       (size_t)b * KP + j));
        }
        dk[j] = key;
        di[j] = id;
    }
    __syncwarp();
    warp_bitonic_smem_pairs(dk, di, nsort, lane);
    uint64_t *out_ids = a.out_blob;
    double *out_d = reinterpret_cast<double *>(a.out_blob + (size_t)a.B * k);
    int32_t *out_c = reinterpret_cast<int32_t *>(a.out_blob + 2 * (size_t)a.B * k);
    for (int i = lane; i < k; i += 32) {
        const size_t o = (size_t)b * k + i;
        if (i < gm.kout) {
            const uint64_t ob = dk[i];
            const uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            out_d[o] = __longlong_as_double((long long)bits);
            out_ids[o] = di[i];
        } else {
            out_d[o] = 0.0;
            out_ids[o] = kKeyMax;
        }
    }
    if (lane == 0) {
        out_c[b] = gm.kout;
        int flag = gm.flag;
        if (gm.kout > 0 && gm.has_outside) {
            const uint64_t ob = dk[gm.kout - 1];
            const uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            double dkv = __longlong_as_double((long long)bits);
            if (a.squared) dkv = __dmul_rn(__dmul_rn(dkv, dkv), 1.0 + 1e-15);
            if (!(dkv < (double)gm.bound - (double)gm.eps)) flag = 1;
        }
        if (gm.kout == 0 && a.kk > 0) flag = 1;
        out_c[a.B + b] = flag;
    }
}

// phase 1: local window of every query -> win_blob ([B*KP keys][B meta])
int launch_shard_window(evdb_store *s, const double *d_q64, const RawCands *raw, int L, int KP, int B, int kk,
                        int metric, const float *eps_q, uint64_t slot_base, const PushTarget &push, cudaStream_t st) {
    if (!raw || s->dtype != EVDB_F32 || KP > kSwMaxKP || L > kRawMaxLists) return EVDB_E_UNSUPPORTED;
    SelectArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = s->rows; a.row_bytes = s->row_bytes; a.norm64 = s->norm64; a.qms64 = s->qms64;
    a.n = s->count; a.d = s->dim; a.q64 = d_q64; a.L = L; a.KP = KP; a.raw = *raw;
    a.kk = kk; a.kstride = kk; a.metric = metric; a.eps_q = eps_q; a.slot_base = slot_base; a.slot_mul = s->slot_mul;
    a.win_mode = 1; a.push = push;
    s->n_launches++;
    return launch_select_warp(a, B, st);
}

static void fill_shard_args(ShardArgs *a, evdb_store *s, const double *d_q64, int B, int KP, int k, int kk, int metric,
                            int rank, int world, uint64_t n_total, uint64_t *g_out, void *g_meta) {
    memset(a, 0, sizeof(*a));
    a->rows = s->rows; a->row_bytes = s->row_bytes; a->norm64 = s->norm64; a->d = s->dim; a->q64 = d_q64;
    a->B = B; a->KP = KP; a->kk = kk; a->k = k; a->metric = metric; a->squared = metric == EVDB_EUCLIDEAN;
    a->world = world; a->rank = rank; a->n_total = n_total; a->strided = s->slot_mul > 1;
    a->g_out = g_out; a->g_meta = (GMeta *)g_meta;
}

size_t shard_gmeta_bytes(int B) { return sizeof(GMeta) * (size_t)B; }

int launch_shard_rerank(evdb_store *s, const double *d_q64, int B, int KP, int k, int kk, int metric, int rank,
                        int world, uint64_t n_total, const ExchangeView &win, const PushTarget &e_push, uint64_t *g_out,
                        void *g_meta, cudaStream_t st) {
    ShardArgs a;
    fill_shard_args(&a, s, d_q64, B, KP, k, kk, metric, rank, world, n_total, g_out, g_meta);
    a.win = win;
    a.e_push = e_push;
    const size_t smem = (size_t)kShWarps * kShPerWarp;
    EVDB_TRY(ensure_func_smem((const void *)shard_rerank_kernel, smem));
    shard_rerank_kernel<<<(B + kShWarps - 1) / kShWarps, kShWarps * kMwGroup * 32, smem, st>>>(a);
    EVDB_CUDA(cudaGetLastError());
    s->n_launches++;
    return EVDB_OK;
}

int launch_shard_final(evdb_store *s, int B, int KP, int k, int kk, int metric, int rank, int world, uint64_t n_total,
                       const ExchangeView &ex, const uint64_t *g_out, const void *g_meta, uint64_t *out_blob,
                       cudaStream_t st) {
    ShardArgs a;
    fill_shard_args(&a, s, nullptr, B, KP, k, kk, metric, rank, world, n_total, const_cast<uint64_t *>(g_out),
                    const_cast<void *>(g_meta));
    a.ex = ex;
    a.out_blob = out_blob;
    shard_final_kernel<<<(B + kShWarps - 1) / kShWarps, kShWarps * 32, 0, st>>>(a);
    EVDB_CUDA(cudaGetLastError());
    s->n_launches++;
    return EVDB_OK;
}

static size_t select_smem(int threads) {
    return sizeof(uint64_t) * kSelSort + sizeof(uint64_t) * 2 * kMaxKP +
           (threads == 1024 ? sizeof(double) * (threads / 32) * 2 * kExactChunk : 0);
}

int launch_select(evdb_store *s, const double *d_q64, const uint64_t *partial, const RawCands *raw, int L, int KP, int B,
                  int kk, int kstride, int metric, float eps_abs, float eps_rel, const float *eps_q,
                  int squared, uint64_t slot_base, uint64_t *d_out_ids, double *d_out_dists, int32_t *d_out_counts,
                  int32_t *d_out_flags, cudaStream_t st) {
    SelectArgs a;
    a.rows = s->rows; a.row_bytes = s->row_bytes; a.norm64 = s->norm64; a.qms64 = s->qms64;
    a.n = s->count; a.d = s->dim; a.q64 = d_q64; a.partial = partial; a.L = L; a.KP = KP;
    memset(&a.raw, 0, sizeof(a.raw));
    if (raw) {
        a.raw = *raw;
        if (L > kRawMaxLists) return EVDB_E_UNSUPPORTED;
    }
    a.kk = kk; a.kstride = kstride; a.metric = metric; a.eps_abs = eps_abs; a.eps_rel = eps_rel;
    a.eps_q = eps_q; a.squared = squared;
    a.win_mode = 0;
    { static int v = -1; if (v < 0) { const char *e = getenv("EVDB_SEL_VARIANT"); v = e ? atoi(e) : 0; } a.variant = v; }
    a.slot_base = slot_base; a.slot_mul = s->slot_mul; a.out_ids = d_out_ids; a.out_dists = d_out_dists;
    a.out_counts = d_out_counts; a.out_flags = d_out_flags;
    if (raw && B >= 8 && KP <= kSwMaxKP && s->dtype == EVDB_F32 && !(a.variant & 32)) {  // GEMM plan, query batch
        s->n_launches++;
        return launch_select_warp(a, B, st);
    }
    const int threads = B >= 8 ? 256 : 1024;
    size_t smem = select_smem(threads);
    void (*fn)(const SelectArgs) = nullptr;
#define EVDB_SEL(DT) fn = threads == 256 ? select_kernel<DT, 256> : select_kernel<DT, 1024>
    switch (s->dtype) {
        case EVDB_F32: EVDB_SEL(EVDB_F32); break;
        case EVDB_BF16: EVDB_SEL(EVDB_BF16); break;
        case EVDB_U8: EVDB_SEL(EVDB_U8); break;
        default: EVDB_SEL(EVDB_U4); break;
    }
#undef EVDB_SEL
    EVDB_TRY(ensure_func_smem((const void *)fn, smem));
    EVDB_CUDA(launch_chained(fn, dim3(B), dim3(threads), smem, st, 1, a));
    s->n_launches++;
    if (a.variant & 16) {
        cudaStreamSynchronize(st);
        unsigned long long h[10], z[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyFromSymbol(h, g_sel_dbg, sizeof(h));
        cudaMemcpyToSymbol(g_sel_dbg, z, sizeof(z));
        fprintf(stderr, "[select dbg] cycles/CTA: counts+scan=%.0f gather=%.0f select=%.0f sort=%.0f ncand=%.0f fold=%.0f final=%.0f nrer=%.1f keys=%.0f\n",
                (double)h[0] / B, (double)h[1] / B, (double)h[2] / B, (double)h[3] / B, (double)h[4] / B, (double)h[5] / B,
                (double)h[6] / B, (double)h[7] / B, (double)h[8] / B);
    }
    return EVDB_OK;
}

// One launch for query prep + scan + selection (small float stores, B <= 8).  G = CTAs per query, partial /
// eps_q / arrive = the store's workspaces.  EVDB_E_UNSUPPORTED when the shape does not fit (caller falls back).
int launch_small_fused(evdb_store *s, const double *d_q64, int B, int KP, int kk, int kstride, int metric, int G, int tpr,
                       uint64_t *partial, float *eps_q, unsigned int *arrive, float eps_abs, float eps_rel,
                       uint64_t slot_base, uint64_t *d_out_ids, double *d_out_dists, int32_t *d_out_counts,
                       int32_t *d_out_flags, cudaStream_t st) {
    if ((s->dtype != EVDB_F32 && s->dtype != EVDB_BF16) || B > 8 || KP > kAppendMaxKP) return EVDB_E_UNSUPPORTED;
    FusedArgs f;
    memset(&f, 0, sizeof(f));
    SelectArgs &a = f.sel;
    a.rows = s->rows; a.row_bytes = s->row_bytes; a.norm64 = s->norm64; a.qms64 = s->qms64;
    a.n = s->count; a.d = s->dim; a.q64 = d_q64; a.partial = partial; a.L = G; a.KP = KP;
    a.kk = kk; a.kstride = kstride; a.metric = metric; a.eps_abs = eps_abs; a.eps_rel = eps_rel;
    a.eps_q = eps_q; a.squared = 0; a.win_mode = 0; a.variant = 0;
    a.slot_base = slot_base; a.slot_mul = s->slot_mul; a.out_ids = d_out_ids; a.out_dists = d_out_dists;
    a.out_counts = d_out_counts; a.out_flags = d_out_flags;
    f.inv_norm = s->inv_norm; f.partial = partial; f.eps_q = eps_q; f.arrive = arrive;
    f.nch = s->nch; f.tpr = tpr; f.G = G;
    const size_t scan_smem = (size_t)s->nch * (s->dtype == EVDB_F32 ? 16 : 32) + scan_list_bytes(KP);
    size_t smem = select_smem(256);
    if (scan_smem > smem) smem = scan_smem;
    if (smem > 100 * 1024) return EVDB_E_UNSUPPORTED;
    void (*fn)(const FusedArgs) = nullptr;
#define EVDB_FUSED(DT)                                                                             \
    fn = metric == EVDB_COSINE ? small_fused_kernel<EVDB_COSINE, DT>                                 \
       : metric == EVDB_EUCLIDEAN ? small_fused_kernel<EVDB_EUCLIDEAN, DT> : small_fused_kernel<EVDB_MANHATTAN, DT>
    if (s->dtype == EVDB_F32) { EVDB_FUSED(EVDB_F32); } else { EVDB_FUSED(EVDB_BF16); }
#undef EVDB_FUSED
    EVDB_TRY(ensure_func_smem((const void *)fn, smem));
    fn<<<dim3(G, B), 256, smem, st>>>(f);
    s->n_launches++;
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// exhaustive fp64 plan
// ----------------------------------------------------------------------------
template <int DTYPE>
__global__ void __launch_bounds__(256) exact_all_kernel(const uint8_t *__restrict__ rows,
                                                        size_t row_bytes,
                                                        const double *__restrict__ norm64,
                                                        const double2 *__restrict__ qms64,
                                                        uint64_t n, int d, const double *__restrict__ q,
                                                        int metric, uint64_t *__restrict__ keys,
                                                        uint32_t *__restrict__ slots) {
    __shared__ double sp_all[8 * 2 * kExactChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sp = sp_all + warp * 2 * kExactChunk;
    for (uint64_t r = (uint64_t)blockIdx.x * 8 + warp; r < n; r += (uint64_t)gridDim.x * 8) {
        double mn = 0.0, sc = 0.0;
        if (DTYPE == EVDB_U8 || DTYPE == EVDB_U4) {
            double2 ms = qms64[r];
            mn = ms.x;
            sc = ms.y;
        }
        double dist = exact_distance_warp<DTYPE>(rows + r * row_bytes, mn, sc, q, d, metric,
                                                 norm64[r], sp, lane);
        if (lane == 0) {
            keys[r] = f64_orderable(dist);
            slots[r] = (uint32_t)r;
        }
    }
}

__global__ void emit_sorted_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ slots,
                                   int kout, int kstride, uint64_t slot_base, uint64_t slot_mul, uint64_t *out_ids,
                                   double *out_dists, int32_t *out_count, int32_t *out_flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kstride; i += gridDim.x * blockDim.x) {
        if (i < kout) {
            uint64_t ob = keys[i];
            uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            out_dists[i] = __longlong_as_double((long long)bits);
            out_ids[i] = slot_base + slots[i] * slot_mul;
        } else {
            out_dists[i] = 0.0;
            out_ids[i] = kKeyMax;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *out_count = kout;
        if (out_flag) *out_flag = 0;
    }
}

int exact_plan_search(evdb_store *s, const double *d_q64, int B, int kk, int kstride, int metric,
                      uint64_t slot_base, uint64_t *d_out_ids, double *d_out_dists,
                      int32_t *d_out_counts, int32_t *d_out_flags, cudaStream_t st) {
    const uint64_t n = s->count;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n, 0, 64, st);
    size_t kb = round_up64(n * sizeof(uint64_t), 256), sb = round_up64(n * sizeof(uint32_t), 256);
    size_t need = 2 * kb + 2 * sb + cub_bytes;
    EVDB_TRY(ensure_bytes(&s->w_tmp, &s->w_tmp_cap, need));
    uint8_t *p = (uint8_t *)s->w_tmp;
    uint64_t *k0 = (uint64_t *)p, *k1 = (uint64_t *)(p + kb);
    uint32_t *s0 = (uint32_t *)(p + 2 * kb), *s1 = (uint32_t *)(p + 2 * kb + sb);
    void *cub_tmp = p + 2 * kb + 2 * sb;
    int grid = s->sm_count * 8;
    int kout = (uint64_t)kk < n ? kk : (int)n;
    for (int b = 0; b < B; ++b) {
        const double *q = d_q64 + (size_t)b * s->dim;
        switch (s->dtype) {
            case EVDB_F32: exact_all_kernel<EVDB_F32><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->norm64, s->qms64, n, s->dim, q, metric, k0, s0); break;
            case EVDB_BF16: exact_all_kernel<EVDB_BF16><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->norm64, s->qms64, n, s->dim, q, metric, k0, s0); break;
            case EVDB_U8: exact_all_kernel<EVDB_U8><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->norm64, s->qms64, n, s->dim, q, metric, k0, s0); break;
            default: exact_all_kernel<EVDB_U4><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->norm64, s->qms64, n, s->dim, q, metric, k0, s0); break;
        }
        EVDB_CUDA(cudaGetLastError());
        // LSD radix sort is stable: equal distances keep ascending slot order
        EVDB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k0, k1, s0, s1, (int64_t)n, 0, 64, st));
        emit_sorted_kernel<<<(kstride + 255) / 256 > 0 ? (kstride + 255) / 256 : 1, 256, 0, st>>>(
            k1, s1, kout, kstride, slot_base, s->slot_mul, d_out_ids + (size_t)b * kstride,
            d_out_dists + (size_t)b * kstride, d_out_counts + b, d_out_flags ? d_out_flags + b : nullptr);
        EVDB_CUDA(cudaGetLastError());
        s->n_launches += 3;
    }
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// G-way merge of per-shard (distance, id) lists after the allgather
// ----------------------------------------------------------------------------
// Rank g's list of query b: ids/dists at [g*stride8 + b*k ..], count at counts[g*stride4 + b]
// (flags likewise, optional).  The same kernel serves separate [G][B][k] arrays and the packed
// per-rank blobs of the one-collective exchange.
__global__ void __launch_bounds__(1024) merge_topk_kernel(const uint64_t *__restrict__ ids,
                                                          const double *__restrict__ dists,
                                                          const int32_t *__restrict__ counts,
                                                          const int32_t *__restrict__ flags,
                                                          size_t stride8, size_t stride4, int G,
                                                          int B, int k, int nsort,
                                                          uint64_t *__restrict__ out_ids,
                                                          double *__restrict__ out_dists,
                                                          int32_t *__restrict__ out_counts,
                                                          int32_t *__restrict__ out_flags,
                                                          const unsigned long long *arrived,
                                                          unsigned long long epoch) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *dk = reinterpret_cast<uint64_t *>(smem);
    uint64_t *di = dk + nsort;
    const int b = blockIdx.x;
    if (arrived) {
        // peer-memory exchange (exchange.cu): rank g's blob is complete once arrived[g] reaches
        // this search's epoch.  Peers run on OTHER GPUs, so this wait depends on no kernel of this
        // device; it is bounded -- a dead peer must trap, never hang the GPU.
        if (threadIdx.x < G) {
            const volatile unsigned long long *f = arrived + threadIdx.x;
            const long long t0 = clock64();
            while (*f < epoch)
                if (clock64() - t0 > 8000000000ll) __trap();
        }
        __threadfence_system();
        __syncthreads();
    }
    int total = 0, flag = 0;
    for (int g = 0; g < G; ++g) {
        total += __ldcg(counts + (size_t)g * stride4 + b);
        if (flags) flag |= __ldcg(flags + (size_t)g * stride4 + b);
    }
    for (int i = threadIdx.x; i < nsort; i += blockDim.x) {
        uint64_t key = kKeyMax, id = kKeyMax;
        if (i < G * k) {
            int g = i / k, j = i % k;
            if (j < __ldcg(counts + (size_t)g * stride4 + b)) {  // L2 loads: peers write these buffers
                size_t o = (size_t)g * stride8 + (size_t)b * k + j;
                key = f64_orderable(__ldcg(dists + o));
                id = __ldcg(ids + o);
            }
        }
        dk[i] = key;
        di[i] = id;
    }
    __syncthreads();
    block_bitonic_sort_pairs(dk, di, nsort);
    int kout = total < k ? total : k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        size_t o = (size_t)b * k + i;
        if (i < kout) {
            uint64_t ob = dk[i];
            uint64_t bits = ob ^ ((ob >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
            out_dists[o] = __longlong_as_double((long long)bits);
            out_ids[o] = di[i];
        } else {
            out_dists[o] = 0.0;
            out_ids[o] = kKeyMax;
        }
    }
    if (threadIdx.x == 0) {
        out_counts[b] = kout;
        if (out_flags) out_flags[b] = flag;
    }
}

// Lists too long for one CTA's shared memory (G * k keys > 8192: K in the thousands): every list is
// already ascending by (distance, id) and ids are unique, so the merged position of an element is
// its own index plus, for every other list, the number of elements there that precede it -- one
// binary search each.  No scratch, no sort; one thread per element.
__global__ void __launch_bounds__(256) merge_rank_kernel(const uint64_t *__restrict__ ids, const double *__restrict__ dists,
                                                         const int32_t *__restrict__ counts, const int32_t *__restrict__ flags,
                                                         size_t stride8, size_t stride4, int G, int B, int k,
                                                         uint64_t *__restrict__ out_ids, double *__restrict__ out_dists,
                                                         int32_t *__restrict__ out_counts, int32_t *__restrict__ out_flags,
                                                         const unsigned long long *arrived, unsigned long long epoch) {
    const int b = blockIdx.y;
    if (arrived) {
        if ((int)threadIdx.x < G) {
            const volatile unsigned long long *f = arrived + threadIdx.x;
            const long long t0 = clock64();
            while (*f < epoch)
                if (clock64() - t0 > 8000000000ll) __trap();
        }
        __threadfence_system();
        __syncthreads();
    }
    int total = 0, flag = 0;
    for (int g = 0; g < G; ++g) {
        total += __ldcg(counts + (size_t)g * stride4 + b);
        if (flags) flag |= __ldcg(flags + (size_t)g * stride4 + b);
    }
    const int kout = total < k ? total : k;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < G * k) {
        const int g = e / k, j = e % k;
        if (j < __ldcg(counts + (size_t)g * stride4 + b)) {
            const size_t o = (size_t)g * stride8 + (size_t)b * k + j;
            const double dv = __ldcg(dists + o);
            const uint64_t key = f64_orderable(dv), id = __ldcg(ids + o);
            int rank = j;
            for (int g2 = 0; g2 < G && rank < kout; ++g2) {
                if (g2 == g) continue;
                const size_t o2 = (size_t)g2 * stride8 + (size_t)b * k;
                int lo = 0, hi = __ldcg(counts + (size_t)g2 * stride4 + b);   // first element of list g2 that does NOT precede (key, id)
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const uint64_t k2 = f64_orderable(__ldcg(dists + o2 + mid)), i2 = __ldcg(ids + o2 + mid);
                    if (k2 < key || (k2 == key && i2 < id)) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            if (rank < kout) {
                out_dists[(size_t)b * k + rank] = dv;
                out_ids[(size_t)b * k + rank] = id;
            }
        }
    }
    if (e >= kout && e < k) {
        out_dists[(size_t)b * k + e] = 0.0;
        out_ids[(size_t)b * k + e] = kKeyMax;
    }
    if (e == 0) {
        out_counts[b] = kout;
        if (out_flags) out_flags[b] = flag;
    }
}

static int merge_launch(const uint64_t *ids, const double *dists, const int32_t *counts, const int32_t *flags,
                        size_t stride8, size_t stride4, int G, int B, int k, uint64_t *out_ids,
                        double *out_dists, int32_t *out_counts, int32_t *out_flags, cudaStream_t st,
                        const unsigned long long *arrived = nullptr, unsigned long long epoch = 0) {
    if (G <= 0 || B <= 0 || k <= 0) return EVDB_E_BAD_ARG;
    int nsort = next_pow2(G * k);
    if (nsort < 2) nsort = 2;
    size_t smem = (size_t)nsort * 16;
    if (smem > 128 * 1024) {
        if (B > 65535) return EVDB_E_UNSUPPORTED;
        const dim3 grid((unsigned)(((size_t)G * k + 255) / 256), (unsigned)B);
        merge_rank_kernel<<<grid, 256, 0, st>>>(ids, dists, counts, flags, stride8, stride4, G, B, k, out_ids, out_dists,
                                                out_counts, out_flags, arrived, epoch);
        EVDB_CUDA(cudaGetLastError());
        return EVDB_OK;
    }
    EVDB_TRY(ensure_func_smem((const void *)merge_topk_kernel, smem));
    const int threads = nsort >= 2048 ? 1024 : (nsort >= 512 ? 256 : 128);
    merge_topk_kernel<<<B, threads, smem, st>>>(ids, dists, counts, flags, stride8, stride4, G, B, k, nsort,
                                                out_ids, out_dists, out_counts, out_flags, arrived, epoch);
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

int launch_merge_topk(const uint64_t *ids, const double *dists, const int32_t *counts, int G, int B,
                      int k, uint64_t *out_ids, double *out_dists, int32_t *out_counts,
                      cudaStream_t st) {
    return merge_launch(ids, dists, counts, nullptr, (size_t)B * k, (size_t)B, G, B, k, out_ids, out_dists,
                        out_counts, nullptr, st);
}

// Packed blobs (one per rank, `blob_words` u64 words apart): [B*k ids][B*k dists][B counts i32][B flags i32]
// blob_stride: u64 words between consecutive ranks' blobs (>= 2*B*k + B)
int launch_merge_topk_packed(const uint64_t *blobs, size_t blob_stride, int G, int B, int k, uint64_t *out_blob,
                             cudaStream_t st, const unsigned long long *arrived, unsigned long long epoch) {
    const size_t nk = (size_t)B * k;
    const int32_t *cf = reinterpret_cast<const int32_t *>(blobs + 2 * nk);
    int32_t *ocf = reinterpret_cast<int32_t *>(out_blob + 2 * nk);
    return merge_launch(blobs, reinterpret_cast<const double *>(blobs + nk), cf, cf + B, blob_stride, 2 * blob_stride, G, B,
                        k, out_blob, reinterpret_cast<double *>(out_blob + nk), ocf, ocf + B, st, arrived, epoch);
}

}  // namespace evdb
