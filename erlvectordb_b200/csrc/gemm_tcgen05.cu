// gemm_tcgen05.cu -- batched cosine / euclidean search as a TMA-fed tcgen05/TMEM GEMM with a
// fused per-query top-k epilogue (family B).
//
// Replaces, for query batches, the same reference text as the scans: the maps:fold of
// cosine_distance/2 over all rows (reference src/vector_store.erl:227-252) and the
// euclidean form of src/vector_utils.erl:38-40 -- here for B queries at once as one GEMM
//     acc[q][r] = sum_k Qh[q][k] * Vh[r][k]          (fp16 operands, fp32 accumulate in TMEM)
//   cosine   : Qh = q/||q||, Vh = v/||v||                      key score = 1 - acc
//   euclidean: Qh = [s*q, 1, 1, 1], Vh = [s*v, h1, h2, h3]     key score = ||q||^2 - (2/s^2)*acc
//              with h1+h2+h3 = -s^2*||v||^2/2 split over three fp16 columns, s a power of two
//              chosen from the store's largest row norm, so acc = s^2*(q.v - ||v||^2/2): the row
//              norms ride in the K dimension and the epilogue is the same for both metrics.
// The GEMM only GENERATES CANDIDATES: key scores carry a rigorous per-query error bound eps
// (fp16 operand rounding + fp32 accumulation), the KP best per query go to select.cu, which
// re-ranks them in exact fp64 and proves the window complete -- returned distances never see
// fp16.
//
// Kernel anatomy (one persistent CTA per SM, 640 threads; d in [512, 1024]: CTA pairs as 2-CTA clusters):
//   warp 0   TMA producer: cp.async.bulk.tensor 2D tiles (128B swizzle) of Q [128 x 64] and
//            V [256 x 64] into a 4-stage shared-memory ring (6 stages of Q + half of V in the pair
//            variant), mbarrier complete_tx
//   warp 1   MMA issuer: one elected thread, tcgen05.mma kind::f16, M=128 (cta_group::1) or M=256
//            (cta_group::2, issued by the pair's leader) x N=256 x K=16, up to 4 per stage;
//            tcgen05.commit frees the stage / publishes the accumulator
//   warp 2   TMEM allocator (512 columns = 2 accumulator stages x 256 fp32 columns)
//   warps 4-19 epilogue (16 warps): thread t owns TMEM lane t == query t of the CTA's 128-query block,
//            four warps per lane quarter take 64 of the 256 columns each; tcgen05.ld 32 columns at a time.
//            The score stream is filtered in the ACCUMULATOR domain: a 3-input-max tree over
//            the 32 values against the query's admission threshold; only a chunk that beats it is
//            expanded, and its hits are appended to the query's private candidate buffer (global
//            memory, L2 resident).  A buffer that nears capacity is reduced to its KP best by value
//            bisection with warp population counts (no sort), which also tightens the threshold.
// A CTA keeps one query block for a whole sweep over its share of the corpus tiles; the 256-row
// corpus tile is shared by the MB CTAs working on different query blocks at the same time (L2
// hits).  Thresholds start from a sampled pre-pass (the same kernel in pooling mode over a strided
// row sample).  The raw candidate buffers go to select.cu unsorted (select_warp_kernel gathers them).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_fp16.h>

#include "internal.h"
#include "topk.cuh"
#include "tc05.cuh"

namespace evdb {

constexpr int GN = 256;        // corpus rows per tile (UMMA N)
constexpr int GK = 64;         // K elements per stage (64 fp16 = one 128-byte swizzle row)
constexpr int GUK = 16;        // UMMA K
constexpr uint32_t kStageABytes = GM * GK * 2;   // 16 KB: the CTA's 128 queries x 64 K
constexpr uint32_t kTailABytes = GM * GUK * 2;   // 4 KB: the norm K-step of the euclidean plan (Q side)
// PAIR = two CTAs of a cluster (an SM pair) run ONE tcgen05.mma.cta_group::2 of M = 256: each CTA
// brings its own 128-query block and HALF of the 256-row corpus tile, so a stage is 32 KB instead
// of 48 KB (a third less L2 -> shared-memory traffic per flop) and six stages fit instead of four.
template <bool PAIR>
struct GemmCfg {
    static constexpr int kStages = PAIR ? 6 : 4;
    static constexpr uint32_t kBRows = PAIR ? GN / 2 : GN;           // corpus rows this CTA loads per tile
    static constexpr uint32_t kStageBBytes = kBRows * GK * 2;
    static constexpr uint32_t kStageBytes = kStageABytes + kStageBBytes;
    static constexpr uint32_t kTailBBytes = kBRows * GUK * 2;
    static constexpr uint32_t kTailBytes = kTailABytes + kTailBBytes;
    static constexpr int kBars = 2 * kStages + 8;                    // full[S] empty[S] tfull[2] tempty[2] nfull[2] nempty[2]
    static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + 2 * kTailBytes + kBars * 8 + 16;
};


struct GemmArgs {
    uint64_t n;          // corpus rows
    int B;               // live queries (rows >= B of the padded batch are never admitted)
    int kblocks;         // ceil(K / 64), K = operand columns (dim, or dim + 3 for euclidean)
    int last_ksteps;     // UMMA K-steps in the last k-block (1..4)
    int tail;            // euclidean: one more K-step from the [rows][16] norm operands (32B swizzle)
    int cap;             // candidate buffer capacity in use (<= kCandCapMax)
    int nt;              // corpus tiles = ceil(n / 256)
    int MB;              // query blocks processed concurrently
    int NG;              // CTAs per query block (2*NG lists per query)
    int nchunks;         // sweeps: query blocks [c*MB, (c+1)*MB)
    int KP;
    uint64_t *cand;      // [sweep][CTA][part][cap][128] append buffers (entry-major: thread t's i-th key at [i][t])
    int *cand_cnt;       // [sweep][CTA][part][128] fill counts at the end of the sweep
    int mode;            // 0 = fused top-k, 1 = pooled key scores of a row sample (threshold seeding)
    float *dump;         // mode 1: [Bpad][dump_ld] best key score of every 32-row chunk
    int dump_ld;
    const uint32_t *thr0;  // mode 0: per-query starting threshold, orderable key score (NULL = none)
    const float *qc0;    // [Bpad] key score = fma(acc, c1, qc0[q])
    float c1;            // < 0
    int debug;           // EVDB_GEMM_DEBUG (measurement only): 1 = no appends, 2 = no TMEM loads, 8 = cycle breakdown, 16 = no Q reloads (pair variant)
    unsigned long long *dbg;  // [CTAs][epilogue warps][8]
};

// ---- the kernel ---------------------------------------------------------------------------
template <bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmVt,
                 const GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem is only guaranteed 16-byte aligned: realign to 1024 for the 128B swizzle
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using Cfg = GemmCfg<PAIR>;
    constexpr int kGemmStages = Cfg::kStages;
    constexpr uint32_t kStageBytes = Cfg::kStageBytes, kTailBytes = Cfg::kTailBytes;
    uint8_t *stage_base = smem;                                            // [stages][A|B]
    uint8_t *tail_base = smem + kGemmStages * kStageBytes;                 // [2][A tail | B tail]
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail_base + 2 * kTailBytes);  // full[S] empty[S] tfull[2] tempty[2] nfull[2] nempty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + Cfg::kBars);
    // PAIR: rank 0 of the cluster issues the MMAs; its full/tempty/nfull barriers collect both CTAs
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KP = a.KP;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kGemmStages);
    const uint32_t tfull0 = smem_u32(bars + 2 * kGemmStages), tempty0 = smem_u32(bars + 2 * kGemmStages + 2);
    const uint32_t nfull0 = smem_u32(bars + 2 * kGemmStages + 4), nempty0 = smem_u32(bars + 2 * kGemmStages + 6);

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kGemmStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, PAIR ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
            mbar_init(nfull0 + 8 * i, 1);
            mbar_init(nempty0 + 8 * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // barriers and TMEM are set up: everything below reads what earlier kernels of the chain wrote
    pdl_wait();
    pdl_launch_dependents();

    const int cta = blockIdx.x;
    const int nCTA = a.MB * a.NG;
    const bool active = cta < nCTA;
    const int mb_local = cta % a.MB, ng = cta / a.MB;
    int my_tiles = 0;
    if (active) my_tiles = (a.nt - ng + a.NG - 1) / a.NG;  // tiles ng, ng+NG, ...

    if (active && warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0, ts = 0;
            uint32_t phase = 0, tphase = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                const int qrow = (c * a.MB + mb_local) * GM;
                for (int t = 0; t < my_tiles; ++t) {
                    const int vrow = (ng + t * a.NG) * GN;
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
                        if (PAIR) {
                            // both CTAs' tiles complete on the LEADER's barrier; it alone posts the byte count
                            const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
                            const bool skipq = (a.debug & 16) && t > 0;   // measurement only: stale Q operand, V traffic alone
                            if (leader) mbar_arrive_expect_tx(full0 + 8 * stage, skipq ? 2 * Cfg::kStageBBytes : 2 * kStageBytes);
                            if (!skipq) tma_load_2d_pair(sa, &tmQ, fb, kb * GK, qrow);
                            tma_load_2d_pair(sa + kStageABytes, &tmV, fb, kb * GK, vrow + (int)(crank * Cfg::kBRows));
                        } else {
                            mbar_arrive_expect_tx(full0 + 8 * stage, kStageBytes);
                            tma_load_2d(sa, &tmQ, full0 + 8 * stage, kb * GK, qrow);
                            tma_load_2d(sa + kStageABytes, &tmV, full0 + 8 * stage, kb * GK, vrow);
                        }
                        if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
                    }
                    if (a.tail) {  // the norm K-step: [128 x 16] ones-pattern and [256 x 16] split row norms
                        mbar_wait(nempty0 + 8 * ts, tphase ^ 1);
                        const uint32_t sa = smem_u32(tail_base + ts * kTailBytes);
                        if (PAIR) {
                            const uint32_t fb = mapa_u32(nfull0 + 8 * ts, 0);
                            if (leader) mbar_arrive_expect_tx(nfull0 + 8 * ts, 2 * kTailBytes);
                            tma_load_2d_pair(sa, &tmQt, fb, 0, qrow);
                            tma_load_2d_pair(sa + kTailABytes, &tmVt, fb, 0, vrow + (int)(crank * Cfg::kBRows));
                        } else {
                            mbar_arrive_expect_tx(nfull0 + 8 * ts, kTailBytes);
                            tma_load_2d(sa, &tmQt, nfull0 + 8 * ts, 0, qrow);
                            tma_load_2d(sa + kTailABytes, &tmVt, nfull0 + 8 * ts, 0, vrow);
                        }
                        if (++ts == 2) { ts = 0; tphase ^= 1; }
                    }
                }
            }
        }
    } else if (active && warp == 1) {
        // ===== MMA issuer (PAIR: the leader CTA only, for both) =====
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * GM : GM, GN);
            int stage = 0, ts = 0;
            uint32_t phase = 0, tphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                for (int t = 0; t < my_tiles; ++t) {
                    mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);  // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)acc * GN;
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
                        const uint64_t adesc = make_sw128_kmajor_desc(sa);
                        const uint64_t bdesc = make_sw128_kmajor_desc(sa + kStageABytes);
                        const int ksteps = kb + 1 == a.kblocks ? a.last_ksteps : GK / GUK;
#pragma unroll
                        for (int k = 0; k < GK / GUK; ++k) {
                            // advance 16 fp16 = 32 bytes along K inside the swizzled row: +2 in >>4 units
                            if (k < ksteps) {
                                if (PAIR) umma_f16_pair(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                                        (uint32_t)((kb | k) != 0));
                                else umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                              (uint32_t)((kb | k) != 0));
                            }
                        }
                        // stage reusable (in both CTAs) once these MMAs retire
                        if (PAIR) umma_commit_pair(empty0 + 8 * stage); else umma_commit(empty0 + 8 * stage);
                        if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
                    }
                    if (a.tail) {
                        mbar_wait(nfull0 + 8 * ts, tphase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tail_base + ts * kTailBytes);
                        if (PAIR) {
                            umma_f16_pair(tmem_d, make_sw32_kmajor_desc(sa), make_sw32_kmajor_desc(sa + kTailABytes), idesc, 1u);
                            umma_commit_pair(nempty0 + 8 * ts);
                        } else {
                            umma_f16(tmem_d, make_sw32_kmajor_desc(sa), make_sw32_kmajor_desc(sa + kTailABytes), idesc, 1u);
                            umma_commit(nempty0 + 8 * ts);
                        }
                        if (++ts == 2) { ts = 0; tphase ^= 1; }
                    }
                    // accumulator complete (each CTA's epilogue drains its own 128 TMEM lanes)
                    if (PAIR) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (active && warp >= 4) {
        // ===== epilogue: 16 warps; thread <-> TMEM lane <-> query, 4 warps per lane quarter split the columns =====
        const int ew = warp - 4;                 // 0..15
        const int lg = ew & 3;                   // == warp % 4: TMEM lanes [32*lg, 32*lg+32)
        const int part = ew >> 2;                // columns [64*part, 64*part+64) of every tile
        const int et = lg * 32 + lane;           // query within the CTA's block
        const float kInf = __int_as_float(0x7f800000);
        const float c1 = a.c1;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long t_wait = 0, t_chunk = 0, t_prune = 0, n_prune = 0;
        for (int c = 0; c < a.nchunks; ++c) {
            const size_t lbase = ((size_t)c * nCTA + cta) * kEpiParts + part;
            // entry i of thread e at [i*GM + e]: lanes appending at similar fill levels share sectors
            uint64_t *cbase = a.cand + lbase * a.cap * GM;
            uint64_t *mybuf = cbase + et;
            int cnt = 0;
            const size_t qglob = (size_t)(c * a.MB + mb_local) * GM + et;
            const float c0 = a.qc0[qglob];
            const bool admit = a.mode == 0 && qglob < (size_t)a.B && !(a.debug & 1);
            // admission threshold: seeded by the sampled pre-pass (a valid upper bound on the
            // query's KP-th best key score), tightened by every prune; held in the accumulator domain
            float tau = a.thr0 ? f32_from_orderable(a.thr0[qglob]) : kInf;
            float thrS = acc_threshold(tau, c0, c1);
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = ng + t * a.NG;
                const uint32_t row0 = (uint32_t)tile * GN;
                long long tw0 = 0, tw1 = 0;
                if (a.debug & 8) tw0 = clock64();
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                tc_fence_after();
                if (a.debug & 8) { tw1 = clock64(); t_wait += tw1 - tw0; }
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)acc * GN + part * (GN / kEpiParts);
                constexpr int kChunks = GN / kEpiParts / 32;  // 32-column chunks per warp per tile
                const uint32_t nrows = (uint32_t)a.n;
                // No register double-buffering: the other three warps of this scheduler cover the
                // load latency (one tcgen05.ld in flight per warp, four per scheduler).
#pragma unroll 1
                for (int cb = 0; cb < kChunks && !(a.debug & 2); ++cb) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + cb * 32, v);
                    tmem_ld_wait();
                    const uint32_t col0 = part * (GN / kEpiParts) + cb * 32;
                    const float m = epi_chunk(v, thrS, row0 + col0, nrows, c0, c1, mybuf, cnt, admit);
                    if (a.mode == 1) {  // sampled pre-pass: best key score of the chunk
                        a.dump[qglob * a.dump_ld + (size_t)tile * 8 + part * kChunks + cb] = fmaf(m, c1, c0);
                        continue;
                    }
                    // a chunk appends at most 32 keys: prune any buffer that could overflow next
                    const unsigned need = __ballot_sync(0xffffffffu, cnt > a.cap - 32);
                    if (need) {
                        long long tp0 = 0;
                        if (a.debug & 8) tp0 = clock64();
                        n_prune += prune_buffers(cbase + lg * 32, need, KP, lane, c0, c1, cnt, tau, thrS);
                        if (a.debug & 8) t_prune += clock64() - tp0;
                    }
                }
                if (a.debug & 8) t_chunk += clock64() - tw1;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(mapa_u32(tempty0 + 8 * acc, 0));  // the leader's issuer waits for both CTAs
                    else mbar_arrive(tempty0 + 8 * acc);
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (a.mode == 0) a.cand_cnt[lbase * GM + et] = cnt;
        }
        if ((a.debug & 8) && lane == 0) {
            unsigned long long *d = a.dbg + ((size_t)cta * kEpiWarps + ew) * 8;
            d[0] = t_wait; d[1] = t_chunk; d[2] = t_prune; d[3] = 0; d[4] = n_prune;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair's MMAs can touch it
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- query preparation -------------------------------------------------------------------------
// cosine   : Qh = q / ||q|| (fp16, zero padded to [Bpad][kpitch]); key = 1 - acc
// euclidean: Qh = [sigma*q, 1, 1, 1]; key = ||q||^2 - (2/sigma^2) * acc  (approximate dist^2)
// eps_q[b] = rigorous bound on |key - ideal| for every row of the store (see the derivation in
// DESIGN.md section 4); +inf marks a query the fp16 operands cannot carry (select.cu then flags it
// and the host path re-issues it through the scan plan).
__global__ void __launch_bounds__(256) prep_queries_gemm_kernel(const double *__restrict__ q64, int B, int d,
                                                                int metric, float sigma, double vmax,
                                                                __half *__restrict__ qh, int kpitch,
                                                                __half *__restrict__ qtail,
                                                                float *__restrict__ qc0,
                                                                float *__restrict__ eps_q) {
    __shared__ double red[8];
    __shared__ double redm[8];
    pdl_launch_dependents();   // first link of the chain (launched the ordinary way): the next one may queue up
    const int b = blockIdx.x;
    __half *o = qh + (size_t)b * kpitch;
    // the norm K-step multiplies the three fp16 pieces of -sigma^2*||v||^2/2 by exactly 1
    if (qtail && threadIdx.x < GUK) qtail[(size_t)b * GUK + threadIdx.x] = __float2half_rn(threadIdx.x < 3 ? 1.0f : 0.0f);
    if (b >= B) {
        for (int i = threadIdx.x; i < kpitch; i += blockDim.x) o[i] = __float2half_rn(0.f);
        if (threadIdx.x == 0) { qc0[b] = metric == EVDB_COSINE ? 1.0f : 0.0f; eps_q[b] = 0.f; }
        return;
    }
    const double *q = q64 + (size_t)b * d;
    double ss = 0.0, mx = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        ss += q[i] * q[i];
        mx = fmax(mx, fabs(q[i]));
    }
    for (int off = 16; off > 0; off >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = ss; redm[threadIdx.x >> 5] = mx; }
    __syncthreads();
    ss = 0.0; mx = 0.0;
    for (int i = 0; i < 8; ++i) { ss += red[i]; mx = fmax(mx, redm[i]); }
    const double u = 5.9604644775390625e-08;  // 2^-24
    if (metric == EVDB_COSINE) {
        const float inv = ss > 0.0 ? (float)(1.0 / sqrt(ss)) : 0.0f;
        for (int i = threadIdx.x; i < kpitch; i += blockDim.x)
            o[i] = __float2half_rn(i < d ? (float)q[i] * inv : 0.f);
        if (threadIdx.x == 0) {
            qc0[b] = 1.0f;
            // both operands are unit vectors rounded to fp16 (2^-11 relative each, 2^-25 absolute
            // in the subnormal range), fp32 accumulation over dim terms in the tensor core
            // (bounded as dim * 2^-22, truncation included), one fp32 subtract
            eps_q[b] = (float)(0.0009765625 * 1.01 + sqrt((double)d) * u + (double)d * 4.0 * u);
        }
    } else {
        const double sg = (double)sigma;
        const bool fits = mx * sg <= 60000.0;
        for (int i = threadIdx.x; i < kpitch; i += blockDim.x) {
            float v = 0.f;
            if (i < d) v = fits ? (float)(q[i] * sg) : 0.f;
            o[i] = __float2half_rn(v);
        }
        if (threadIdx.x == 0) {
            const double K = (double)(d + 3);
            const double Q = sg * sqrt(ss), V = sg * vmax;
            const double e_dot = (0.0009765625 + 4.8e-7) * Q * V + 0.5 * u * sqrt(K) * (Q + V) * 1.001 + K * u * u * 4.0;
            const double e_norm = 0.5 * V * V * 4.66e-10 + 3.0 * 0.5 * u;
            const double e_acc = K * 4.0 * u * (Q * V + 0.5 * V * V);
            const double scale = 2.0 / (sg * sg);
            const double e_key = (ss + scale * (Q * V + 0.5 * V * V)) * 3.0 * u;
            const double eps = 1.01 * ((e_dot + e_norm + e_acc) * scale + e_key);
            qc0[b] = (float)ss;
            eps_q[b] = fits && eps < 3.0e38 ? (float)eps : __int_as_float(0x7f800000);
        }
    }
}

// ---- euclidean operands: fp16(sigma*v) [pitch] and the norm tail {h1, h2, h3, 0 x 13}, one warp per row ----
__global__ void __launch_bounds__(256) build_l2_shadow_kernel(const uint8_t *__restrict__ rows, size_t row_bytes,
                                                              const double *__restrict__ norm64, int d,
                                                              int pitch, float sigma, uint64_t slot0, uint64_t n,
                                                              __half *__restrict__ out, __half *__restrict__ tail) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t i = (uint64_t)blockIdx.x * 8 + warp; i < n; i += (uint64_t)gridDim.x * 8) {
        const uint64_t r = slot0 + i;
        const float *fr = reinterpret_cast<const float *>(rows + r * row_bytes);
        __half *sh = out + r * (size_t)pitch;
        for (int c = lane; c < pitch; c += 32) sh[c] = __float2half_rn(c < d ? fr[c] * sigma : 0.f);
        if (lane < GUK) {
            const double nr = norm64[r] * (double)sigma;
            double x = -0.5 * nr * nr;
            const __half h1 = __double2half(x);
            x -= (double)__half2float(h1);
            const __half h2 = __double2half(x);
            x -= (double)__half2float(h2);
            const __half h3 = __double2half(x);
            tail[r * GUK + lane] = lane == 0 ? h1 : lane == 1 ? h2 : lane == 2 ? h3 : __float2half_rn(0.f);
        }
    }
}

__global__ void max_norm_kernel(const double *__restrict__ norm64, uint64_t n, unsigned long long *out) {
    double m = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = fmax(m, norm64[i]);
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));  // m >= 0
}

// ---- threshold seeding: KP-th smallest pooled key score of each query over the sampled rows ------
// One CTA per query, values in registers, bisection on the order-preserving bits with block-wide
// counts.  Every pooled value is the key score of an actual (distinct) row, so the KP-th smallest is
// an upper bound on the query's global KP-th best key score.
constexpr int kSeedThreads = 256;
constexpr int kSeedMaxVpt = 32;  // pooled values per query <= 256 * 32 (sample <= 262144 rows)
// The search stops after 10 four-way rounds (the bracket is then 2^-20 of the value range) and
// returns the bracket's upper end, which still has >= need values at or below it.
template <int kSeedVpt>
__global__ void __launch_bounds__(kSeedThreads) seed_threshold_kernel(const float *__restrict__ dump, int ld,
                                                                      int S, int need, uint32_t *__restrict__ thr0) {
    __shared__ int s_cnt[6];
    __shared__ uint32_t s_lo, s_hi;
    pdl_wait();                // the pooled scores come from the pre-pass
    pdl_launch_dependents();
    // blockIdx.y = group of <= 8192 pooled values: the `need`-th smallest of each group, maximum over
    // the groups (atomicMax, thr0 zeroed before) -- with need = ceil(KP / groups) at least KP values
    // of the whole sample are at or below that maximum, so it still bounds the KP-th best
    const int q = blockIdx.x;
    const int g0 = blockIdx.y * (kSeedThreads * kSeedMaxVpt);
    const float *row = dump + (size_t)q * ld + g0;
    S = S - g0 < kSeedThreads * kSeedMaxVpt ? S - g0 : kSeedThreads * kSeedMaxVpt;
    uint32_t v[kSeedVpt];
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int i = 0; i < kSeedVpt; ++i) {
        const int idx = i * kSeedThreads + threadIdx.x;
        v[i] = idx < S ? f32_orderable(row[idx]) : 0xFFFFFFFFu;
        mn = min(mn, v[i]);
        if (idx < S) mx = max(mx, v[i]);
    }
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_lo = 0xFFFFFFFFu; s_hi = 0u; }
    __syncthreads();
    atomicMin(&s_lo, __reduce_min_sync(0xffffffffu, mn));
    atomicMax(&s_hi, __reduce_max_sync(0xffffffffu, mx));
    __syncthreads();
    uint32_t lo = s_lo, hi = s_hi;  // invariant: count(v <= hi) >= need
    int it = 0;
    while (lo < hi && it < 10) {  // 4-way search: three pivots per round, one barrier pair per round
        const uint32_t span = hi - lo;
        const uint32_t qd = span >> 2;
        const uint32_t p2 = lo + (span >> 1);
        const uint32_t p1 = qd ? lo + qd : p2;
        const uint32_t p3 = qd ? p2 + qd : p2;
        int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int i = 0; i < kSeedVpt; ++i) {
            c1 += v[i] <= p1 ? 1 : 0;
            c2 += v[i] <= p2 ? 1 : 0;
            c3 += v[i] <= p3 ? 1 : 0;
        }
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        int *cc = s_cnt + (it & 1) * 3;
        if ((threadIdx.x & 31) == 0) { atomicAdd(cc, c1); atomicAdd(cc + 1, c2); atomicAdd(cc + 2, c3); }
        __syncthreads();
        const int t1 = cc[0], t2 = cc[1], t3 = cc[2];
        if (threadIdx.x < 3) s_cnt[((it + 1) & 1) * 3 + threadIdx.x] = 0;
        if (t1 >= need) hi = p1;
        else if (t2 >= need) { lo = p1 + 1; hi = p2; }
        else if (t3 >= need) { lo = p2 + 1; hi = p3; }
        else lo = p3 + 1;
        ++it;
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(thr0 + q, hi);
}

// ---- host side ------------------------------------------------------------------------------

// rows x cols fp16, `pitch_elems` elements between consecutive rows of the MAP (a multiple of the
// storage pitch selects every step-th stored row: the strided sample needs no gather)
static int make_map(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                    uint32_t box_rows, bool tail = false) {
    encode_tiled_fn enc = get_encode();
    if (!enc) return EVDB_E_CUDA;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)(tail ? GUK : GK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, tail ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? EVDB_OK : EVDB_E_CUDA;
}

// One launch of the GEMM kernel: plain for a lone query block, as 2-CTA clusters otherwise.
static int launch_gemm_kernel(bool pair, int grid, cudaStream_t st, const CUtensorMap &tmQ, const CUtensorMap &tmV,
                              const CUtensorMap &tmQt, const CUtensorMap &tmVt, const GemmArgs &a) {
    if (!pair) {
        const size_t smem = GemmCfg<false>::kSmem;
        EVDB_TRY(ensure_func_smem((const void *)gemm_topk_kernel<false>, smem));
        EVDB_CUDA(launch_chained(gemm_topk_kernel<false>, dim3(grid), dim3(kGemmThreads), smem, st, 1, tmQ, tmV, tmQt, tmVt, a));
        return EVDB_OK;
    }
    const size_t smem = GemmCfg<true>::kSmem;
    EVDB_TRY(ensure_func_smem((const void *)gemm_topk_kernel<true>, smem));
    EVDB_CUDA(launch_chained(gemm_topk_kernel<true>, dim3(grid), dim3(kGemmThreads), smem, st, 2, tmQ, tmV, tmQt, tmVt, a));
    return EVDB_OK;
}

bool gemm_plan_supported(evdb_store *s, int metric, int B, int KP) {
    if (s->dtype != EVDB_F32 || !s->shadow) return false;
    if (metric != EVDB_COSINE && metric != EVDB_EUCLIDEAN) return false;
    if (KP > kGemmMaxKP || B < 1) return false;
    if (s->count < (uint64_t)GN) return false;
    if (metric == EVDB_COSINE && s->shadow_valid < s->count) return false;
    return get_encode() != nullptr;
}

int gemm_kp(int KP) { return KP <= 32 ? 32 : (KP <= 64 ? 64 : (KP <= 128 ? 128 : KP)); }
int gemm_max_batch() { return kMaxSweeps * 8 * GM; }

// Bring the euclidean operand column up to date: rows [l2_valid, count) are converted with the
// current scale; the scale itself (a power of two with sigma * max||v|| in [64, 128)) is redone,
// and the whole column rebuilt, when a larger row norm has arrived.
static int ensure_l2_shadow(evdb_store *s, cudaStream_t st) {
    const int pitch = s->spitch;
    bool rebuild = false;
    if (!s->shadow_l2 || s->l2_pitch != pitch || s->l2_cap < s->capacity) {
        if (s->shadow_l2) {
            EVDB_CUDA(cudaStreamSynchronize(st));
            cudaFree(s->shadow_l2); s->shadow_l2 = nullptr;
            cudaFree(s->l2_tail); s->l2_tail = nullptr;
        }
        EVDB_CUDA(cudaMalloc((void **)&s->shadow_l2, (size_t)s->capacity * pitch * sizeof(__half)));
        EVDB_CUDA(cudaMalloc((void **)&s->l2_tail, (size_t)s->capacity * GUK * sizeof(__half)));
        s->l2_pitch = pitch;
        s->l2_cap = s->capacity;
        rebuild = true;
    }
    if (!s->d_scalar) EVDB_CUDA(cudaMalloc((void **)&s->d_scalar, 64));
    if (s->max_norm_dirty) {
        // the bound only ever needs to cover the live rows; recomputed after ingest, never after delete
        EVDB_CUDA(cudaMemsetAsync(s->d_scalar, 0, 8, st));
        max_norm_kernel<<<s->sm_count * 4, 256, 0, st>>>(s->norm64, s->count, (unsigned long long *)s->d_scalar);
        EVDB_CUDA(cudaGetLastError());
        unsigned long long bits = 0;
        EVDB_CUDA(cudaMemcpyAsync(&bits, s->d_scalar, 8, cudaMemcpyDeviceToHost, st));
        EVDB_CUDA(cudaStreamSynchronize(st));
        double m;
        memcpy(&m, &bits, 8);
        s->max_norm = m;
        s->max_norm_dirty = 0;
        s->n_launches++;
    }
    float sigma = 1.0f;
    if (s->max_norm > 0.0) {
        int x;
        frexp(s->max_norm, &x);  // max_norm = m * 2^x, m in [0.5, 1)  =>  max_norm * 2^(7-x) in [64, 128)
        int e = 7 - x;
        if (e > 100) e = 100;
        if (e < -100) e = -100;
        sigma = (float)ldexp(1.0, e);
    }
    if (sigma != s->l2_sigma) rebuild = true;
    if (rebuild) { s->l2_valid = 0; s->l2_sigma = sigma; }
    if (s->l2_valid < s->count) {
        const uint64_t n = s->count - s->l2_valid;
        uint64_t blocks = (n + 7) / 8;
        const uint64_t cap = (uint64_t)s->sm_count * 16;
        build_l2_shadow_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(
            s->rows, s->row_bytes, s->norm64, s->dim, pitch, sigma, s->l2_valid, n, s->shadow_l2, s->l2_tail);
        EVDB_CUDA(cudaGetLastError());
        s->n_launches++;
        s->l2_valid = s->count;
    }
    return EVDB_OK;
}

// ingest hook: keep rows [slot0, slot0+n) of an existing euclidean operand column current.  A row
// whose norm outgrows the scale is caught at the next search (max_norm_dirty -> new sigma -> rebuild).
int launch_l2_shadow_rows(evdb_store *s, uint64_t slot0, uint64_t n, cudaStream_t st) {
    if (!s->shadow_l2 || n == 0 || s->l2_sigma == 0.f) return EVDB_OK;
    if (slot0 > s->l2_valid || slot0 + n > s->l2_cap) return EVDB_OK;  // picked up by ensure_l2_shadow
    uint64_t blocks = (n + 7) / 8;
    const uint64_t cap = (uint64_t)s->sm_count * 16;
    build_l2_shadow_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(
        s->rows, s->row_bytes, s->norm64, s->dim, s->l2_pitch, s->l2_sigma, slot0, n, s->shadow_l2, s->l2_tail);
    EVDB_CUDA(cudaGetLastError());
    s->n_launches++;
    if (slot0 + n > s->l2_valid) s->l2_valid = slot0 + n;
    return EVDB_OK;
}

// thr[q] (zeroed by the caller, combined by atomicMax) <- an upper bound on query q's KP-th best key score, from
// the [Bpad][pooled] best-of-32-rows key scores a pooling pass left in `dump` (also used by gemm_i8.cu)
int launch_seed_thresholds(const float *dump, int pooled, int Bpad, int KP, uint32_t *thr, cudaStream_t st) {
    const int vpt = (pooled + kSeedThreads - 1) / kSeedThreads;
    const int groups = (pooled + kSeedThreads * kSeedMaxVpt - 1) / (kSeedThreads * kSeedMaxVpt);
    const int need = (KP + groups - 1) / groups;
    const dim3 sgrid(Bpad, groups);
    const int pooled_i = pooled;
    if (vpt <= 4) EVDB_CUDA(launch_chained(seed_threshold_kernel<4>, sgrid, dim3(kSeedThreads), 0, st, 1, dump, pooled_i, pooled_i, need, thr));
    else if (vpt <= 8) EVDB_CUDA(launch_chained(seed_threshold_kernel<8>, sgrid, dim3(kSeedThreads), 0, st, 1, dump, pooled_i, pooled_i, need, thr));
    else if (vpt <= 16) EVDB_CUDA(launch_chained(seed_threshold_kernel<16>, sgrid, dim3(kSeedThreads), 0, st, 1, dump, pooled_i, pooled_i, need, thr));
    else EVDB_CUDA(launch_chained(seed_threshold_kernel<kSeedMaxVpt>, sgrid, dim3(kSeedThreads), 0, st, 1, dump, pooled_i, pooled_i, need, thr));
    return EVDB_OK;
}

int launch_gemm_topk(evdb_store *s, const double *d_q64, int B, int KP, int metric, int *lists_per_query,
                     const float **d_eps_q, RawCands *raw, cudaStream_t st) {
    const bool l2 = metric == EVDB_EUCLIDEAN;
    if (l2) EVDB_TRY(ensure_l2_shadow(s, st));
    const int kcols = s->dim;
    const int kpitch = s->spitch;
    const __half *vcol = l2 ? s->shadow_l2 : s->shadow;
    const float sigma = l2 ? s->l2_sigma : 1.0f;
    const int nblocks_q = (B + GM - 1) / GM;
    // concurrent query blocks: the largest power of two <= 8 that does not exceed what exists
    int MB = 1;
    while (MB * 2 <= nblocks_q && MB * 2 <= 8) MB *= 2;
    const int nchunks = (nblocks_q + MB - 1) / MB;
    if (nchunks > kMaxSweeps) return EVDB_E_BAD_ARG;  // search_core splits larger batches
    const int Bpad = nchunks * MB * GM;
    int NG = s->sm_count / MB;
    const int nt = (int)((s->count + GN - 1) / GN);
    if (NG > nt) NG = nt;
    const int nCTA = MB * NG;
    const int cap = KP <= 32 ? 128 : kCandCapMax;
    // >= 2 concurrent query blocks: CTAs (2p, 2p+1) share a row group and run as one cta_group::2 pair
    // Measured on B200 (tools/sweep.py, r01): +1 % at d = 768, -3 % at d = 128 (the issuer waits for the slower
    // of two epilogues), -4 % at d = 1536: the kernel runs at the power cap, not at the L2 feed limit.
    bool pair = MB >= 2 && kcols >= 512 && kcols <= 1024;
    { const char *e = getenv("EVDB_GEMM_PAIR"); if (e) pair = MB >= 2 && atoi(e) != 0; }
    const uint32_t vbox = pair ? GN / 2 : GN;

    // workspace: [Qh][Q tail][qc0][eps_q][thr0][cand_cnt][cand]
    const size_t qh_bytes = round_up64((size_t)Bpad * kpitch * sizeof(__half), 256);
    const size_t qt_bytes = round_up64((size_t)Bpad * GUK * sizeof(__half), 256);
    const size_t vec_bytes = round_up64((size_t)Bpad * 4, 256);
    const size_t cnt_bytes = round_up64((size_t)nchunks * nCTA * kEpiParts * GM * sizeof(int), 256);
    const size_t cand_bytes = (size_t)nchunks * nCTA * kEpiParts * cap * GM * sizeof(uint64_t);
    EVDB_TRY(ensure_bytes((void **)&s->w_qh, &s->w_qh_cap, qh_bytes + qt_bytes + 3 * vec_bytes + cnt_bytes + cand_bytes));
    uint8_t *wp = (uint8_t *)s->w_qh;
    __half *qh = (__half *)wp; wp += qh_bytes;
    __half *qtail = (__half *)wp; wp += qt_bytes;
    float *qc0 = (float *)wp; wp += vec_bytes;
    float *eps_q = (float *)wp; wp += vec_bytes;
    uint32_t *thr = (uint32_t *)wp; wp += vec_bytes;
    int *cand_cnt = (int *)wp; wp += cnt_bytes;
    uint64_t *cand = (uint64_t *)wp;

    EVDB_CUDA(cudaMemsetAsync(thr, 0, sizeof(uint32_t) * (size_t)Bpad, st));   // seeded thresholds accumulate by atomicMax
    prep_queries_gemm_kernel<<<Bpad, 256, 0, st>>>(d_q64, B, s->dim, metric, sigma, s->max_norm, qh, kpitch,
                                                   l2 ? qtail : nullptr, qc0, eps_q);
    EVDB_CUDA(cudaGetLastError());
    CUtensorMap tmQ, tmV, tmQt, tmVt;
    EVDB_TRY(make_map(&tmQ, qh, (uint64_t)Bpad, (uint64_t)kpitch, (uint64_t)kpitch, GM));
    EVDB_TRY(make_map(&tmV, vcol, s->count, (uint64_t)kpitch, (uint64_t)kpitch, vbox));
    if (l2) {
        EVDB_TRY(make_map(&tmQt, qtail, (uint64_t)Bpad, GUK, GUK, GM, true));
        EVDB_TRY(make_map(&tmVt, s->l2_tail, s->count, GUK, GUK, vbox, true));
    } else {
        tmQt = tmQ;  // never dereferenced
        tmVt = tmV;
    }
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B;
    a.kblocks = (kcols + GK - 1) / GK;
    a.last_ksteps = (kcols - (a.kblocks - 1) * GK + GUK - 1) / GUK;
    a.tail = l2 ? 1 : 0;
    a.cap = cap;
    a.MB = MB;
    a.nchunks = nchunks;
    a.KP = KP;
    a.cand = cand;
    a.cand_cnt = cand_cnt;
    a.qc0 = qc0;
    a.c1 = l2 ? -2.0f / (sigma * sigma) : -1.0f;

    // ---- sampled pre-pass: seed every query's admission threshold ----
    // S strided sample rows (a TMA map with a multiplied row stride), the best key score of every
    // 32-row chunk pooled, KP-th smallest per query selected.  Kills the warm-up of the main sweep.
    const uint32_t *thr0 = nullptr;
    const char *noseed = getenv("EVDB_GEMM_NOSEED");
    uint64_t S = (uint64_t)8192 * KP;   // admitted keys per query ~ count * KP / S: wider windows get a larger sample
    { const char *e = getenv("EVDB_GEMM_SAMPLE"); if (e && atoi(e) >= 1024) S = (uint64_t)atoi(e); }
    while (S * 16 > s->count && S > 1024) S >>= 1;
    if (S >= (uint64_t)64 * KP && S * 16 <= s->count && !(noseed && atoi(noseed))) {
        const uint64_t step = s->count / S;
        const int snt = (int)(S / GN);
        int sNG = s->sm_count / MB;
        if (sNG > snt) sNG = snt;
        const int pooled = (int)(S / 32);
        EVDB_TRY(ensure_bytes((void **)&s->w_seed, &s->w_seed_cap, (size_t)Bpad * pooled * sizeof(float)));
        float *dump = (float *)s->w_seed;
        CUtensorMap tmVs, tmVts = tmVt;
        EVDB_TRY(make_map(&tmVs, vcol, S, (uint64_t)kpitch, (uint64_t)kpitch * step, vbox));
        if (l2) EVDB_TRY(make_map(&tmVts, s->l2_tail, S, GUK, (uint64_t)GUK * step, vbox, true));
        GemmArgs p = a;
        p.n = S; p.nt = snt; p.NG = sNG; p.mode = 1; p.dump = dump; p.dump_ld = pooled;
        EVDB_TRY(launch_gemm_kernel(pair, MB * sNG, st, tmQ, tmVs, tmQt, tmVts, p));
        EVDB_TRY(launch_seed_thresholds(dump, pooled, Bpad, KP, thr, st));   // (thr was zeroed before the query prep)
        s->n_launches += 2;
        thr0 = thr;
    }
    a.thr0 = thr0;
    a.n = s->count;
    a.nt = nt;
    a.NG = NG;
    { const char *e = getenv("EVDB_GEMM_DEBUG"); a.debug = e ? atoi(e) : 0; }
    a.dbg = nullptr;
    static unsigned long long *g_dbg = nullptr;
    if (a.debug & 8) {
        if (!g_dbg) cudaMalloc((void **)&g_dbg, sizeof(unsigned long long) * 148 * kEpiWarps * 8);
        a.dbg = g_dbg;
    }
    prof_begin(s, st);
    EVDB_TRY(launch_gemm_kernel(pair, nCTA, st, tmQ, tmV, tmQt, tmVt, a));
    prof_end(s, st);
    if (a.debug & 8) {
        cudaStreamSynchronize(st);
        static unsigned long long h[148 * kEpiWarps * 8];
        cudaMemcpy(h, g_dbg, sizeof(h), cudaMemcpyDeviceToHost);
        double sum[5] = {0, 0, 0, 0, 0};
        int nw = nCTA * kEpiWarps;
        for (int i = 0; i < nw; ++i) for (int j = 0; j < 5; ++j) sum[j] += (double)h[i * 8 + j];
        fprintf(stderr, "[gemm dbg] per epilogue warp (cycles): wait=%.0f chunk=%.0f (of which prune=%.0f) prunes=%.1f\n",
                sum[0] / nw, sum[1] / nw, sum[2] / nw, sum[4] / nw);
    }
    s->n_launches += 2;
    raw->cand = cand; raw->cnt = cand_cnt; raw->cap = cap; raw->nCTA = nCTA; raw->MB = MB; raw->NG = NG;
    raw->parts = kEpiParts; raw->gm = GM;
    *lists_per_query = kEpiParts * NG;
    *d_eps_q = eps_q;
    return EVDB_OK;
}

}  // namespace evdb
