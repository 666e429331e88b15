// gemm_i8.cu -- query BATCHES against a quantization_8bit store on the tensor cores (family B'):
// tcgen05.mma.kind::i8 over the stored u8 codes and the query's 8-bit digit planes, int32
// accumulators in TMEM, the same fused per-query top-k epilogue as gemm_tcgen05.cu.
//
// Replaces, for batches, what the dp4a scan (scan.cu) replaces for one query: the fold of
// cosine_distance/2 (reference src/vector_store.erl:227-252) over rows that
// vector_persistence:decompress_if_needed (src/vector_persistence.erl:276-284) rebuilt as
// y_i = Min + c_i*Scale (src/vector_compression.erl:180-183).  Same algebra, same integers:
//     q.y = Min*sum(q) + Scale*sum(q_i c_i),  q on the 16-bit fixed-point grid of prep_queries_kernel,
//     Q = a*2^8 + b (a = signed high digit, b = unsigned low digit)
//     Sa[q][r] = sum_k a[q][k]*c[r][k]   (MMA 1: A = s8, B = u8, D = s32)
//     Sb[q][r] = sum_k b[q][k]*c[r][k]   (MMA 2: A = u8, B = u8, D = s32)
// Sa and Sb are the EXACT integers the dp4a scan forms (evdb_debug_quant_dots); the epilogue combines
// them with the row's {scale, min}/||y|| into the candidate key in fp32.  The keys only GENERATE
// CANDIDATES: select.cu re-ranks them in exact fp64 from the codes and proves the window complete.
//
// Kernel anatomy (one persistent CTA per SM, 640 threads, no shadow column -- the codes ARE the operand):
//   warp 0   TMA producer: per 128-byte K block a [128 x 128 B] tile of each digit plane of the CTA's
//            query block and a [128 rows x 128 B] tile of codes (128B swizzle) -> 4-stage ring
//   warp 1   MMA issuer: per K block up to 4 K-steps (K = 32) x 2 planes, M = 128, N = 128
//   warp 2   TMEM allocator (512 columns = 2 accumulator stages x {Sa[128], Sb[128]})
//   warps 4-19  epilogue: thread <-> TMEM lane <-> query; the 4 warps of a lane quarter take 32 of the
//            tile's 128 rows each: two tcgen05.ld (Sa, Sb), per column S = 256*Sa + Sb in fp32,
//            x = cx*S + cy*sum(Q) with the row's coefficients (staged per warp in shared memory),
//            then the shared accumulator-domain filter / append / prune (tc05.cuh) with the per-query
//            key = 1 - x * fx/||q||.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"
#include "topk.cuh"
#include "tc05.cuh"

namespace evdb {

constexpr int QN = 128;        // corpus rows per tile (UMMA N)
constexpr int QKB = 128;       // K bytes per stage (one 128-byte swizzle row of codes)
constexpr int QUK = 32;        // UMMA K of kind::i8
constexpr int kQStages = 4;
constexpr uint32_t kQPlaneBytes = GM * QKB;            // 16 KB: one digit plane of the query block
constexpr uint32_t kQCodeBytes = QN * QKB;             // 16 KB: 128 rows of codes
constexpr uint32_t kQStageBytes = 2 * kQPlaneBytes + kQCodeBytes;
constexpr int kQBars = 2 * kQStages + 4;               // full[S] empty[S] tfull[2] tempty[2]
constexpr size_t kQCoefBytes = (size_t)kEpiWarps * 32 * sizeof(float2);
constexpr size_t kQSmem = 1024 + (size_t)kQStages * kQStageBytes + kQCoefBytes + kQBars * 8 + 16;
constexpr int kQMaxDim = 16384;                        // d * 255 * 255 < 2^31

struct QGemmArgs {
    uint64_t n;          // corpus rows
    int B;               // live queries
    int kblocks;         // ceil(dpad / 128)
    int last_ksteps;     // K-steps in the last K block (1..4)
    int cap;             // candidate buffer capacity in use
    int nt;              // corpus tiles = ceil(n / 128)
    int MB, NG, nchunks, KP;
    uint64_t *cand;      // [sweep][CTA][part][cap][128]
    int *cand_cnt;       // [sweep][CTA][part][128]
    const QStat *qstat;  // [B]
    const float2 *qcoef; // [n] {scale, min}/||y||
};

// kind::i8 instruction descriptor: (u8|s8) x u8 -> s32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(bool a_signed, int M, int N) {
    return (2u << 4)                              // c_format = S32
         | ((a_signed ? 1u : 0u) << 7)            // a_format: 0 = u8, 1 = s8
         | (0u << 10)                             // b_format = u8
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// FAST: |Sa| < 2^22 and 0 <= Sb < 2^23 (d <= 128): int -> float by the mantissa trick (two full-rate
// instructions instead of a quarter-rate I2F); exact either way below 2^24.
template <bool FAST>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_topk_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                    const __grid_constant__ CUtensorMap tmV, const QGemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *stage_base = smem;                                                // [stages][Qa | Qb | codes]
    float2 *coef_base = reinterpret_cast<float2 *>(smem + kQStages * kQStageBytes);   // [epilogue warp][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kQStages * kQStageBytes + kQCoefBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kQBars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KP = a.KP;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kQStages);
    const uint32_t tfull0 = smem_u32(bars + 2 * kQStages), tempty0 = smem_u32(bars + 2 * kQStages + 2);

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kQStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                // the digit planes and query statistics come from prep_queries_kernel
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
            int stage = 0;
            uint32_t phase = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                int qrow = (c * a.MB + mb_local) * GM;
                if (qrow >= a.B) qrow = 0;   // a padding block of the last sweep: any rows will do, nothing is admitted
                for (int t = 0; t < my_tiles; ++t) {
                    const int vrow = (ng + t * a.NG) * QN;
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_u32(stage_base + stage * kQStageBytes);
                        mbar_arrive_expect_tx(full0 + 8 * stage, kQStageBytes);
                        tma_load_2d(sa, &tmQa, full0 + 8 * stage, kb * QKB, qrow);
                        tma_load_2d(sa + kQPlaneBytes, &tmQb, full0 + 8 * stage, kb * QKB, qrow);
                        tma_load_2d(sa + 2 * kQPlaneBytes, &tmV, full0 + 8 * stage, kb * QKB, vrow);
                        if (++stage == kQStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (active && warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc_a = make_idesc_i8(true, GM, QN);    // signed high digit
            constexpr uint32_t idesc_b = make_idesc_i8(false, GM, QN);   // unsigned low digit
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                for (int t = 0; t < my_tiles; ++t) {
                    mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);  // epilogue drained this accumulator pair
                    tc_fence_after();
                    const uint32_t tmem_a = tmem_base + (uint32_t)acc * (2 * QN);
                    const uint32_t tmem_b = tmem_a + QN;
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + stage * kQStageBytes);
                        const uint64_t adesc = make_sw128_kmajor_desc(sa);
                        const uint64_t bdesc = make_sw128_kmajor_desc(sa + kQPlaneBytes);
                        const uint64_t vdesc = make_sw128_kmajor_desc(sa + 2 * kQPlaneBytes);
                        const int ksteps = kb + 1 == a.kblocks ? a.last_ksteps : QKB / QUK;
#pragma unroll
                        for (int k = 0; k < QKB / QUK; ++k) {
                            // advance 32 codes = 32 bytes along K inside the swizzled row: +2 in >>4 units
                            if (k < ksteps) {
                                const uint32_t accum = (uint32_t)((kb | k) != 0);
                                umma_i8(tmem_a, adesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_a, accum);
                                umma_i8(tmem_b, bdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_b, accum);
                            }
                        }
                        umma_commit(empty0 + 8 * stage);   // stage reusable once these MMAs retire
                        if (++stage == kQStages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(tfull0 + 8 * acc);         // both accumulators complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (active && warp >= 4) {
        // ===== epilogue =====
        const int ew = warp - 4;                 // 0..15
        const int lg = ew & 3;                   // == warp % 4: TMEM lanes [32*lg, 32*lg+32)
        const int part = ew >> 2;                // rows [32*part, 32*part+32) of every tile
        const int et = lg * 32 + lane;           // query within the CTA's block
        const float kInf = __int_as_float(0x7f800000);
        float2 *wcoef = coef_base + ew * 32;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t nrows = (uint32_t)a.n;
        for (int c = 0; c < a.nchunks; ++c) {
            const size_t lbase = ((size_t)c * nCTA + cta) * kEpiParts + part;
            uint64_t *cbase = a.cand + lbase * a.cap * GM;
            uint64_t *mybuf = cbase + et;
            int cnt = 0;
            const size_t qglob = (size_t)(c * a.MB + mb_local) * GM + et;
            const bool live = qglob < (size_t)a.B;
            // key = 1 - (cx*S*fx + cy*sum(q^)) / ||q|| = fma(x, c1, 1), x = cx*S + cy*Cq, c1 = -fx/||q|| (per query)
            float c1 = -1.0f, Cq = 0.f;
            if (live) {
                const QStat qs = a.qstat[qglob];
                Cq = qs.sum / qs.fx;                       // sum of the fixed-point query, exact scaling
                c1 = -(qs.inv_norm * qs.fx);
                if (!(c1 < 0.f)) c1 = -1e-30f;             // zero query: every key is 1.0
            }
            const float c0 = 1.0f;
            float tau = kInf;
            float thrS = -kInf;
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = ng + t * a.NG;
                const uint32_t row0 = (uint32_t)tile * QN + (uint32_t)part * 32;
                // this warp's 32 rows' coefficients: fetched before the accumulator is awaited
                {
                    const uint32_t r = row0 + lane;
                    wcoef[lane] = r < nrows ? __ldg(a.qcoef + r) : make_float2(0.f, 0.f);
                }
                __syncwarp();
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)acc * (2 * QN) + part * 32;
                uint32_t va[32], vb[32];
                tmem_ld_32x32b_x32(taddr, va);
                tmem_ld_32x32b_x32(taddr + QN, vb);
                tmem_ld_wait();
                // both accumulators are in registers: hand the TMEM stage back before the arithmetic
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float fa, fb;
                    if (FAST) {
                        fa = __int_as_float((int)va[j] + 0x4B400000) - 12582912.0f;   // |Sa| < 2^22
                        fb = __int_as_float((int)vb[j] + 0x4B000000) - 8388608.0f;    // 0 <= Sb < 2^23
                    } else {
                        fa = __int2float_rn((int)va[j]);
                        fb = __int2float_rn((int)vb[j]);
                    }
                    const float2 co = wcoef[j];
                    const float S = fmaf(fa, 256.0f, fb);
                    va[j] = __float_as_uint(fmaf(co.x, S, co.y * Cq));
                }
                __syncwarp();   // wcoef is rewritten at the top of the next tile
                (void)epi_chunk(va, thrS, row0, nrows, c0, c1, mybuf, cnt, live);
                const unsigned need = __ballot_sync(0xffffffffu, cnt > a.cap - 32);
                if (need) prune_buffers(cbase + lg * 32, need, KP, lane, c0, c1, cnt, tau, thrS);
            }
            a.cand_cnt[lbase * GM + et] = cnt;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// [rows][cols] bytes, `pitch` bytes between rows, box = 128 rows x 128 bytes, 128B swizzle
static int make_map_u8(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch) {
    encode_tiled_fn enc = get_encode();
    if (!enc) return EVDB_E_CUDA;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch};
    cuuint32_t box[2] = {(cuuint32_t)QKB, (cuuint32_t)GM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? EVDB_OK : EVDB_E_CUDA;
}

bool qgemm_plan_supported(evdb_store *s, int metric, int B, int KP) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("EVDB_QGEMM"); on = e ? atoi(e) : 1; }
    if (!on || s->dtype != EVDB_U8 || metric != EVDB_COSINE || kQPlanes != 2) return false;
    if (KP > kGemmMaxKP || B < 1 || s->dim > kQMaxDim) return false;
    if (s->count < (uint64_t)QN) return false;
    return get_encode() != nullptr;
}

// Candidates of a query batch against a U8 store.  Leaves the digit planes / statistics / grid bound of
// launch_prep_queries in the store's workspaces (w_qdig, w_qstat, w_qeps) and the raw candidate
// buffers in w_qh; *d_eps_q = the per-query grid bound (the caller adds the arithmetic bound).
int launch_qgemm_topk(evdb_store *s, const double *d_q64, int B, int KP, int *lists_per_query,
                      const float **d_eps_q, RawCands *raw, cudaStream_t st) {
    const int nblocks_q = (B + GM - 1) / GM;
    int MB = 1;
    while (MB * 2 <= nblocks_q && MB * 2 <= 8) MB *= 2;
    const int nchunks = (nblocks_q + MB - 1) / MB;
    if (nchunks > kMaxSweeps) return EVDB_E_BAD_ARG;  // search_core splits larger batches
    int NG = s->sm_count / MB;
    const int nt = (int)((s->count + QN - 1) / QN);
    if (NG > nt) NG = nt;
    const int nCTA = MB * NG;
    const int cap = KP <= 32 ? 128 : kCandCapMax;

    const size_t cnt_bytes = round_up64((size_t)nchunks * nCTA * kEpiParts * GM * sizeof(int), 256);
    const size_t cand_bytes = (size_t)nchunks * nCTA * kEpiParts * cap * GM * sizeof(uint64_t);
    EVDB_TRY(ensure_bytes((void **)&s->w_qh, &s->w_qh_cap, cnt_bytes + cand_bytes));
    int *cand_cnt = (int *)s->w_qh;
    uint64_t *cand = (uint64_t *)((uint8_t *)s->w_qh + cnt_bytes);

    EVDB_TRY(launch_prep_queries(s, d_q64, B, EVDB_COSINE, st));   // [B][2][dpad] digits, QStat, grid bound
    CUtensorMap tmQa, tmQb, tmV;
    const uint64_t dp = (uint64_t)s->dpad;
    EVDB_TRY(make_map_u8(&tmQa, s->w_qdig, (uint64_t)B, dp, (uint64_t)kQPlanes * dp));
    EVDB_TRY(make_map_u8(&tmQb, s->w_qdig + dp, (uint64_t)B, dp, (uint64_t)kQPlanes * dp));
    EVDB_TRY(make_map_u8(&tmV, s->rows, s->count, dp, (uint64_t)s->row_bytes));

    QGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.n = s->count;
    a.B = B;
    a.kblocks = (s->dpad + QKB - 1) / QKB;
    a.last_ksteps = (s->dpad - (a.kblocks - 1) * QKB + QUK - 1) / QUK;
    a.cap = cap;
    a.nt = nt;
    a.MB = MB; a.NG = NG; a.nchunks = nchunks; a.KP = KP;
    a.cand = cand; a.cand_cnt = cand_cnt;
    a.qstat = s->w_qstat;
    a.qcoef = s->qcoef;
    const bool fast = s->dim <= 128;
    const void *fn = fast ? (const void *)gemm_i8_topk_kernel<true> : (const void *)gemm_i8_topk_kernel<false>;
    EVDB_TRY(ensure_func_smem(fn, kQSmem));
    prof_begin(s, st);
    if (fast) EVDB_CUDA(launch_chained(gemm_i8_topk_kernel<true>, dim3(nCTA), dim3(kGemmThreads), kQSmem, st, 1, tmQa, tmQb, tmV, a));
    else EVDB_CUDA(launch_chained(gemm_i8_topk_kernel<false>, dim3(nCTA), dim3(kGemmThreads), kQSmem, st, 1, tmQa, tmQb, tmV, a));
    prof_end(s, st);
    s->n_launches += 1;
    raw->cand = cand; raw->cnt = cand_cnt; raw->cap = cap; raw->nCTA = nCTA; raw->MB = MB; raw->NG = NG;
    raw->parts = kEpiParts; raw->gm = GM;
    *lists_per_query = kEpiParts * NG;
    *d_eps_q = s->w_qeps;
    return EVDB_OK;
}

}  // namespace evdb
