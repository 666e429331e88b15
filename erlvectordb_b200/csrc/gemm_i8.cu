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
// Kernel anatomy (one persistent CTA per SM, 384 threads, no shadow column -- the codes ARE the operand):
//   warp 0   TMA producer: per 128-byte K block a [128 x 128 B] tile of each digit plane of the CTA's
//            query block and a [128 rows x 128 B] tile of codes (128B swizzle) -> 4-stage ring; for
//            d <= 384 the query planes are loaded ONCE per sweep and stay resident, the ring (6-10
//            stages) carries codes only
//   warp 1   MMA issuer: per K block up to 4 K-steps (K = 32) x 2 planes, M = 128, N = 128
//   warp 2   TMEM allocator (512 columns = 2 accumulator stages x {Sa[128], Sb[128]})
//   warps 4-11  epilogue (8 warps, 168 registers): thread <-> TMEM lane <-> query; the 2 warps of a lane quarter
//            take 64 of the tile's 128 rows each, as two 32-row chunks whose high-digit sums are fetched together
//            (a 16-warp, one-chunk variant is kept for A/B).  Per chunk: a coarse filter on Sa alone (d <= 128),
//            then -- only if some lane's bound beats its threshold -- tcgen05.ld of Sb, per column
//            S = 256*Sa + Sb in fp32, x = cx*S + cy*sum(Q) with the row's coefficients (staged per warp in
//            shared memory), and the shared accumulator-domain filter / append / prune (tc05.cuh) with the
//            per-query key = 1 - x * fx/||q||.
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
constexpr uint32_t kQPlaneBytes = GM * QKB;            // 16 KB: one digit plane of the query block, one K block
constexpr uint32_t kQCodeBytes = QN * QKB;             // 16 KB: 128 rows of codes, one K block
constexpr uint32_t kQRingBytes = 192 * 1024;           // operand shared memory: resident query planes + the ring
// RES (d <= 384): the CTA's query block stays in shared memory for a whole sweep -- both digit planes, all K
// blocks, loaded once -- and the ring carries codes only (16 KB stages): a third of the L2 -> shared-memory
// traffic of streaming {Qa, Qb, codes} per K block.  Measured +-0 at d = 96, where the epilogue paces the
// kernel (EVDB_QGEMM_RES=0 for the A/B); kept because the traffic it removes is what bounds the long-row shapes.
constexpr int kQResMaxKBlocks = 3;
constexpr int kQMaxStages = 10;
constexpr int kQBars = 2 * kQMaxStages + 6;            // full[S] empty[S] tfull[2] tempty[2] qfull qempty
constexpr size_t kQCoefBytes = (size_t)kEpiWarps * 32 * sizeof(float4);   // per epilogue warp: 32 rows x {256 cx, K', cy, cx}
constexpr size_t kQSmem = 1024 + (size_t)kQRingBytes + kQCoefBytes + kQBars * 8 + 16;
constexpr int kQMaxDim = 16384;                        // d * 255 * 255 < 2^31

struct QGemmArgs {
    uint64_t n;          // corpus rows
    int B;               // live queries
    int kblocks;         // ceil(dpad / 128)
    int last_ksteps;     // K-steps in the last K block (1..4)
    int cap;             // candidate buffer capacity in use
    int nt;              // corpus tiles = ceil(n / 128)
    int MB, NG, nchunks, KP;
    int stages;          // ring depth: 4 x {Qa, Qb, codes} streamed, or (192 KB - resident planes) / 16 KB of codes
    uint64_t *cand;      // [sweep][CTA][part][cap][128]
    int *cand_cnt;       // [sweep][CTA][part][128]
    const QStat *qstat;  // [B]
    const float2 *qcoef; // [n * coef_step] {scale, min}/||y|| (the sampled pre-pass reads every coef_step-th row's)
    uint64_t coef_step;
    int mode;            // 0 = fused top-k, 1 = pooled key scores of a strided row sample (threshold seeding)
    float *dump;         // mode 1: [Bpad][dump_ld] best key score of every 32-row chunk
    int dump_ld;
    const uint32_t *thr0;  // mode 0: per-query starting threshold, orderable key score (NULL = none)
    int coarse;          // FAST: run the high-digit coarse filter before fetching the low-digit sums
    float sb_max;        // FAST: upper bound of the low-digit sum, 255 * 255 * d, plus the slack of the coarse filter
    int debug;           // EVDB_QGEMM_DEBUG (measurement only): 1 = no epilogue arithmetic, 2 = one MMA per tile,
                         // 4 = no TMEM loads, 8 = no code TMA after the first tile, 16 = no coarse filter
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
template <bool FAST, bool RES, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_i8_topk_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                    const __grid_constant__ CUtensorMap tmV, const QGemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kStages = a.stages;
    const uint32_t kStageBytes = RES ? kQCodeBytes : 2 * kQPlaneBytes + kQCodeBytes;
    uint8_t *qres_base = smem;                                                 // RES: [K block][Qa | Qb]
    uint8_t *stage_base = smem + (RES ? (size_t)a.kblocks * 2 * kQPlaneBytes : 0);   // [stages][codes] or [stages][Qa | Qb | codes]
    float4 *coef_base = reinterpret_cast<float4 *>(smem + kQRingBytes);        // [epilogue warp][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kQRingBytes + kQCoefBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kQBars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KP = a.KP;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kQMaxStages);
    const uint32_t tfull0 = smem_u32(bars + 2 * kQMaxStages), tempty0 = smem_u32(bars + 2 * kQMaxStages + 2);
    const uint32_t qfull = smem_u32(bars + 2 * kQMaxStages + 4), qempty = smem_u32(bars + 2 * kQMaxStages + 5);

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, EW);
        }
        mbar_init(qfull, 1);
        mbar_init(qempty, 1);
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
                if (RES) {
                    // the sweep's query planes: once the previous sweep's MMAs have retired
                    mbar_wait(qempty, (uint32_t)(c & 1) ^ 1);
                    mbar_arrive_expect_tx(qfull, (uint32_t)a.kblocks * 2 * kQPlaneBytes);
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        const uint32_t qa = smem_u32(qres_base + (size_t)kb * 2 * kQPlaneBytes);
                        tma_load_2d(qa, &tmQa, qfull, kb * QKB, qrow);
                        tma_load_2d(qa + kQPlaneBytes, &tmQb, qfull, kb * QKB, qrow);
                    }
                }
                for (int t = 0; t < my_tiles; ++t) {
                    const int vrow = (ng + t * a.NG) * QN;
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kStageBytes);
                        if (RES && (a.debug & 8) && t > 0) {   // measurement only: stale codes, no TMA traffic
                            mbar_arrive(full0 + 8 * stage);
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(full0 + 8 * stage, kStageBytes);
                        if (RES) {
                            tma_load_2d(sa, &tmV, full0 + 8 * stage, kb * QKB, vrow);
                        } else {
                            tma_load_2d(sa, &tmQa, full0 + 8 * stage, kb * QKB, qrow);
                            tma_load_2d(sa + kQPlaneBytes, &tmQb, full0 + 8 * stage, kb * QKB, qrow);
                            tma_load_2d(sa + 2 * kQPlaneBytes, &tmV, full0 + 8 * stage, kb * QKB, vrow);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
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
                if (RES) {
                    mbar_wait(qfull, (uint32_t)(c & 1));
                    tc_fence_after();
                }
                for (int t = 0; t < my_tiles; ++t) {
                    mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);  // epilogue drained this accumulator pair
                    tc_fence_after();
                    const uint32_t tmem_a = tmem_base + (uint32_t)acc * (2 * QN);
                    const uint32_t tmem_b = tmem_a + QN;
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kStageBytes);
                        const uint32_t qa = RES ? smem_u32(qres_base + (size_t)kb * 2 * kQPlaneBytes) : sa;
                        const uint64_t adesc = make_sw128_kmajor_desc(qa);
                        const uint64_t bdesc = make_sw128_kmajor_desc(qa + kQPlaneBytes);
                        const uint64_t vdesc = make_sw128_kmajor_desc(RES ? sa : sa + 2 * kQPlaneBytes);
                        const int ksteps = kb + 1 == a.kblocks ? a.last_ksteps : QKB / QUK;
#pragma unroll
                        for (int k = 0; k < QKB / QUK; ++k) {
                            // advance 32 codes = 32 bytes along K inside the swizzled row: +2 in >>4 units
                            if (k < ksteps && !((a.debug & 2) && (kb | k) != 0)) {
                                const uint32_t accum = (uint32_t)((kb | k) != 0);
                                umma_i8(tmem_a, adesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_a, accum);
                                umma_i8(tmem_b, bdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_b, accum);
                            }
                        }
                        umma_commit(empty0 + 8 * stage);   // stage reusable once these MMAs retire
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(tfull0 + 8 * acc);         // both accumulators complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                if (RES) umma_commit(qempty);              // the resident planes may be replaced
            }
        }
    } else if (active && warp >= 4) {
        // ===== epilogue =====
        // EW epilogue warps: 16 (one 32-row chunk of every tile per warp) or 8 (two chunks per warp, 168 registers per thread)
        constexpr int kParts = EW / 4;           // candidate lists per query and CTA
        constexpr int kCpw = 4 / kParts;         // 32-row chunks per warp and tile
        const int ew = warp - 4;                 // 0..EW-1
        const int lg = ew & 3;                   // == warp % 4: TMEM lanes [32*lg, 32*lg+32)
        const int part = ew >> 2;                // rows [32*kCpw*part, 32*kCpw*(part+1)) of every tile
        const int et = lg * 32 + lane;           // query within the CTA's block
        const float kInf = __int_as_float(0x7f800000);
        const uint32_t wcoef = smem_u32(coef_base + ew * 32);   // [32] float4; explicit shared-space accesses (the realigned base pointer is generic)
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t nrows = (uint32_t)a.n;
        for (int c = 0; c < a.nchunks; ++c) {
            const size_t lbase = ((size_t)c * nCTA + cta) * kParts + part;
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
            const bool admit = live && a.mode == 0;
            // admission threshold: seeded by the sampled pre-pass, tightened by every prune; held in the x domain
            float tau = (a.thr0 && live) ? f32_from_orderable(a.thr0[qglob]) : kInf;
            float thrS = acc_threshold(tau, c0, c1);
            // the coefficients of the warp's next 32-row chunks, fetched two chunks ahead (a dependent global load
            // per chunk would put its whole latency on every chunk of the warp)
            auto load_coef = [&](int st) -> float2 {   // step st = chunk st % kCpw of the warp's tile st / kCpw
                const int t = st / kCpw;
                const uint32_t r = (uint32_t)(ng + t * a.NG) * QN + (uint32_t)(part * kCpw + st % kCpw) * 32 + lane;
                return (t < my_tiles && r < nrows) ? __ldg(a.qcoef + (size_t)r * a.coef_step) : make_float2(0.f, 0.f);
            };
            float2 co_next = load_coef(0), co_next2 = load_coef(1);
            const bool coarse = FAST && a.mode == 0 && a.coarse && !(a.debug & 16);
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = ng + t * a.NG;
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                tc_fence_after();
                // one 32-row chunk; PRE: its high-digit sums were fetched before the loop (both chunks' loads in flight together)
                auto do_chunk = [&](const int cb, uint32_t (&va)[32], const bool pre) {
                const int chunk = part * kCpw + cb;      // 32-row chunk of the tile
                const uint32_t row0 = (uint32_t)tile * QN + (uint32_t)chunk * 32;
                {
                    // row constants of the coarse filter (FAST): with F = float bits of (Sa + 0x4B400000) = 12582912 + Sa,
                    //   x <= cx*(256 Sa + SbMax) + cy*Cq = (256 cx)*F + (cy*Cq + K'),  K' = cx*SbMax' - 256 cx * 12582912
                    // K' is formed in fp64 and rounded UP; SbMax' carries the slack that covers every fp32 rounding of
                    // both evaluations (2048 units of S per cx, 1.5 |cy| for the cy*Cq products, |Cq| <= 2^22)
                    float kpf = 0.f;
                    if (coarse) {
                        const double cx = (double)co_next.x, cy = (double)co_next.y;
                        const double kp = cx * (double)a.sb_max + 1.5 * fabs(cy) - cx * 256.0 * 12582912.0;
                        kpf = (float)kp;
                        if ((double)kpf < kp) kpf = __int_as_float(__float_as_int(kpf) + (kpf >= 0.f ? 1 : -1));
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wcoef + lane * 16), "f"(co_next.x * 256.0f),
                                 "f"(kpf), "f"(co_next.y), "f"(co_next.x) : "memory");   // {256 cx, K', cy, cx}
                }
                __syncwarp();
                co_next = co_next2;
                co_next2 = load_coef(t * kCpw + cb + 2);
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)acc * (2 * QN) + chunk * 32;
                if (!pre && !(a.debug & 4)) {
                    tmem_ld_32x32b_x32(taddr, va);
                    tmem_ld_wait();
                }
                bool exact = !coarse;
                if (coarse && !(a.debug & 1)) {
                    // high digit only: an upper bound of every x of the chunk, three instructions per row
                    float g[8];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float c256, cy, kp, cx;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c256), "=f"(kp), "=f"(cy), "=f"(cx) : "r"(wcoef + j * 16));
                        const float F = __int_as_float((int)va[j] + 0x4B400000);
                        const float xu = fmaf(c256, F, fmaf(cy, Cq, kp));
                        g[j >> 2] = (j & 3) ? fmaxf(g[j >> 2], xu) : xu;
                    }
                    const float mu = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
                    exact = __any_sync(0xffffffffu, admit && mu > thrS);   // (padding lanes of a ragged batch hold no threshold)
                }
                float m = 0.f;
                if (exact && !(a.debug & 1)) {
                    uint32_t vb[32];
                    tmem_ld_32x32b_x32(taddr + QN, vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float fa, fb, cy, cx;
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cy), "=f"(cx) : "r"(wcoef + j * 16 + 8));
                        if (FAST) {
                            fa = __int_as_float((int)va[j] + 0x4B400000) - 12582912.0f;   // |Sa| < 2^22
                            fb = __int_as_float((int)vb[j] + 0x4B000000) - 8388608.0f;    // 0 <= Sb < 2^23
                        } else {
                            fa = __int2float_rn((int)va[j]);
                            fb = __int2float_rn((int)vb[j]);
                        }
                        const float S = fmaf(fa, 256.0f, fb);
                        vb[j] = __float_as_uint(fmaf(cx, S, cy * Cq));
                    }
                    m = epi_chunk(vb, thrS, row0, nrows, c0, c1, mybuf, cnt, admit);
                }
                __syncwarp();   // wcoef is rewritten at the top of the next chunk
                if (cb == kCpw - 1) {
                    // hand the TMEM stage back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                if (a.mode == 1) {   // sampled pre-pass: best key score of the chunk
                    a.dump[qglob * a.dump_ld + (size_t)tile * 4 + chunk] = fmaf(m, c1, c0);
                    return;
                }
                if (exact) {
                    const unsigned need = __ballot_sync(0xffffffffu, cnt > a.cap - 32);
                    if (need) prune_buffers(cbase + lg * 32, need, KP, lane, c0, c1, cnt, tau, thrS);
                }
                };
                if constexpr (kCpw == 2) {
                    uint32_t v0[32], v1[32];
                    const uint32_t t0 = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)acc * (2 * QN) + part * kCpw * 32;
                    tmem_ld_32x32b_x32(t0, v0);
                    tmem_ld_32x32b_x32(t0 + 32, v1);
                    tmem_ld_wait();
                    do_chunk(0, v0, true);
                    do_chunk(1, v1, true);
                } else {
#pragma unroll 1
                    for (int cb = 0; cb < kCpw; ++cb) {
                        uint32_t va[32];
                        do_chunk(cb, va, false);
                    }
                }
            }
            if (a.mode == 0) a.cand_cnt[lbase * GM + et] = cnt;
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

    const int Bpad = nchunks * MB * GM;
    const size_t thr_bytes = round_up64((size_t)Bpad * sizeof(uint32_t), 256);
    const size_t cnt_bytes = round_up64((size_t)nchunks * nCTA * kEpiParts * GM * sizeof(int), 256);
    const size_t cand_bytes = (size_t)nchunks * nCTA * kEpiParts * cap * GM * sizeof(uint64_t);
    EVDB_TRY(ensure_bytes((void **)&s->w_qh, &s->w_qh_cap, thr_bytes + cnt_bytes + cand_bytes));
    uint32_t *thr = (uint32_t *)s->w_qh;
    int *cand_cnt = (int *)((uint8_t *)s->w_qh + thr_bytes);
    uint64_t *cand = (uint64_t *)((uint8_t *)s->w_qh + thr_bytes + cnt_bytes);
    EVDB_CUDA(cudaMemsetAsync(thr, 0, sizeof(uint32_t) * (size_t)Bpad, st));   // seeded thresholds accumulate by atomicMax

    EVDB_TRY(launch_prep_queries(s, d_q64, B, EVDB_COSINE, st));   // [B][2][dpad] digits, QStat, grid bound
    CUtensorMap tmQa, tmQb, tmV;
    const uint64_t dp = (uint64_t)s->dpad;
    EVDB_TRY(make_map_u8(&tmQa, s->w_qdig, (uint64_t)B, dp, (uint64_t)kQPlanes * dp));
    EVDB_TRY(make_map_u8(&tmQb, s->w_qdig + dp, (uint64_t)B, dp, (uint64_t)kQPlanes * dp));
    // the code tiles may run past a row's end (into the next row; the allocation carries slack for the last one):
    // the digit planes are zero there, so the products vanish.  The alternative (EVDB_QGEMM_VEXT=0) lets TMA's
    // out-of-bounds fill pad rows shorter than the 128-byte box; measured equal at 12.5 M x 96 (10.40 / 10.54 ms).
    static int vext = -1;
    if (vext < 0) { const char *e = getenv("EVDB_QGEMM_VEXT"); vext = e ? atoi(e) : 1; }
    const uint64_t vcols = vext ? (dp + QKB - 1) / QKB * QKB : dp;
    EVDB_TRY(make_map_u8(&tmV, s->rows, s->count, vcols, (uint64_t)s->row_bytes));

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
    a.coef_step = 1;
    a.sb_max = 65025.0f * (float)s->dim + 2048.0f;
    { const char *e = getenv("EVDB_QGEMM_DEBUG"); a.debug = e ? atoi(e) : 0; }
    const bool fast = s->dim <= 128;
    bool use_res = a.kblocks <= kQResMaxKBlocks;
    { const char *e = getenv("EVDB_QGEMM_RES"); if (e) use_res = use_res && atoi(e) != 0; }   // A/B: 0 = always stream the planes
    a.stages = use_res ? (int)((kQRingBytes - (size_t)a.kblocks * 2 * kQPlaneBytes) / kQCodeBytes) : 4;
    if (a.stages > kQMaxStages) a.stages = kQMaxStages;
    // epilogue warps: 8 x 168 registers, two 32-row chunks of every tile per warp with both TMEM loads in flight together
    // (no spills, room for instruction-level parallelism), or 16 x 96 registers, one chunk each (EVDB_QGEMM_EW=16).
    // Same-box A/B, kernel ms at 16 / 8 warps: 12.5 M x 96 B = 1024 10.6 / 8.35, B = 64 1.45 / 1.02; 4 M x 256 4.31 / 2.68;
    // 1 M x 768 1.57 / 1.46 (4 warps x 221 registers: 13.0 at 12.5 M x 96)
    static int ew = -1;
    if (ew < 0) { const char *e = getenv("EVDB_QGEMM_EW"); ew = (e && atoi(e) == 16) ? 16 : 8; }
    const int parts = ew / 4;
    // the coarse filter pays when registers are short (16 warps: 11.5 -> 10.5 ms at 12.5 M x 96) and when few lanes of a
    // warp hold a live query (8 warps, B = 8: 0.97 -> 0.85 ms); with 168 registers and full warps the exact path straight
    // away is faster (B = 1024: 8.59 -> 8.18 ms, k = 100: 9.85 -> 8.51, 4 M x 128: 3.28 -> 2.79)
    static int coarse_env = -2;
    if (coarse_env == -2) { const char *e = getenv("EVDB_QGEMM_COARSE"); coarse_env = e ? atoi(e) : -1; }
    a.coarse = coarse_env >= 0 ? coarse_env : ((ew == 16 || B < 64) ? 1 : 0);
    void (*fn)(CUtensorMap, CUtensorMap, CUtensorMap, QGemmArgs) = nullptr;
#define EVDB_QSEL(EWN)                                                                                             \
    fn = use_res ? (fast ? gemm_i8_topk_kernel<true, true, EWN> : gemm_i8_topk_kernel<false, true, EWN>)           \
                 : (fast ? gemm_i8_topk_kernel<true, false, EWN> : gemm_i8_topk_kernel<false, false, EWN>)
    if (ew == 8) { EVDB_QSEL(8); } else { EVDB_QSEL(16); }
#undef EVDB_QSEL
    const int threads = 128 + 32 * ew;
    EVDB_TRY(ensure_func_smem((const void *)fn, kQSmem));
    // ---- sampled pre-pass (as in gemm_tcgen05.cu): S strided rows pooled per 32-row chunk seed every query's threshold ----
    const char *noseed = getenv("EVDB_GEMM_NOSEED");
    uint64_t S = (uint64_t)8192 * KP;
    while (S * 16 > s->count && S > 1024) S >>= 1;
    if (S >= (uint64_t)64 * KP && S * 16 <= s->count && !(noseed && atoi(noseed))) {
        const uint64_t step = s->count / S;
        const int snt = (int)(S / QN);
        int sNG = s->sm_count / MB;
        if (sNG > snt) sNG = snt;
        const int pooled = (int)(S / 32);
        EVDB_TRY(ensure_bytes((void **)&s->w_seed, &s->w_seed_cap, (size_t)Bpad * pooled * sizeof(float)));
        CUtensorMap tmVs;
        EVDB_TRY(make_map_u8(&tmVs, s->rows, S, vcols, (uint64_t)s->row_bytes * step));
        QGemmArgs p = a;
        p.n = S; p.nt = snt; p.NG = sNG; p.mode = 1; p.dump = (float *)s->w_seed; p.dump_ld = pooled; p.coef_step = step;
        EVDB_CUDA(launch_chained(fn, dim3(MB * sNG), dim3(threads), kQSmem, st, 1, tmQa, tmQb, tmVs, p));
        EVDB_TRY(launch_seed_thresholds((const float *)s->w_seed, pooled, Bpad, KP, thr, st));
        s->n_launches += 2;
        a.thr0 = thr;
    }
    prof_begin(s, st);
    EVDB_CUDA(launch_chained(fn, dim3(nCTA), dim3(threads), kQSmem, st, 1, tmQa, tmQb, tmV, a));
    prof_end(s, st);
    s->n_launches += 1;
    raw->cand = cand; raw->cnt = cand_cnt; raw->cap = cap; raw->nCTA = nCTA; raw->MB = MB; raw->NG = NG;
    raw->parts = parts; raw->gm = GM;
    *lists_per_query = parts * NG;
    *d_eps_q = s->w_qeps;
    return EVDB_OK;
}

}  // namespace evdb
