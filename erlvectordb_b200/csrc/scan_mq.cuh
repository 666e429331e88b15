// scan_mq.cuh -- the multi-query float scan (batches on the plans without a GEMM form), instantiated per
// element type in scan_mq_f32.cu / scan_mq_bf16.cu: 54 kernels each, compiled in parallel (one translation
// unit with all 108 took 4.5 minutes).
#pragma once
#include "scan_common.cuh"

namespace evdb {

// ----------------------------------------------------------------------------
// The same scan for Q queries at once (batches on the plans without a GEMM form: manhattan, BF16
// stores): every 16-byte row chunk is loaded ONCE and used against Q queries held in shared memory,
// so a batch of B costs ceil(B/Q) passes over the rows instead of B.  R rows x Q queries accumulate per lane.
// The pass stays HBM-bound while the FP32 work per chunk fits under the load time: ~8 instructions
// per query and chunk for manhattan (sub, |.|+add x 4), 128 FP32 lanes against 23 row bytes per clock
// and SM -> Q <= ~10; beyond that a batched manhattan scan is FP32-pipe-bound.
// ----------------------------------------------------------------------------
template <int METRIC, int DTYPE, int TPR, int Q>
__global__ void __launch_bounds__(kScanWarps * 32, (DTYPE == EVDB_BF16 && Q > 2) ? 1 : 2)
scan_float_mq_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();                // the prepared query comes from prep_queries_kernel
    pdl_launch_dependents();
    // rows in flight per group.  F32: R * Q = 8 accumulators, two CTAs per SM (measured on B200, 1 M x 1536 manhattan:
    // 3.5 k QPS against 2.6 k with 16 accumulators and one CTA per SM); BF16 (two query float4s per chunk: twice the
    // shared-memory reads per row byte) gains from reusing every query value for more rows: 16 accumulators, one CTA
    constexpr int R = DTYPE == EVDB_BF16 ? (Q == 8 ? 2 : 4) : 8 / Q;
    constexpr int QPC = (DTYPE == EVDB_F32) ? 1 : 2;
    constexpr int GPW = 32 / TPR;
    const int nch = a.nch, KP = a.KP;
    const int qf4 = nch * QPC;                       // float4s per query
    const int cap = append_cap(KP);
    float4 *sq = reinterpret_cast<float4 *>(smem);   // [Q][qf4]
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + (size_t)Q * qf4 * 16);   // [Q][kScanWarps * cap]
    const size_t lstride = (size_t)kScanWarps * cap;
    const int b0 = blockIdx.y * Q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int qi = 0; qi < Q; ++qi) {
        const int b = b0 + qi < a.B ? b0 + qi : a.B - 1;   // a ragged last group repeats the last query (its output is not stored)
        const float4 *qsrc = reinterpret_cast<const float4 *>(a.q32 + (size_t)b * a.q32_stride);
        for (int i = threadIdx.x; i < qf4; i += blockDim.x) sq[(size_t)qi * qf4 + i] = qsrc[i];
    }
    __syncthreads();
    float q_inv[Q];
    uint64_t thr[Q];
    int cnt[Q];
#pragma unroll
    for (int qi = 0; qi < Q; ++qi) {
        q_inv[qi] = a.qstat[b0 + qi < a.B ? b0 + qi : a.B - 1].inv_norm;
        thr[qi] = kKeyMax;
        cnt[qi] = 0;
    }
    const int g = lane / TPR, gl = lane % TPR;
    const uint64_t rows_per_wi = (uint64_t)GPW * R;
    const uint64_t total_wi = (a.n + rows_per_wi - 1) / rows_per_wi;
    for (uint64_t wi = (uint64_t)blockIdx.x * kScanWarps + warp; wi < total_wi; wi += (uint64_t)gridDim.x * kScanWarps) {
        const uint64_t base = wi * rows_per_wi;
        const uint4 *rp[R];
        uint64_t rix[R];
        bool valid[R];
        float inv[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint64_t r = base + (uint64_t)j * GPW + g;
            valid[j] = r < a.n;
            rix[j] = r;
            rp[j] = reinterpret_cast<const uint4 *>(a.rows + (valid[j] ? r : a.n - 1) * a.row_bytes);
            inv[j] = (METRIC == EVDB_COSINE && valid[j] && gl == 0) ? __ldg(a.inv_norm + rix[j]) : 0.f;
        }
        float4 acc[R][Q];
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
            for (int qi = 0; qi < Q; ++qi) acc[j][qi] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int c = gl; c < nch; c += TPR) {
            uint4 v[R];
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = ldg_stream_u4(rp[j] + c);
#pragma unroll
            for (int qi = 0; qi < Q; ++qi) {
                if (DTYPE == EVDB_F32) {
                    const float4 q0 = sq[(size_t)qi * qf4 + c];
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        acc_f4<METRIC>(acc[j][qi], make_float4(__uint_as_float(v[j].x), __uint_as_float(v[j].y),
                                                               __uint_as_float(v[j].z), __uint_as_float(v[j].w)), q0);
                } else {
                    const float4 q0 = sq[(size_t)qi * qf4 + 2 * c], q1 = sq[(size_t)qi * qf4 + 2 * c + 1];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        acc_f4<METRIC>(acc[j][qi], bf16x4_lo(v[j]), q0);
                        acc_f4<METRIC>(acc[j][qi], bf16x4_hi(v[j]), q1);
                    }
                }
            }
        }
#pragma unroll
        for (int qi = 0; qi < Q; ++qi) {
            uint64_t *mine = lists + (size_t)qi * lstride + (size_t)warp * cap;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                float sacc = (acc[j][qi].x + acc[j][qi].y) + (acc[j][qi].z + acc[j][qi].w);
#pragma unroll
                for (int o = TPR / 2; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
                float score;
                if (METRIC == EVDB_COSINE) score = (inv[j] == 0.f || q_inv[qi] == 0.f) ? 1.0f : 1.0f - sacc * inv[j] * q_inv[qi];
                else if (METRIC == EVDB_EUCLIDEAN) score = sqrtf(sacc);
                else score = sacc;
                const uint64_t key = (valid[j] && gl == 0) ? make_key(score, (uint32_t)rix[j]) : kKeyMax;
                offer_append(key, thr[qi], mine, cnt[qi], cap, KP, lane);
            }
        }
    }
    // per query: every warp's best <= KP keys compacted to lists_q[w*KP ..], CTA merge, one KP-key list out
#pragma unroll 1
    for (int qi = 0; qi < Q; ++qi) {
        uint64_t *lq = lists + (size_t)qi * lstride;
        uint64_t *mine = lq + (size_t)warp * cap;
        int c = cnt[qi];
        uint64_t t = thr[qi];
        if (c > KP) warp_buf_prune(mine, c, t, KP, lane);
        uint64_t e[kAppendMaxKP / 32];
#pragma unroll
        for (int r = 0; r < kAppendMaxKP / 32; ++r) {
            const int i = r * 32 + lane;
            e[r] = (i < c && i < KP) ? mine[i] : kKeyMax;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kAppendMaxKP / 32; ++r) {
            const int i = r * 32 + lane;
            if (i < KP) lq[(size_t)warp * KP + i] = e[r];
        }
        __syncthreads();
        block_bitonic_sort(lq, kScanWarps * KP);
        if (b0 + qi < a.B)
            for (int i = threadIdx.x; i < KP; i += blockDim.x) a.partial[((size_t)(b0 + qi) * a.G + blockIdx.x) * KP + i] = lq[i];
        __syncthreads();
    }
}

template <int METRIC, int DTYPE, int Q>
static scan_fn_t pick_float_mq_t(int tpr) {
    switch (tpr) {
        case 1: return scan_float_mq_kernel<METRIC, DTYPE, 1, Q>;
        case 2: return scan_float_mq_kernel<METRIC, DTYPE, 2, Q>;
        case 4: return scan_float_mq_kernel<METRIC, DTYPE, 4, Q>;
        case 8: return scan_float_mq_kernel<METRIC, DTYPE, 8, Q>;
        case 16: return scan_float_mq_kernel<METRIC, DTYPE, 16, Q>;
        default: return scan_float_mq_kernel<METRIC, DTYPE, 32, Q>;
    }
}
template <int METRIC, int DTYPE>
static scan_fn_t pick_float_mq(int tpr, int Q) {
    return Q == 8 ? pick_float_mq_t<METRIC, DTYPE, 8>(tpr) : Q == 4 ? pick_float_mq_t<METRIC, DTYPE, 4>(tpr)
                                                                    : pick_float_mq_t<METRIC, DTYPE, 2>(tpr);
}
template <int DTYPE>
static scan_fn_t pick_float_mq_m(int metric, int tpr, int Q) {
    return metric == EVDB_COSINE ? pick_float_mq<EVDB_COSINE, DTYPE>(tpr, Q)
         : metric == EVDB_EUCLIDEAN ? pick_float_mq<EVDB_EUCLIDEAN, DTYPE>(tpr, Q)
                                    : pick_float_mq<EVDB_MANHATTAN, DTYPE>(tpr, Q);
}

}  // namespace evdb
