// exact.cuh -- fp64 distances in the reference's exact operation order.
//
// The reference (src/vector_store.erl:238-252, src/vector_utils.erl:38-43)
// forms every product/difference independently and then folds them left to
// right with lists:sum/1, in IEEE binary64 with no FMA.  A warp reproduces
// that bit for bit: 32 lanes compute the d independent terms in parallel
// (__dmul_rn/__dsub_rn: never contracted), stage them in shared memory, and a
// single lane performs the strictly sequential fold.  Lane 1 folds the query's
// own squares at the same time, so the query norm costs no extra latency.
#pragma once
#include "common.cuh"

namespace evdb {

constexpr int kExactChunk = 128;  // terms staged per round; smem = 2*128 doubles per warp

template <int DTYPE>
__device__ __forceinline__ double row_elem(const uint8_t *row, int i, double mn, double sc) {
    if (DTYPE == EVDB_F32) {
        return (double)reinterpret_cast<const float *>(row)[i];
    } else if (DTYPE == EVDB_BF16) {
        return (double)__bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(row)[i]);
    } else if (DTYPE == EVDB_U8) {
        // decompress_8bit_quantization: Min + (Q * Scale)   (vector_compression.erl:180-183)
        return __dadd_rn(mn, __dmul_rn((double)row[i], sc));
    } else {
        // unpack_4bit_values: first element in the high nibble (vector_compression.erl:321-329)
        uint8_t b = row[i >> 1];
        uint32_t c = (i & 1) ? (b & 0x0Fu) : (b >> 4);
        return __dadd_rn(mn, __dmul_rn((double)c, sc));
    }
}

// Left-to-right sum of cnt staged terms onto acc (one lane).  Terms come from shared memory 16 at
// a time into registers of their own, the next 16 requested before the current 16 are added, so
// the dependent DADD chain (8.6 cycles a step on B200) never waits for a load.
__device__ __forceinline__ double fold_staged(double acc, const double *src, int cnt) {
    constexpr int U = 16;
    int i = 0;
    if (cnt >= U) {
        double t[U];
#pragma unroll
        for (int k = 0; k < U; ++k) t[k] = src[k];
#pragma unroll 1
        for (; i + 2 * U <= cnt; i += U) {
            double n[U];
#pragma unroll
            for (int k = 0; k < U; ++k) n[k] = src[i + U + k];
#pragma unroll
            for (int k = 0; k < U; ++k) acc = __dadd_rn(acc, t[k]);
#pragma unroll
            for (int k = 0; k < U; ++k) t[k] = n[k];
        }
#pragma unroll
        for (int k = 0; k < U; ++k) acc = __dadd_rn(acc, t[k]);
        i += U;
    }
    for (; i < cnt; ++i) acc = __dadd_rn(acc, src[i]);
    return acc;
}

// vector_norm/1 of a stored row (src/vector_store.erl:251-252): returned to every lane.
template <int DTYPE>
__device__ double exact_norm_warp(const uint8_t *row, double mn, double sc, int d, double *sp,
                                  int lane) {
    double s = 0.0;
    for (int base = 0; base < d; base += kExactChunk) {
        int cnt = min(kExactChunk, d - base);
        for (int t = lane; t < cnt; t += kWarp) {
            double v = row_elem<DTYPE>(row, base + t, mn, sc);
            sp[t] = __dmul_rn(v, v);
        }
        __syncwarp();
        if (lane == 0) s = fold_staged(s, sp, cnt);
        __syncwarp();
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    return __dsqrt_rn(s);
}

// Distance of query q (fp64, length d) to a stored row, reference order.
// vnorm = cached exact vector_norm of the row (cosine only).
template <int DTYPE>
__device__ double exact_distance_warp(const uint8_t *row, double mn, double sc, const double *q,
                                      int d, int metric, double vnorm, double *sp, int lane) {
    double s = 0.0, sq = 0.0;
    double *sp2 = sp + kExactChunk;
    constexpr int kPer = kExactChunk / kWarp;  // terms per lane per round
    // operands of the NEXT chunk are loaded while one lane folds the current one
    double nv[kPer], nq[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int t = i * kWarp + lane;
        nv[i] = t < d ? row_elem<DTYPE>(row, t, mn, sc) : 0.0;
        nq[i] = t < d ? q[t] : 0.0;
    }
    for (int base = 0; base < d; base += kExactChunk) {
        int cnt = min(kExactChunk, d - base);
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int t = i * kWarp + lane;
            const double v = nv[i], qq = nq[i];
            if (t < cnt) {
                if (metric == EVDB_COSINE) {
                    sp[t] = __dmul_rn(qq, v);    // dot_product: X*Y
                    sp2[t] = __dmul_rn(qq, qq);  // vector_norm(Query): X*X
                } else if (metric == EVDB_EUCLIDEAN) {
                    double t0 = __dsub_rn(qq, v);  // vector_subtract
                    sp[t] = __dmul_rn(t0, t0);
                } else {
                    sp[t] = fabs(__dsub_rn(qq, v));  // abs(X - Y)
                }
            }
        }
        const int nb = base + kExactChunk;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int t = nb + i * kWarp + lane;
            nv[i] = t < d ? row_elem<DTYPE>(row, t, mn, sc) : 0.0;
            nq[i] = t < d ? q[t] : 0.0;
        }
        __syncwarp();
        // lane 0 folds the products, lane 1 (cosine) the query squares -- in the same instruction stream
        if (lane < (metric == EVDB_COSINE ? 2 : 1)) {
            const double r = fold_staged(lane == 0 ? s : sq, lane == 0 ? sp : sp2, cnt);
            if (lane == 0) s = r; else sq = r;
        }
        __syncwarp();
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    if (metric == EVDB_COSINE) {
        sq = __shfl_sync(0xffffffffu, sq, 1);
        double n1 = __dsqrt_rn(sq);
        double n2 = vnorm;
        if (n1 == 0.0 || n2 == 0.0) return 1.0;  // cosine_distance clauses {0.0,_} / {_,0.0}
        return __dsub_rn(1.0, __ddiv_rn(s, __dmul_rn(n1, n2)));
    } else if (metric == EVDB_EUCLIDEAN) {
        return __dsqrt_rn(s);
    }
    return s;
}

// ----------------------------------------------------------------------------
// Lane-parallel variant (select.cu): every lane folds ONE row by itself, strictly left to right,
// so a warp re-ranks up to 32 candidates for the price of one sequential chain -- the fp64 pipe
// issues per warp instruction, whatever the number of active lanes.  A lane with `qlane` set
// folds the query's own squares instead (vector_norm(Query) of cosine_distance/2).
// ----------------------------------------------------------------------------
template <int DTYPE>
__device__ __forceinline__ void load8(const uint8_t *row, int t8, double mn, double sc, double (&v)[8]) {
    if (DTYPE == EVDB_F32) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(row) + 2 * t8);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(row) + 2 * t8 + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if (DTYPE == EVDB_BF16) {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(row) + t8);
        const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = (double)__uint_as_float(u[i] << 16);
            v[2 * i + 1] = (double)__uint_as_float(u[i] & 0xFFFF0000u);
        }
    } else if (DTYPE == EVDB_U8) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(row) + t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t c = ((i < 4 ? w.x : w.y) >> (8 * (i & 3))) & 0xFFu;
            v[i] = __dadd_rn(mn, __dmul_rn((double)c, sc));   // Min + (Q * Scale)
        }
    } else {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(row) + t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t byte = (w >> (8 * (i >> 1))) & 0xFFu;
            const uint32_t c = (i & 1) ? (byte & 0x0Fu) : (byte >> 4);   // first element in the high nibble
            v[i] = __dadd_rn(mn, __dmul_rn((double)c, sc));
        }
    }
}

// Returns the reference's left-to-right sum for this lane's row: sum(q*v) (cosine; sum(q*q) on the
// qlane), sum((q-v)^2) (euclidean), sum(|q-v|) (manhattan).  q: the fp64 query in global memory.
template <int DTYPE>
__device__ __forceinline__ double exact_fold_lane(const uint8_t *row, double mn, double sc,
                                                         const double *__restrict__ q, int d, int metric, bool qlane) {
    double s = 0.0;
    const int full = d >> 3;
    for (int t8 = 0; t8 < full; ++t8) {
        double v[8], qq[8];
        load8<DTYPE>(row, t8, mn, sc, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) qq[i] = __ldg(q + 8 * t8 + i);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double x = qlane ? qq[i] : v[i];
            double term;
            if (metric == EVDB_COSINE) term = __dmul_rn(qq[i], x);
            else {
                const double t0 = __dsub_rn(qq[i], x);
                term = metric == EVDB_EUCLIDEAN ? __dmul_rn(t0, t0) : fabs(t0);
            }
            s = __dadd_rn(s, term);
        }
    }
    for (int t = full << 3; t < d; ++t) {
        const double qv = q[t];
        const double x = qlane ? qv : row_elem<DTYPE>(row, t, mn, sc);
        double term;
        if (metric == EVDB_COSINE) term = __dmul_rn(qv, x);
        else {
            const double t0 = __dsub_rn(qv, x);
            term = metric == EVDB_EUCLIDEAN ? __dmul_rn(t0, t0) : fabs(t0);
        }
        s = __dadd_rn(s, term);
    }
    return s;
}

}  // namespace evdb
