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
        if (lane == 0) {
#pragma unroll 8
            for (int i = 0; i < cnt; ++i) s = __dadd_rn(s, sp[i]);
        }
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
        if (lane == 0) {
#pragma unroll 8
            for (int i = 0; i < cnt; ++i) s = __dadd_rn(s, sp[i]);
        } else if (lane == 1 && metric == EVDB_COSINE) {
#pragma unroll 8
            for (int i = 0; i < cnt; ++i) sq = __dadd_rn(sq, sp2[i]);
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

}  // namespace evdb
