// ingest.cu -- device-side row finalisation, 8/4-bit codecs, synthetic fill.
//
// The codecs restate compress_{8,4}bit_quantization / decompress_* of the
// reference (src/vector_compression.erl:166-204, find_min_max :306-309,
// pack_4bit_values :311-319) in fp64 on the device: IEEE division and
// round-half-away-from-zero make the codes bit-identical to erlang:round/1.
#include "exact.cuh"
#include "internal.h"

namespace evdb {

// ----------------------------------------------------------------------------
// finalize: per-row cached scalars (+ bf16 shadow), one warp per row
// ----------------------------------------------------------------------------
template <int DTYPE>
__global__ void __launch_bounds__(256) finalize_rows_kernel(uint8_t *__restrict__ rows,
                                                            size_t row_bytes, int d, int spitch,
                                                            uint64_t slot0, uint64_t n,
                                                            double *__restrict__ norm64,
                                                            float *__restrict__ inv_norm,
                                                            float *__restrict__ norm_sq,
                                                            float2 *__restrict__ qcoef,
                                                            const double2 *__restrict__ qms64,
                                                            __half *__restrict__ shadow,
                                                            const double *__restrict__ src64,
                                                            const float *__restrict__ src32, int dpad) {
    __shared__ double sp_all[8 * kExactChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sp = sp_all + warp * kExactChunk;
    for (uint64_t i = (uint64_t)blockIdx.x * 8 + warp; i < n; i += (uint64_t)gridDim.x * 8) {
        const uint64_t r = slot0 + i;
        const uint8_t *row = rows + r * row_bytes;
        if ((DTYPE == EVDB_F32 || DTYPE == EVDB_BF16) && (src64 || src32)) {
            // fused ingest of float stores: the staged source row is narrowed into place by this warp first
            // (one launch per upsert instead of two), then read back below for the cached values
            for (int c = lane; c < dpad; c += 32) {
                const float v = c < d ? (src64 ? (float)src64[i * (uint64_t)d + c] : src32[i * (uint64_t)d + c]) : 0.0f;
                if (DTYPE == EVDB_F32) reinterpret_cast<float *>(rows + r * row_bytes)[c] = v;
                else reinterpret_cast<__nv_bfloat16 *>(rows + r * row_bytes)[c] = __float2bfloat16_rn(v);
            }
            __syncwarp();
        }
        double mn = 0.0, sc = 0.0;
        if (DTYPE == EVDB_U8 || DTYPE == EVDB_U4) {
            double2 ms = qms64[r];
            mn = ms.x;
            sc = ms.y;
        }
        double nrm = exact_norm_warp<DTYPE>(row, mn, sc, d, sp, lane);
        if (lane == 0) {
            norm64[r] = nrm;
            float inv = nrm > 0.0 ? (float)(1.0 / nrm) : 0.0f;
            inv_norm[r] = inv;
            norm_sq[r] = (float)(nrm * nrm);
            if (DTYPE == EVDB_U8 || DTYPE == EVDB_U4)
                qcoef[r] = nrm > 0.0 ? make_float2((float)(sc / nrm), (float)(mn / nrm))
                                     : make_float2(0.f, 0.f);
        }
        if (DTYPE == EVDB_F32 && shadow) {
            // tcgen05 operand: the unit-norm row in fp16 (|x| <= 1: no overflow for any input scale)
            const float *fr = reinterpret_cast<const float *>(row);
            const float inv = nrm > 0.0 ? (float)(1.0 / nrm) : 0.0f;
            __half *sh = shadow + r * (size_t)spitch;
            for (int c = lane; c < spitch; c += 32) sh[c] = __float2half_rn(c < d ? fr[c] * inv : 0.0f);
        }
    }
}

int launch_finalize_rows(evdb_store *s, uint64_t slot0, uint64_t n, cudaStream_t st, const void *d_src, bool src_f64) {
    if (n == 0) return EVDB_OK;
    const double *src64 = d_src && src_f64 ? (const double *)d_src : nullptr;
    const float *src32 = d_src && !src_f64 ? (const float *)d_src : nullptr;
    uint64_t blocks = (n + 7) / 8;
    uint64_t cap = (uint64_t)s->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
#define EVDB_FIN(DT)                                                                              \
    finalize_rows_kernel<DT><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->dim, s->spitch, slot0, \
                                                   n, s->norm64, s->inv_norm, s->norm_sq,         \
                                                   s->qcoef, s->qms64, s->shadow, src64, src32, s->dpad)
    switch (s->dtype) {
        case EVDB_F32: EVDB_FIN(EVDB_F32); break;
        case EVDB_BF16: EVDB_FIN(EVDB_BF16); break;
        case EVDB_U8: EVDB_FIN(EVDB_U8); break;
        default: EVDB_FIN(EVDB_U4); break;
    }
#undef EVDB_FIN
    s->n_launches++;
    EVDB_CUDA(cudaGetLastError());
    if (s->shadow && slot0 <= s->shadow_valid && slot0 + n > s->shadow_valid) s->shadow_valid = slot0 + n;
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// quantizers: one warp per row.  Source = fp64 rows, fp32 rows or the synthetic
// generator (seed != 0 path: value(row0 + r, c)).
// ----------------------------------------------------------------------------
struct RowSrc {
    const double *r64;
    const float *r32;
    uint64_t seed, row0, rstride;
    int synth;
    int d;
    __device__ __forceinline__ double at(uint64_t r, int c) const {
        if (synth) return (double)synth_value(seed, (row0 + r * rstride) * (uint64_t)d + (uint64_t)c);
        if (r64) return r64[r * (uint64_t)d + c];
        return (double)r32[r * (uint64_t)d + c];
    }
};

template <int DTYPE>
__global__ void __launch_bounds__(256) quantize_rows_kernel(const RowSrc src, uint64_t n, int d,
                                                            uint8_t *__restrict__ codes,
                                                            size_t code_row_bytes,
                                                            double2 *__restrict__ ms64,
                                                            double *__restrict__ maxs,
                                                            uint8_t *__restrict__ ok) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double levels = (DTYPE == EVDB_U8) ? 255.0 : 15.0;
    for (uint64_t r = (uint64_t)blockIdx.x * 8 + warp; r < n; r += (uint64_t)gridDim.x * 8) {
        // find_min_max/1
        double lo = src.at(r, 0), hi = lo;
        for (int c = lane; c < d; c += 32) {
            double v = src.at(r, c);
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        for (int o = 16; o > 0; o >>= 1) {
            double l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
            lo = l2 < lo ? l2 : lo;
            hi = h2 > hi ? h2 : hi;
        }
        const double scale = __ddiv_rn(__dsub_rn(hi, lo), levels);
        const bool good = scale != 0.0;
        uint8_t *out = codes + r * code_row_bytes;
        if (DTYPE == EVDB_U8) {
            for (int c = lane; c < (int)code_row_bytes; c += 32) {
                uint8_t q = 0;
                if (good && c < d) q = (uint8_t)round(__ddiv_rn(__dsub_rn(src.at(r, c), lo), scale));
                out[c] = q;
            }
        } else {
            for (int j = lane; j < (int)code_row_bytes; j += 32) {
                uint32_t q0 = 0, q1 = 0;
                if (good && 2 * j < d) q0 = (uint32_t)round(__ddiv_rn(__dsub_rn(src.at(r, 2 * j), lo), scale));
                if (good && 2 * j + 1 < d) q1 = (uint32_t)round(__ddiv_rn(__dsub_rn(src.at(r, 2 * j + 1), lo), scale));
                out[j] = (uint8_t)((q0 << 4) | q1);  // <<V1:4, V2:4>>
            }
        }
        if (lane == 0) {
            // Max == Min: the reference raises badarith and keeps the raw (constant) vector;
            // {min, scale = 0} with all-zero codes decodes to exactly that vector.
            ms64[r] = make_double2(lo, good ? scale : 0.0);
            if (maxs) maxs[r] = hi;
            if (ok) ok[r] = good ? 1 : 0;
        }
    }
}

int launch_quantize_rows(int dtype, const double *d_rows64, const float *d_rows32, uint64_t n, int d,
                         uint8_t *codes, size_t code_row_bytes, double2 *ms64, double *maxs,
                         uint8_t *ok, cudaStream_t st) {
    if (n == 0) return EVDB_OK;
    RowSrc src{d_rows64, d_rows32, 0, 0, 1, 0, d};
    uint64_t blocks = (n + 7) / 8;
    int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
    if (dtype == EVDB_U8) quantize_rows_kernel<EVDB_U8><<<grid, 256, 0, st>>>(src, n, d, codes, code_row_bytes, ms64, maxs, ok);
    else quantize_rows_kernel<EVDB_U4><<<grid, 256, 0, st>>>(src, n, d, codes, code_row_bytes, ms64, maxs, ok);
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

__global__ void dequantize_rows_kernel(int dtype, const uint8_t *__restrict__ codes,
                                       size_t code_row_bytes, const double2 *__restrict__ ms64,
                                       uint64_t n, int d, double *__restrict__ out) {
    uint64_t total = n * (uint64_t)d;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / d;
        int c = (int)(i % d);
        double2 ms = ms64[r];
        const uint8_t *row = codes + r * code_row_bytes;
        out[i] = dtype == EVDB_U8 ? row_elem<EVDB_U8>(row, c, ms.x, ms.y)
                                  : row_elem<EVDB_U4>(row, c, ms.x, ms.y);
    }
}

int launch_dequantize_rows(int dtype, const uint8_t *codes, size_t code_row_bytes,
                           const double2 *ms64, uint64_t n, int d, double *out, cudaStream_t st) {
    if (n == 0) return EVDB_OK;
    uint64_t total = n * (uint64_t)d;
    uint64_t blocks = (total + 255) / 256;
    int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    dequantize_rows_kernel<<<grid, 256, 0, st>>>(dtype, codes, code_row_bytes, ms64, n, d, out);
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// synthetic corpus (bench/tests): generated in place, never crosses PCIe
// ----------------------------------------------------------------------------
template <int DTYPE>
__global__ void __launch_bounds__(256) fill_float_kernel(uint8_t *__restrict__ rows, size_t row_bytes,
                                                         int d, int dpad, uint64_t seed,
                                                         uint64_t row0, uint64_t rstride, uint64_t n) {
    const uint64_t total = n * (uint64_t)dpad;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / dpad;
        int c = (int)(i % dpad);
        float v = c < d ? synth_value(seed, (row0 + r * rstride) * (uint64_t)d + (uint64_t)c) : 0.0f;
        if (DTYPE == EVDB_F32) reinterpret_cast<float *>(rows + r * row_bytes)[c] = v;
        else reinterpret_cast<__nv_bfloat16 *>(rows + r * row_bytes)[c] = __float2bfloat16_rn(v);
    }
}

int launch_fill_synthetic(evdb_store *s, uint64_t seed, uint64_t row0, uint64_t rstride, uint64_t n, cudaStream_t st) {
    if (n == 0) return EVDB_OK;
    if (s->dtype == EVDB_F32 || s->dtype == EVDB_BF16) {
        uint64_t blocks = (n * (uint64_t)s->dpad + 255) / 256;
        int grid = (int)(blocks < (uint64_t)s->sm_count * 16 ? blocks : (uint64_t)s->sm_count * 16);
        if (s->dtype == EVDB_F32) fill_float_kernel<EVDB_F32><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->dim, s->dpad, seed, row0, rstride, n);
        else fill_float_kernel<EVDB_BF16><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->dim, s->dpad, seed, row0, rstride, n);
    } else {
        RowSrc src{nullptr, nullptr, seed, row0, rstride, 1, s->dim};
        uint64_t blocks = (n + 7) / 8;
        int grid = (int)(blocks < (uint64_t)s->sm_count * 8 ? blocks : (uint64_t)s->sm_count * 8);
        if (s->dtype == EVDB_U8) quantize_rows_kernel<EVDB_U8><<<grid, 256, 0, st>>>(src, n, s->dim, s->rows, s->row_bytes, s->qms64, nullptr, nullptr);
        else quantize_rows_kernel<EVDB_U4><<<grid, 256, 0, st>>>(src, n, s->dim, s->rows, s->row_bytes, s->qms64, nullptr, nullptr);
    }
    s->n_launches++;
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// vector_utils (reference src/vector_utils.erl:28-57): pairwise functions of two fp64 vectors, one
// warp per pair, in the reference's operation order (independent products, strictly sequential
// lists:sum folds, no FMA).  Lane 0 folds the cross terms, lanes 1 / 2 the squares of a / b.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vector_utils_kernel(int op, const double *__restrict__ a, const double *__restrict__ b,
                                                           uint64_t n, int d, double *__restrict__ out) {
    __shared__ double sp_all[8 * 3 * kExactChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sp = sp_all + warp * 3 * kExactChunk;
    for (uint64_t p = (uint64_t)blockIdx.x * 8 + warp; p < n; p += (uint64_t)gridDim.x * 8) {
        const double *x = a + p * (uint64_t)d, *y = b ? b + p * (uint64_t)d : x;
        double s = 0.0;   // lane 0: cross terms; lane 1: sum x*x; lane 2: sum y*y
        for (int base = 0; base < d; base += kExactChunk) {
            const int cnt = min(kExactChunk, d - base);
            for (int t = lane; t < cnt; t += kWarp) {
                const double u = x[base + t], v = y[base + t];
                double c;
                if (op == EVDB_VU_EUCLIDEAN) { const double df = __dsub_rn(u, v); c = __dmul_rn(df, df); }   // vector_subtract, then X*X
                else if (op == EVDB_VU_MANHATTAN) c = fabs(__dsub_rn(u, v));
                else c = __dmul_rn(u, v);
                sp[t] = c;
                sp[kExactChunk + t] = __dmul_rn(u, u);
                sp[2 * kExactChunk + t] = __dmul_rn(v, v);
            }
            __syncwarp();
            if (lane < 3) s = fold_staged(s, sp + lane * kExactChunk, cnt);
            __syncwarp();
        }
        const double cross = __shfl_sync(0xffffffffu, s, 0), xx = __shfl_sync(0xffffffffu, s, 1), yy = __shfl_sync(0xffffffffu, s, 2);
        if (lane == 0) {
            double r;
            const double n1 = __dsqrt_rn(xx), n2 = __dsqrt_rn(yy);
            switch (op) {
                case EVDB_VU_COSINE_SIMILARITY: r = (n1 == 0.0 || n2 == 0.0) ? 0.0 : __ddiv_rn(cross, __dmul_rn(n1, n2)); break;
                case EVDB_VU_COSINE_DISTANCE: r = (n1 == 0.0 || n2 == 0.0) ? 1.0 : __dsub_rn(1.0, __ddiv_rn(cross, __dmul_rn(n1, n2))); break;
                case EVDB_VU_EUCLIDEAN: r = __dsqrt_rn(cross); break;
                case EVDB_VU_MANHATTAN: r = cross; break;
                case EVDB_VU_DOT: r = cross; break;
                default: r = n1; break;   // EVDB_VU_NORM
            }
            out[p] = r;
        }
    }
}

int launch_vector_utils(int op, const double *d_a, const double *d_b, uint64_t n, int d, double *d_out, cudaStream_t st) {
    if (n == 0) return EVDB_OK;
    const uint64_t blocks = (n + 7) / 8;
    vector_utils_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(op, d_a, d_b, n, d, d_out);
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

}  // namespace evdb
