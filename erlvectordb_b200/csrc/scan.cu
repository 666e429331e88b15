// scan.cu -- HBM-bound distance scans with fused partial top-k (family A).
//
// Replaces the maps:fold over all N entries in perform_search/3
// (reference src/vector_store.erl:227-231) and the per-row cosine_distance /
// dot_product / vector_norm passes (:238-252) for small query batches, the
// manhattan/euclidean forms of src/vector_utils.erl:38-43, and the scan over
// quantization_8bit / quantization_4bit codes (src/vector_compression.erl:166-204).
//
// Layout: rows are dense, row-major, padded to 16-byte chunks.  A group of TPR
// lanes owns a row; each lane streams 128-bit chunks of R rows at a time
// (ld.global.nc.L1::no_allocate), the query sits in shared memory, partial
// sums are combined with warp shuffles, and each warp keeps its KP best
// (score, slot) keys in a sorted shared-memory list guarded by a register
// threshold.  The CTA bitonic-merges its warps' lists and writes KP keys; the
// select kernel (select.cu) merges CTAs and re-ranks exactly in fp64.
#include <stdio.h>
#include <stdlib.h>

#include "internal.h"
#include "topk.cuh"
#include "scan_common.cuh"

namespace evdb {

// ----------------------------------------------------------------------------
// query preparation
// ----------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T block_reduce(T v, T *red, bool is_max) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? (u > v ? u : v) : (v + u);
    }
    if (lane == 0) red[warp] = v;
    __syncthreads();
    T r = red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? (red[i] > r ? red[i] : r) : (r + red[i]);
    __syncthreads();
    return r;
}

// One CTA per query: narrow to fp32 (zero padded), norms, and -- for quantized
// stores -- the fixed-point query (8*kQPlanes bits) split into 8-bit digit planes so
// that sum(Q_i * c_i) is an exact integer computed with dp4a.
__global__ void __launch_bounds__(256) prep_queries_kernel(const double *__restrict__ q64, int d,
                                                           float *__restrict__ q32, int q32_stride,
                                                           uint8_t *__restrict__ qdig,
                                                           int qdig_stride, int dtype, int metric,
                                                           QStat *__restrict__ qstat, float *__restrict__ qeps) {
    __shared__ double redd[8];
    __shared__ long long redl[8];
    pdl_launch_dependents();   // first link of a scan plan's chain (common.cuh)
    const int b = blockIdx.x;
    const double *q = q64 + (size_t)b * d;
    double ss = 0.0, sm = 0.0, mx = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        double v = q[i];
        ss += v * v;
        sm += v;
        mx = fmax(mx, fabs(v));
    }
    ss = block_reduce<double>(ss, redd, false);
    sm = block_reduce<double>(sm, redd, false);
    if (q32) {
        float *o = q32 + (size_t)b * q32_stride;
        // The float scans see q narrowed to fp32.  For euclidean / manhattan that moves every score by
        // up to the narrowing residual (triangle inequality: | ||q32-v|| - ||q-v|| | <= ||q-q32||_2,
        // | |q32-v|_1 - |q-v|_1 | <= |q-q32|_1) -- an ABSOLUTE error the relative arithmetic bound does
        // not cover when rows lie much closer to the query than ||q||.  It becomes the query's eps_q
        // (0 for a query that is exact in fp32; cosine's absolute bound already covers 2^-24 ||q||).
        double r2 = 0.0, r1 = 0.0;
        for (int i = threadIdx.x; i < q32_stride; i += blockDim.x) {
            const float f = i < d ? (float)q[i] : 0.0f;
            o[i] = f;
            if (i < d) {
                const double df = fabs(q[i] - (double)f);
                r2 += df * df;
                r1 += df;
            }
        }
        if (qeps) {
            r2 = block_reduce<double>(r2, redd, false);
            r1 = block_reduce<double>(r1, redd, false);
            if (threadIdx.x == 0) {
                float e = 0.f;
                if (metric == EVDB_EUCLIDEAN) e = (float)(sqrt(r2) * 1.0000002);
                else if (metric == EVDB_MANHATTAN) e = (float)(r1 * 1.0000002);
                if (e > 0.f) e = nextafterf(e, __int_as_float(0x7f800000));   // the casts round to nearest: never below the bound
                qeps[b] = e;
            }
        }
    }
    QStat st;
    st.norm_sq = (float)ss;
    st.inv_norm = ss > 0.0 ? (float)(1.0 / sqrt(ss)) : 0.0f;
    st.sum = (float)sm;
    st.fx = 0.0f;
    if (dtype == EVDB_U8 || dtype == EVDB_U4) {
        mx = block_reduce<double>(mx, redd, true);
        constexpr int kBits = 8 * kQPlanes;
        int e = 0;
        if (mx > 0.0) {
            int x;
            frexp(mx, &x);  // mx = m * 2^x, m in [0.5,1)  =>  |q| * 2^(kBits-1-x) < 2^(kBits-1)
            e = kBits - 1 - x;
        }
        double sc = ldexp(1.0, e);
        const double qmax = (double)((1 << (kBits - 1)) - 1), qmin = -(double)(1 << (kBits - 1));
        uint8_t *p0 = qdig + (size_t)b * kQPlanes * qdig_stride;
        long long qsum = 0;
        for (int i = threadIdx.x; i < qdig_stride; i += blockDim.x) {
            int Q = 0;
            if (i < d) {
                double t = rint(q[i] * sc);
                t = fmin(fmax(t, qmin), qmax);
                Q = (int)t;
            }
            qsum += Q;
            int pos = i;
            if (dtype == EVDB_U4) {
                // codes: byte j = (elem 2j << 4) | elem 2j+1.  (w >> 4) & 0x0F0F0F0F yields the
                // even elements of a word, w & 0x0F0F0F0F the odd ones: lay the digits out to match,
                // even-element digits of all chunks first, then the odd ones (both halves are read
                // at a 16-byte stride across lanes: no shared-memory bank conflicts).
                int c = i >> 5, r = i & 31, w = r >> 3, e8 = r & 7;
                pos = ((e8 & 1) ? (qdig_stride >> 1) : 0) + c * 16 + w * 4 + (e8 >> 1);
            }
#pragma unroll
            for (int pl = 0; pl < kQPlanes; ++pl)   // plane 0 = the signed high digit
                p0[(size_t)pl * qdig_stride + pos] = (uint8_t)((Q >> (8 * (kQPlanes - 1 - pl))) & 0xFF);
        }
        // |q_i - Q_i 2^-e| <= 2^-e (half a unit from rounding, up to one where the clamp bites) and
        // max|q| 2^e >= 2^(kBits-2):  |dq . y| <= max|q| 2^-(kBits-2) ||y||_1 <= ... sqrt(d) ||y||
        if (threadIdx.x == 0 && qeps)
            qeps[b] = ss > 0.0 ? (float)(1.01 * ldexp(1.0, -(kBits - 2)) * sqrt((double)d) * mx / sqrt(ss)) : 0.f;
        qsum = block_reduce<long long>(qsum, redl, false);
        st.fx = (float)ldexp(1.0, -e);
        st.sum = (float)((double)qsum * ldexp(1.0, -e));
    }
    if (threadIdx.x == 0) qstat[b] = st;
}

int launch_prep_queries(evdb_store *s, const double *d_q64, int B, int metric, cudaStream_t st) {
    bool quant = s->dtype == EVDB_U8 || s->dtype == EVDB_U4;
    int q32_stride = quant ? 0 : s->dpad;
    int qdig_stride = quant ? s->dpad : 0;
    if (!quant) EVDB_TRY(ensure_bytes((void **)&s->w_q32, &s->w_q32_cap, sizeof(float) * (size_t)B * q32_stride));
    else EVDB_TRY(ensure_bytes((void **)&s->w_qdig, &s->w_qdig_cap, (size_t)B * kQPlanes * qdig_stride));
    EVDB_TRY(ensure_bytes((void **)&s->w_qeps, &s->w_qeps_cap, sizeof(float) * (size_t)B));
    EVDB_TRY(ensure_bytes((void **)&s->w_qstat, &s->w_qstat_cap, sizeof(QStat) * (size_t)B));
    prep_queries_kernel<<<B, 256, 0, st>>>(d_q64, s->dim, quant ? nullptr : s->w_q32, q32_stride,
                                           s->w_qdig, qdig_stride, s->dtype, metric, s->w_qstat, s->w_qeps);
    s->n_launches++;
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// fp32 / bf16 rows: cosine, euclidean, manhattan
// ----------------------------------------------------------------------------
template <int METRIC, int DTYPE, int TPR, int R>
__global__ void __launch_bounds__(kScanWarps * 32, DTYPE == EVDB_F32 ? 4 : 3)   // F32: 4 CTAs (32 warps) per SM -- the loads in flight are what hides HBM latency
scan_float_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();                // the prepared query comes from prep_queries_kernel
    pdl_launch_dependents();
    constexpr int QPC = (DTYPE == EVDB_F32) ? 1 : 2;  // query float4s per 16-byte row chunk
    constexpr int GPW = 32 / TPR;                      // row groups per warp
    const int nch = a.nch, KP = a.KP;
    float4 *sq = reinterpret_cast<float4 *>(smem);
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + (size_t)nch * QPC * 16);
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const float4 *qsrc = reinterpret_cast<const float4 *>(a.q32 + (size_t)b * a.q32_stride);
    for (int i = threadIdx.x; i < nch * QPC; i += blockDim.x) sq[i] = qsrc[i];
    if (KP > kAppendMaxKP)
        for (int i = threadIdx.x; i < kScanWarps * KP; i += blockDim.x) lists[i] = kKeyMax;
    __syncthreads();
    const float q_inv = a.qstat[b].inv_norm;
    WarpCands wc;
    wc.init(lists, KP, warp);

    const int g = lane / TPR, gl = lane % TPR;
    const uint64_t rows_per_wi = (uint64_t)GPW * R;
    const uint64_t total_wi = (a.n + rows_per_wi - 1) / rows_per_wi;
    for (uint64_t wi = (uint64_t)blockIdx.x * kScanWarps + warp; wi < total_wi;
         wi += (uint64_t)gridDim.x * kScanWarps) {
        const uint64_t base = wi * rows_per_wi;
        const uint4 *rp[R];
        uint64_t rix[R];
        bool valid[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            uint64_t r = base + (uint64_t)j * GPW + g;
            valid[j] = r < a.n;
            rix[j] = r;
            rp[j] = reinterpret_cast<const uint4 *>(a.rows + (valid[j] ? r : a.n - 1) * a.row_bytes);
        }
        // the per-row side value is requested up front, with the rows: issued after the
        // reduction it would put a second DRAM latency on every warp-iteration
        float inv[R];
#pragma unroll
        for (int j = 0; j < R; ++j)
            inv[j] = (METRIC == EVDB_COSINE && valid[j] && gl == 0) ? __ldg(a.inv_norm + rix[j]) : 0.f;
        float4 acc[R];
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int c = gl; c < nch; c += TPR) {
            uint4 v[R];
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = ldg_stream_u4(rp[j] + c);
            if (DTYPE == EVDB_F32) {
                const float4 q0 = sq[c];
#pragma unroll
                for (int j = 0; j < R; ++j)
                    acc_f4<METRIC>(acc[j],
                                   make_float4(__uint_as_float(v[j].x), __uint_as_float(v[j].y),
                                               __uint_as_float(v[j].z), __uint_as_float(v[j].w)),
                                   q0);
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
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            float score;
            if (METRIC == EVDB_COSINE) {
                score = (inv[j] == 0.f || q_inv == 0.f) ? 1.0f : 1.0f - sacc * inv[j] * q_inv;
            } else if (METRIC == EVDB_EUCLIDEAN) {
                score = sqrtf(sacc);
            } else {
                score = sacc;
            }
            uint64_t key = (valid[j] && gl == 0) ? make_key(score, (uint32_t)rix[j]) : kKeyMax;
            wc.offer(key, lane);
        }
    }
    wc.finish(lists, warp, lane);
    cta_merge_and_store(lists, KP, a.partial + ((size_t)b * a.G + blockIdx.x) * KP);
}

// one 16-byte chunk (index c) of R rows against the three digit planes of the query
template <int DTYPE, int R>
__device__ __forceinline__ void quant_chunk(const uint4 (&v)[R], int c, const uint4 *sd, int plane_u4,
                                            int (&A)[R], int (&Bm)[R], int (&Cl)[R]) {
    if (DTYPE == EVDB_U8) {
        const uint4 d0 = sd[c], d1 = sd[plane_u4 + c];
        const uint4 d2 = kQPlanes == 3 ? sd[2 * plane_u4 + c] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            A[j] = dp4a_su((int)d0.x, v[j].x, A[j]); A[j] = dp4a_su((int)d0.y, v[j].y, A[j]);
            A[j] = dp4a_su((int)d0.z, v[j].z, A[j]); A[j] = dp4a_su((int)d0.w, v[j].w, A[j]);
            Bm[j] = (int)dp4a_uu(d1.x, v[j].x, (uint32_t)Bm[j]); Bm[j] = (int)dp4a_uu(d1.y, v[j].y, (uint32_t)Bm[j]);
            Bm[j] = (int)dp4a_uu(d1.z, v[j].z, (uint32_t)Bm[j]); Bm[j] = (int)dp4a_uu(d1.w, v[j].w, (uint32_t)Bm[j]);
            if (kQPlanes == 3) {
                Cl[j] = (int)dp4a_uu(d2.x, v[j].x, (uint32_t)Cl[j]); Cl[j] = (int)dp4a_uu(d2.y, v[j].y, (uint32_t)Cl[j]);
                Cl[j] = (int)dp4a_uu(d2.z, v[j].z, (uint32_t)Cl[j]); Cl[j] = (int)dp4a_uu(d2.w, v[j].w, (uint32_t)Cl[j]);
            }
        }
    } else {
        const int half = plane_u4 >> 1;   // = chunks per row
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        const uint4 e0 = sd[c], o0 = sd[half + c];
        const uint4 e1 = sd[plane_u4 + c], o1 = sd[plane_u4 + half + c];
        const uint4 e2 = kQPlanes == 3 ? sd[2 * plane_u4 + c] : z4, o2 = kQPlanes == 3 ? sd[2 * plane_u4 + half + c] : z4;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            const uint32_t E0[4] = {e0.x, e0.y, e0.z, e0.w}, O0[4] = {o0.x, o0.y, o0.z, o0.w};
            const uint32_t E1[4] = {e1.x, e1.y, e1.z, e1.w}, O1[4] = {o1.x, o1.y, o1.z, o1.w};
            const uint32_t E2[4] = {e2.x, e2.y, e2.z, e2.w}, O2[4] = {o2.x, o2.y, o2.z, o2.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint32_t hi = (w[t] >> 4) & 0x0F0F0F0Fu, lo = w[t] & 0x0F0F0F0Fu;
                A[j] = dp4a_su((int)E0[t], hi, A[j]); A[j] = dp4a_su((int)O0[t], lo, A[j]);
                Bm[j] = (int)dp4a_uu(E1[t], hi, (uint32_t)Bm[j]); Bm[j] = (int)dp4a_uu(O1[t], lo, (uint32_t)Bm[j]);
                if (kQPlanes == 3) {
                    Cl[j] = (int)dp4a_uu(E2[t], hi, (uint32_t)Cl[j]); Cl[j] = (int)dp4a_uu(O2[t], lo, (uint32_t)Cl[j]);
                }
            }
        }
    }
}

// reduce the three digit sums over the TPR lanes of a row and form the key (lane gl == 0 holds it)
template <int TPR>
__device__ __forceinline__ uint64_t quant_key(int sa, int sb, int sc, bool owner, float2 co, const QStat &qs,
                                              uint64_t row) {
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        if (kQPlanes == 3) sc += __shfl_xor_sync(0xffffffffu, sc, o);
    }
    uint64_t key = kKeyMax;
    if (owner) {
        long long S = kQPlanes == 3 ? ((long long)sa << 16) + ((long long)sb << 8) + (long long)sc
                                    : ((long long)sa << 8) + (long long)sb;
        float dotn = fmaf(co.x, __ll2float_rn(S) * qs.fx, co.y * qs.sum);
        bool zero = (co.x == 0.f && co.y == 0.f) || qs.inv_norm == 0.f;
        float score = zero ? 1.0f : 1.0f - dotn * qs.inv_norm;
        key = make_key(score, (uint32_t)row);
    }
    return key;
}

// Diagnostics (evdb_debug_quant_dots): the integer digit-plane sums of chosen rows, formed by the SAME
// quant_chunk arithmetic the scans run, so a test can compare them with integer arithmetic on the host.
template <int DTYPE>
__global__ void __launch_bounds__(256) quant_dots_debug_kernel(const uint8_t *__restrict__ rows, size_t row_bytes, int nch,
                                                               const uint8_t *__restrict__ qdig, int qdig_stride,
                                                               const uint32_t *__restrict__ slots, int n,
                                                               long long *__restrict__ out_S, int *__restrict__ out_planes,
                                                               int *__restrict__ out_csum) {
    constexpr int UPC = (DTYPE == EVDB_U8) ? 1 : 2;
    const int lane = threadIdx.x & 31, w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= n) return;
    const uint4 *sd = reinterpret_cast<const uint4 *>(qdig);
    const int plane_u4 = nch * UPC;
    const uint4 *rp = reinterpret_cast<const uint4 *>(rows + (size_t)slots[w] * row_bytes);
    int A[1] = {0}, Bm[1] = {0}, Cl[1] = {0};
    uint32_t cs = 0;
    for (int c = lane; c < nch; c += 32) {
        uint4 v[1] = {rp[c]};
        quant_chunk<DTYPE, 1>(v, c, sd, plane_u4, A, Bm, Cl);
        const uint32_t ww[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (DTYPE == EVDB_U8) cs = dp4a_uu(0x01010101u, ww[t], cs);
            else { cs = dp4a_uu(0x01010101u, (ww[t] >> 4) & 0x0F0F0F0Fu, cs); cs = dp4a_uu(0x01010101u, ww[t] & 0x0F0F0F0Fu, cs); }
        }
    }
    int sa = A[0], sb = Bm[0], sc = Cl[0];
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        sc += __shfl_xor_sync(0xffffffffu, sc, o);
        cs += __shfl_xor_sync(0xffffffffu, cs, o);
    }
    if (lane == 0) {
        out_S[w] = kQPlanes == 3 ? ((long long)sa << 16) + ((long long)sb << 8) + (long long)sc : ((long long)sa << 8) + (long long)sb;
        out_planes[3 * w] = sa; out_planes[3 * w + 1] = sb; out_planes[3 * w + 2] = sc;
        out_csum[w] = (int)cs;
    }
}

int debug_quant_dots(evdb_store *s, const double *d_q64, const uint32_t *d_slots, int n, long long *d_S, int *d_planes,
                     int *d_csum, float *h_fx, cudaStream_t st) {
    EVDB_TRY(launch_prep_queries(s, d_q64, 1, EVDB_COSINE, st));
    const int grid = (n + 7) / 8;
    if (s->dtype == EVDB_U8)
        quant_dots_debug_kernel<EVDB_U8><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->nch, s->w_qdig, s->dpad, d_slots, n, d_S, d_planes, d_csum);
    else
        quant_dots_debug_kernel<EVDB_U4><<<grid, 256, 0, st>>>(s->rows, s->row_bytes, s->nch, s->w_qdig, s->dpad, d_slots, n, d_S, d_planes, d_csum);
    EVDB_CUDA(cudaGetLastError());
    QStat qs;
    EVDB_CUDA(cudaMemcpyAsync(&qs, s->w_qstat, sizeof(qs), cudaMemcpyDeviceToHost, st));
    EVDB_CUDA(cudaStreamSynchronize(st));
    *h_fx = qs.fx;
    return EVDB_OK;
}

// ----------------------------------------------------------------------------
// u8 / packed-u4 codes: cosine of an unquantised query against Min + c*Scale
//   q.y = Min*sum(q) + Scale*sum(q_i c_i);  sum(Q_i c_i) is an exact integer:
//   Q = a*2^8 + b (a signed, b unsigned digit; a third digit with EVDB_QPLANES=3), one dp4a per digit and word.
// ----------------------------------------------------------------------------
template <int DTYPE, int TPR, int R>
__global__ void __launch_bounds__(kScanWarps * 32, 2)
scan_quant_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();                // the prepared query comes from prep_queries_kernel
    pdl_launch_dependents();
    constexpr int GPW = 32 / TPR;
    constexpr int UPC = (DTYPE == EVDB_U8) ? 1 : 2;  // uint4 of digits per plane per row chunk
    const int nch = a.nch, KP = a.KP;
    const int plane_u4 = nch * UPC;                  // uint4 per plane
    uint4 *sd = reinterpret_cast<uint4 *>(smem);     // [kQPlanes][plane_u4]
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + (size_t)kQPlanes * plane_u4 * 16);
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const uint4 *dsrc = reinterpret_cast<const uint4 *>(a.qdig + (size_t)b * kQPlanes * a.qdig_stride);
    for (int i = threadIdx.x; i < kQPlanes * plane_u4; i += blockDim.x) sd[i] = dsrc[i];
    if (KP > kAppendMaxKP)
        for (int i = threadIdx.x; i < kScanWarps * KP; i += blockDim.x) lists[i] = kKeyMax;
    __syncthreads();
    const QStat qs = a.qstat[b];
    WarpCands wc;
    wc.init(lists, KP, warp);

    const int g = lane / TPR, gl = lane % TPR;
    const uint64_t rows_per_wi = (uint64_t)GPW * R;
    const uint64_t total_wi = (a.n + rows_per_wi - 1) / rows_per_wi;
    for (uint64_t wi = (uint64_t)blockIdx.x * kScanWarps + warp; wi < total_wi;
         wi += (uint64_t)gridDim.x * kScanWarps) {
        const uint64_t base = wi * rows_per_wi;
        const uint4 *rp[R];
        uint64_t rix[R];
        bool valid[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            uint64_t r = base + (uint64_t)j * GPW + g;
            valid[j] = r < a.n;
            rix[j] = r;
            rp[j] = reinterpret_cast<const uint4 *>(a.rows + (valid[j] ? r : a.n - 1) * a.row_bytes);
        }
        // per-row coefficients requested up front, together with the codes
        float2 co[R];
#pragma unroll
        for (int j = 0; j < R; ++j)
            co[j] = (valid[j] && gl == 0) ? __ldg(a.qcoef + rix[j]) : make_float2(0.f, 0.f);  // {scale/||y||, min/||y||}
        int A[R], Bm[R], Cl[R];
#pragma unroll
        for (int j = 0; j < R; ++j) A[j] = Bm[j] = Cl[j] = 0;
#pragma unroll 2
        for (int c = gl; c < nch; c += TPR) {
            uint4 v[R];
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = ldg_stream_u4(rp[j] + c);
            quant_chunk<DTYPE, R>(v, c, sd, plane_u4, A, Bm, Cl);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            uint64_t key = quant_key<TPR>(A[j], Bm[j], Cl[j], valid[j] && gl == 0, co[j], qs, rix[j]);
            wc.offer(key, lane);
        }
    }
    wc.finish(lists, warp, lane);
    cta_merge_and_store(lists, KP, a.partial + ((size_t)b * a.G + blockIdx.x) * KP);
}

// ----------------------------------------------------------------------------
// The same scan with the codes staged through shared memory by the TMA unit.
//
// The register-fed kernel above keeps at most (warps x rows x chunks) 16-byte loads in flight and
// its dp4a work holds 128 registers, i.e. 16 warps per SM: memory latency and integer issue do
// not overlap fully (75 % of the HBM peak).  Here a producer warp streams whole tiles of rows
// (dense rows => one contiguous cp.async.bulk per tile, plus one for the tile's coefficients)
// into a ring of `stages` shared-memory buffers, full/empty mbarriers per stage; the bytes in
// flight are the ring, whatever the consumers are doing.  Eight consumer warps run the same
// dp4a arithmetic out of shared memory.  Lane -> chunk assignment is rotated per row group
// (`bank_mul`, chosen on the host) so that the 8 lanes of a quarter-warp hit 8 distinct 16-byte
// bank groups even when the row pitch is a multiple of 128 bytes.
// A tile is consumed by `tma_wt` warps (WT x 32/TPR x R rows): the 8 consumer warps form 8/WT groups
// that take the CTA's tiles in turn, so a tile stays ~25 KB whatever the row length while every
// warp still works on R = 2 rows at once (the query digits it reads from shared memory are
// amortised over them: (3 + R)/R shared-memory bytes per row byte for u8, (6 + R)/R for u4).
// The ring depth is a multiple of the group count, so a stage always belongs to one group.
// Tiles are whole; the ragged tail of the store goes through the direct-load path in one CTA.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool sbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {  // bounded: a protocol bug traps
    if (sbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!sbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kTmaMaxStages = 8;

template <int DTYPE, int TPR, int R>
__global__ void __launch_bounds__((kScanWarps + 1) * 32, 2)
scan_quant_tma_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];   // (the runtime places dynamic shared memory on a >= 128-byte boundary; stages are multiples of 128 B)
    pdl_wait();                // the prepared query comes from prep_queries_kernel
    pdl_launch_dependents();
    constexpr int GPW = 32 / TPR;
    constexpr int UPC = (DTYPE == EVDB_U8) ? 1 : 2;
    const int WT = a.tma_wt, NG = kScanWarps / WT;     // warps per tile, consumer groups
    const int TR = WT * GPW * R;                       // rows per tile
    const int nch = a.nch, KP = a.KP, S = a.tma_stages;
    const int plane_u4 = nch * UPC;
    const uint32_t code_bytes = (uint32_t)TR * (uint32_t)a.row_bytes, coef_bytes = TR * 8u;
    uint8_t *ring = smem;                              // [S][tma_stage_bytes]
    uint4 *sd = reinterpret_cast<uint4 *>(smem + (size_t)S * a.tma_stage_bytes);
    uint64_t *lists = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(sd) + (size_t)kQPlanes * plane_u4 * 16);
    __shared__ __align__(8) uint64_t bars[2 * kTmaMaxStages];   // full[S], empty[S]
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool consumer = warp < kScanWarps;

    const uint4 *dsrc = reinterpret_cast<const uint4 *>(a.qdig + (size_t)b * kQPlanes * a.qdig_stride);
    for (int i = threadIdx.x; i < kQPlanes * plane_u4; i += blockDim.x) sd[i] = dsrc[i];
    if (KP > kAppendMaxKP)
        for (int i = threadIdx.x; i < kScanWarps * KP; i += blockDim.x) lists[i] = kKeyMax;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            sbar_init(s_u32(&bars[i]), 1);
            sbar_init(s_u32(&bars[kTmaMaxStages + i]), WT);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const QStat qs = a.qstat[b];
    WarpCands wc;
    wc.init(lists, KP, consumer ? warp : 0);

    const uint64_t n_tiles = a.n / TR;
    if (!consumer) {
        if (lane == 0) {
            uint32_t i = 0;
            for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
                const uint32_t st = i % S, use = i / S;
                if (use > 0) sbar_wait(s_u32(&bars[kTmaMaxStages + st]), (use - 1) & 1);
                const uint32_t full = s_u32(&bars[st]);
                const uint32_t dst = s_u32(ring + (size_t)st * a.tma_stage_bytes);
                sbar_expect_tx(full, code_bytes + coef_bytes);
                bulk_g2s(dst, a.rows + t * (uint64_t)code_bytes, code_bytes, full);
                bulk_g2s(dst + code_bytes, a.qcoef + t * TR, coef_bytes, full);
            }
        }
        __syncwarp();
    } else {
        const int g = lane / TPR, gl = lane % TPR;
        int shift = (g * a.bank_mul) & 7;
        if (shift >= nch) shift %= nch;
        const int row_in_tile0 = (warp % WT) * GPW * R + g;
        uint32_t i = warp / WT;                        // this group's first tile of the CTA's sequence
        for (uint64_t t = blockIdx.x + (uint64_t)i * gridDim.x; t < n_tiles; t += (uint64_t)NG * gridDim.x, i += NG) {
            const uint32_t st = i % S, use = i / S;
            sbar_wait(s_u32(&bars[st]), use & 1);
            const uint8_t *codes = ring + (size_t)st * a.tma_stage_bytes;
            const float2 *cf = reinterpret_cast<const float2 *>(codes + code_bytes);
            const uint4 *rp[R];
            float2 co[R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int r = row_in_tile0 + j * GPW;
                rp[j] = reinterpret_cast<const uint4 *>(codes + (size_t)r * a.row_bytes);
                co[j] = gl == 0 ? cf[r] : make_float2(0.f, 0.f);
            }
            int A[R], Bm[R], Cl[R];
#pragma unroll
            for (int j = 0; j < R; ++j) A[j] = Bm[j] = Cl[j] = 0;
#pragma unroll 2
            for (int c0 = gl; c0 < nch; c0 += TPR) {
                int c = c0 + shift;
                if (c >= nch) c -= nch;
                uint4 v[R];
#pragma unroll
                for (int j = 0; j < R; ++j) v[j] = rp[j][c];
                quant_chunk<DTYPE, R>(v, c, sd, plane_u4, A, Bm, Cl);
            }
            __syncwarp();
            if (lane == 0) sbar_arrive(s_u32(&bars[kTmaMaxStages + st]));   // this warp is done with the stage
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const uint64_t row = t * TR + (uint64_t)(row_in_tile0 + j * GPW);
                uint64_t key = quant_key<TPR>(A[j], Bm[j], Cl[j], gl == 0, co[j], qs, row);
                wc.offer(key, lane);
            }
        }
        // ragged tail (< TR rows): direct loads, in the CTA whose turn it would have been
        if (blockIdx.x == (unsigned)(n_tiles % gridDim.x)) {
            const uint64_t rows_per_wi = (uint64_t)GPW * R;
            for (uint64_t base = n_tiles * TR + (uint64_t)warp * rows_per_wi; base < a.n;
                 base += (uint64_t)kScanWarps * rows_per_wi) {
                const uint4 *rp[R];
                uint64_t rix[R];
                bool valid[R];
                float2 co[R];
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    uint64_t r = base + (uint64_t)j * GPW + g;
                    valid[j] = r < a.n;
                    rix[j] = r;
                    rp[j] = reinterpret_cast<const uint4 *>(a.rows + (valid[j] ? r : a.n - 1) * a.row_bytes);
                    co[j] = (valid[j] && gl == 0) ? __ldg(a.qcoef + r) : make_float2(0.f, 0.f);
                }
                int A[R], Bm[R], Cl[R];
#pragma unroll
                for (int j = 0; j < R; ++j) A[j] = Bm[j] = Cl[j] = 0;
                for (int c = gl; c < nch; c += TPR) {
                    uint4 v[R];
#pragma unroll
                    for (int j = 0; j < R; ++j) v[j] = ldg_stream_u4(rp[j] + c);
                    quant_chunk<DTYPE, R>(v, c, sd, plane_u4, A, Bm, Cl);
                }
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    uint64_t key = quant_key<TPR>(A[j], Bm[j], Cl[j], valid[j] && gl == 0, co[j], qs, rix[j]);
                    wc.offer(key, lane);
                }
            }
        }
    }
    wc.finish(lists, warp, lane, consumer);
    cta_merge_and_store(lists, KP, a.partial + ((size_t)b * a.G + blockIdx.x) * KP);
}

// ----------------------------------------------------------------------------
// dispatch
// ----------------------------------------------------------------------------


static int pick_tpr(int nch, int dtype) {
    int t = 1;
    static int env_div = -1;
    if (env_div < 0) { const char *e = getenv("EVDB_SCAN_CPL"); env_div = e && atoi(e) > 0 ? atoi(e) : 0; }
    (void)dtype;
    const int div = env_div ? env_div : 4;  // measured on B200 (tools/sweep.py): 4 beats 2 for every dtype
    while (t < 32 && (t < 2 ? nch >= 2 : t * 2 <= nch / div)) t <<= 1;  // >= 2 lanes per row (whole sectors), then >= `div` chunks per lane
    return t;
}

template <int METRIC, int DTYPE>
static scan_fn_t pick_float(int tpr) {
    switch (tpr) {
        case 1: return scan_float_kernel<METRIC, DTYPE, 1, 4>;
        case 2: return scan_float_kernel<METRIC, DTYPE, 2, 4>;
        case 4: return scan_float_kernel<METRIC, DTYPE, 4, 4>;
        case 8: return scan_float_kernel<METRIC, DTYPE, 8, 4>;
        case 16: return scan_float_kernel<METRIC, DTYPE, 16, 4>;
        default: return scan_float_kernel<METRIC, DTYPE, 32, 4>;
    }
}
template <int DTYPE>
static scan_fn_t pick_quant(int tpr) {
    switch (tpr) {
        case 1: return scan_quant_kernel<DTYPE, 1, 4>;
        case 2: return scan_quant_kernel<DTYPE, 2, 4>;
        case 4: return scan_quant_kernel<DTYPE, 4, 4>;
        case 8: return scan_quant_kernel<DTYPE, 8, 4>;
        case 16: return scan_quant_kernel<DTYPE, 16, 4>;
        default: return scan_quant_kernel<DTYPE, 32, 2>;
    }
}

constexpr int kTmaR = 2;   // rows per lane group in flight: 2 beat 4 on every shape measured (tools/sweep_quant.sh)
template <int DTYPE>
static scan_fn_t pick_quant_tma(int tpr) {
    switch (tpr) {
        case 1: return scan_quant_tma_kernel<DTYPE, 1, kTmaR>;
        case 2: return scan_quant_tma_kernel<DTYPE, 2, kTmaR>;
        case 4: return scan_quant_tma_kernel<DTYPE, 4, kTmaR>;
        case 8: return scan_quant_tma_kernel<DTYPE, 8, kTmaR>;
        case 16: return scan_quant_tma_kernel<DTYPE, 16, kTmaR>;
        default: return scan_quant_tma_kernel<DTYPE, 32, kTmaR>;
    }
}

// Shared-memory wavefronts of one warp's loads (row chunks + query digits) for a candidate chunk
// rotation `m`: lane (g, gl) reads chunk (it*TPR + gl + ((g*m)&7)) mod nch of row g; an LDS.128
// serves a quarter-warp per wavefront when its 8 lanes touch 8 distinct 16-byte bank groups.
static int bank_cost(int nch, int tpr, int m, int R, int digit_loads) {
    int cost = 0;
    for (int it = 0; it * tpr < nch; ++it)
        for (int q = 0; q < 4; ++q) {
            int rowu[8] = {0}, sdu[8] = {0}, sd_addr[8][8], sd_n[8] = {0};
            for (int l = 8 * q; l < 8 * q + 8; ++l) {
                int g = l / tpr, gl = l % tpr, c0 = it * tpr + gl;
                if (c0 >= nch) continue;
                int sh = (g * m) & 7;
                if (sh >= nch) sh %= nch;
                int c = c0 + sh;
                if (c >= nch) c -= nch;
                rowu[(g * nch + c) & 7]++;
                int u = c & 7;
                bool seen = false;
                for (int i = 0; i < sd_n[u]; ++i) seen |= sd_addr[u][i] == c;
                if (!seen) { sd_addr[u][sd_n[u]++] = c; sdu[u]++; }
            }
            int mr = 0, ms = 0;
            for (int u = 0; u < 8; ++u) { mr = rowu[u] > mr ? rowu[u] : mr; ms = sdu[u] > ms ? sdu[u] : ms; }
            cost += R * mr + digit_loads * ms;
        }
    return cost;
}

// Tile plan of the TMA-staged quantized scan: pure host arithmetic (no device state), exported for
// the CPU tests as evdb_debug_scan_tile_plan.  Returns false when the register-fed kernel should run.
struct TmaTilePlan {
    int tpr, wt, stages, stage_bytes, tile_rows, bank_mul;
    size_t smem;   // dynamic shared memory of the launch: ring + query digits + candidate lists
};

static bool choose_tma_tile_plan(int dtype, int nch, size_t row_bytes, int KP, uint64_t count, int sm_count,
                                 TmaTilePlan *o) {
    const bool u8 = dtype == EVDB_U8;
    if (!u8 && dtype != EVDB_U4) return false;
    if (row_bytes < 64) return false;   // 32-byte rows: the register-fed kernel is level or better
    static int env_div = -1, env_wt = -1;
    if (env_div < 0) { const char *e = getenv("EVDB_SCAN_TMA_DIV"); env_div = e && atoi(e) > 0 ? atoi(e) : 4; }
    if (env_wt < 0) { const char *e = getenv("EVDB_SCAN_TMA_WT"); env_wt = e ? atoi(e) : 0; }
    int t2 = 1;
    while (t2 < 32 && (t2 < 2 ? nch >= 2 : t2 * 2 <= nch / env_div)) t2 <<= 1;
    const int gpw = 32 / t2;
    const size_t qbytes = (size_t)nch * (u8 ? 16 : 32) * kQPlanes;
    const size_t fixed = qbytes + scan_list_bytes(KP) + 1024;
    const size_t half = 112 * 1024, whole = 224 * 1024;
    int best_wt = 0, best_s = 0;
    size_t best_stage = 0;
    // preference: two CTAs per SM, then the widest tile group.  A stage always belongs to the same
    // group (stages % groups == 0: a group never waits on a barrier phase it did not see complete),
    // and every group has a second stage in flight.
    for (int pass = 0; pass < 2 && !best_wt; ++pass)
        for (int wt = kScanWarps; wt >= 1 && !best_wt; wt >>= 1) {
            if (env_wt && wt != env_wt) continue;
            const int ng = kScanWarps / wt;
            const size_t tile = (size_t)wt * gpw * kTmaR * (row_bytes + 8);
            const size_t stage = (tile + 127) / 128 * 128;
            const size_t room = pass == 0 ? half : whole;
            if (tile > 48 * 1024 || fixed + stage > room) continue;
            int stages = (int)((room - fixed) / stage);
            if (stages > kTmaMaxStages) stages = kTmaMaxStages;
            stages -= stages % ng;
            if (stages < (ng == 1 ? 3 : 2 * ng)) continue;
            best_wt = wt; best_s = stages; best_stage = stage;
        }
    if (!best_wt) return false;
    const uint64_t tile_rows = (uint64_t)best_wt * gpw * kTmaR;
    if (count / tile_rows < (uint64_t)4 * sm_count) return false;
    int best = 0, best_cost = bank_cost(nch, t2, 0, kTmaR, (u8 ? 1 : 2) * kQPlanes);
    if (t2 < 8)
        for (int m = 1; m < 8; ++m) {
            const int c = bank_cost(nch, t2, m, kTmaR, (u8 ? 1 : 2) * kQPlanes);
            if (c < best_cost) { best_cost = c; best = m; }
        }
    o->tpr = t2;
    o->wt = best_wt;
    o->stages = best_s;
    o->stage_bytes = (int)best_stage;
    o->tile_rows = (int)tile_rows;
    o->bank_mul = best;
    o->smem = qbytes + scan_list_bytes(KP) + (size_t)best_s * best_stage;
    return true;
}

struct ScanPlan {
    scan_fn_t fn;
    size_t smem;
    int tpr, rows_per_wi, threads;
    int mq;            // queries per pass of the multi-query float scan (0: one query per pass)
    bool tma;
    int stages, stage_bytes, bank_mul, tile_rows, wt;
};

static int scan_tma_mode() {   // EVDB_SCAN_TMA=0 forces the register-fed quantized scan (A/B measurements)
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("EVDB_SCAN_TMA"); mode = e ? atoi(e) : 1; }
    return mode;
}

static int scan_mq_mode() {   // EVDB_SCAN_MQ=0: batches as B single-query passes (A/B measurements); N = force N queries per pass
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("EVDB_SCAN_MQ"); mode = e ? atoi(e) : 1; }
    return mode;
}

static int make_scan_plan(evdb_store *s, int metric, int KP, int B, ScanPlan *p) {
    int tpr = pick_tpr(s->nch, s->dtype);
    p->tpr = tpr;
    p->threads = kScanWarps * 32;
    p->tma = false;
    p->mq = 0;
    p->stages = p->stage_bytes = p->bank_mul = p->tile_rows = 0;
    p->wt = kScanWarps;
    int R = 4;
    size_t qbytes;
    switch (s->dtype) {
        case EVDB_F32:
            p->fn = metric == EVDB_COSINE ? pick_float<EVDB_COSINE, EVDB_F32>(tpr)
                  : metric == EVDB_EUCLIDEAN ? pick_float<EVDB_EUCLIDEAN, EVDB_F32>(tpr)
                                             : pick_float<EVDB_MANHATTAN, EVDB_F32>(tpr);
            qbytes = (size_t)s->nch * 16;
            break;
        case EVDB_BF16:
            p->fn = metric == EVDB_COSINE ? pick_float<EVDB_COSINE, EVDB_BF16>(tpr)
                  : metric == EVDB_EUCLIDEAN ? pick_float<EVDB_EUCLIDEAN, EVDB_BF16>(tpr)
                                             : pick_float<EVDB_MANHATTAN, EVDB_BF16>(tpr);
            qbytes = (size_t)s->nch * 32;
            break;
        case EVDB_U8:
        case EVDB_U4: {
            if (metric != EVDB_COSINE) return EVDB_E_UNSUPPORTED;
            const bool u8 = s->dtype == EVDB_U8;
            p->fn = u8 ? pick_quant<EVDB_U8>(tpr) : pick_quant<EVDB_U4>(tpr);
            qbytes = (size_t)s->nch * (u8 ? 16 : 32) * kQPlanes;
            if (tpr == 32) R = 2;
            // TMA-staged variant: whole tiles through a shared-memory ring (large stores only:
            // a small one is latency-bound and spreads better one warp-iteration per warp)
            TmaTilePlan tp;
            if (scan_tma_mode() && choose_tma_tile_plan(s->dtype, s->nch, s->row_bytes, KP, s->count, s->sm_count, &tp)) {
                p->tma = true;
                tpr = tp.tpr;
                p->tpr = tp.tpr;
                p->fn = u8 ? pick_quant_tma<EVDB_U8>(tp.tpr) : pick_quant_tma<EVDB_U4>(tp.tpr);
                p->threads = (kScanWarps + 1) * 32;
                p->stages = tp.stages;
                p->stage_bytes = tp.stage_bytes;
                p->tile_rows = tp.tile_rows;
                p->wt = tp.wt;
                p->bank_mul = tp.bank_mul;
                R = kTmaR;
                qbytes += (size_t)tp.stages * tp.stage_bytes;
                static int dbg = -1;
                if (dbg < 0) { const char *e = getenv("EVDB_SCAN_DEBUG"); dbg = e && atoi(e) ? 1 : 0; }
                if (dbg)
                    fprintf(stderr, "[evdb scan] tma plan: nch=%d tpr=%d R=%d wt=%d stages=%d stage=%d B rotation=%d\n",
                            s->nch, tp.tpr, kTmaR, tp.wt, tp.stages, tp.stage_bytes, tp.bank_mul);
            }
            break;
        }
        default:
            return EVDB_E_BAD_ARG;
    }
    p->rows_per_wi = (32 / tpr) * R;
    p->smem = qbytes + scan_list_bytes(KP);
    if (p->smem > 200 * 1024 && !p->tma) return EVDB_E_UNSUPPORTED;
    // float stores, a batch, an append-buffer window: Q queries share every row load
    const int mode = scan_mq_mode();
    if ((s->dtype == EVDB_F32 || s->dtype == EVDB_BF16) && B >= 2 && KP <= kAppendMaxKP && mode != 0) {
        int Q = B >= 8 ? 8 : (B >= 4 ? 4 : 2);
        if (mode == 2 || mode == 4 || mode == 8) Q = mode;
        while (Q > 2 && (size_t)Q * p->smem > 100 * 1024) Q >>= 1;   // two CTAs per SM
        if ((size_t)Q * p->smem <= 200 * 1024) {
            p->mq = Q;
            p->fn = s->dtype == EVDB_F32 ? pick_float_mq_f32(metric, tpr, Q) : pick_float_mq_bf16(metric, tpr, Q);
            p->rows_per_wi = (32 / tpr) * (s->dtype == EVDB_BF16 ? (Q == 8 ? 2 : 4) : 8 / Q);
            p->smem = (size_t)Q * p->smem;
        }
    }
    return EVDB_OK;
}

int scan_grid_size(evdb_store *s, int metric, int KP, int B, int *G_out) {
    ScanPlan p;
    EVDB_TRY(make_scan_plan(s, metric, KP, B, &p));
    if (p.smem > 48 * 1024)
        EVDB_TRY(ensure_func_smem((const void *)p.fn, p.smem));
    int occ = 1;
    EVDB_TRY(cached_occupancy((const void *)p.fn, p.threads, p.smem, &occ));
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    uint64_t cap = (uint64_t)s->sm_count * occ;
    uint64_t want;
    if (p.tma) {
        want = s->count / p.tile_rows;   // one CTA per tile at most
    } else {
        uint64_t total_wi = (s->count + p.rows_per_wi - 1) / p.rows_per_wi;
        // small stores are latency-bound: one warp-iteration per warp spreads them over the most SMs
        want = (total_wi + (uint64_t)kScanWarps - 1) / (uint64_t)kScanWarps;
    }
    uint64_t G = want < cap ? want : cap;
    if (G < 1) G = 1;
    *G_out = (int)G;
    return EVDB_OK;
}

int scan_small_plan(evdb_store *s, int metric, int KP, int *G_out, int *tpr_out) {
    ScanPlan p;
    EVDB_TRY(make_scan_plan(s, metric, KP, 1, &p));
    *tpr_out = p.tpr;
    return scan_grid_size(s, metric, KP, 1, G_out);
}

int launch_scan(evdb_store *s, int metric, const ScanArgs &a0, cudaStream_t st) {
    ScanPlan p;
    EVDB_TRY(make_scan_plan(s, metric, a0.KP, a0.B, &p));
    ScanArgs a = a0;
    a.tma_stages = p.stages;
    a.tma_stage_bytes = p.stage_bytes;
    a.bank_mul = p.bank_mul;
    a.tma_wt = p.wt;
    if (p.smem > 48 * 1024) EVDB_TRY(ensure_func_smem((const void *)p.fn, p.smem));
    dim3 grid(a.G, p.mq ? (a.B + p.mq - 1) / p.mq : a.B);
    EVDB_CUDA(launch_chained(p.fn, grid, dim3(p.threads), p.smem, st, 1, a));
    s->n_launches++;
    return EVDB_OK;
}

}  // namespace evdb

// Diagnostics (CPU tests): the tile plan the TMA-staged quantized scan would use for a store of this
// shape.  Returns 1 and fills out[8] = {lanes per row, warps per tile, stages, stage bytes, rows per
// tile, chunk rotation, dynamic shared memory, consumer groups}, 0 if the register-fed kernel would
// run, a negative EVDB_E_* for bad arguments.  Touches no device.
extern "C" int evdb_debug_scan_tile_plan(int dtype, int dim, int window, uint64_t count, int sm_count, int32_t *out) {
    if (!out || dim <= 0 || window <= 0 || sm_count <= 0) return EVDB_E_BAD_ARG;
    if (dtype != EVDB_U8 && dtype != EVDB_U4) return EVDB_E_BAD_ARG;
    const int dpad = dtype == EVDB_U8 ? (dim + 15) / 16 * 16 : (dim + 31) / 32 * 32;
    const int nch = dtype == EVDB_U8 ? dpad / 16 : dpad / 32;
    const size_t row_bytes = dtype == EVDB_U8 ? (size_t)dpad : (size_t)dpad / 2;
    evdb::TmaTilePlan tp;
    if (!evdb::choose_tma_tile_plan(dtype, nch, row_bytes, window, count, sm_count, &tp)) return 0;
    out[0] = tp.tpr; out[1] = tp.wt; out[2] = tp.stages; out[3] = tp.stage_bytes;
    out[4] = tp.tile_rows; out[5] = tp.bank_mul; out[6] = (int32_t)tp.smem; out[7] = evdb::kScanWarps / tp.wt;
    return 1;
}
