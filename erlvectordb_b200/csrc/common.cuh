// common.cuh -- shared device/host helpers for libevdb_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/evdb.h"

namespace evdb {

constexpr int kWarp = 32;
constexpr uint64_t kKeyMax = 0xFFFFFFFFFFFFFFFFull;

// ---- thread-local CUDA error text (evdb_last_cuda_error) -------------------
void set_cuda_error(cudaError_t e, const char *file, int line);

#define EVDB_CUDA(expr)                                         \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) {                                \
            ::evdb::set_cuda_error(_e, __FILE__, __LINE__);     \
            return _e == cudaErrorMemoryAllocation ? EVDB_E_OOM \
                                                   : EVDB_E_CUDA; \
        }                                                       \
    } while (0)

#define EVDB_TRY(expr)             \
    do {                           \
        int _rc = (expr);          \
        if (_rc != EVDB_OK) return _rc; \
    } while (0)

// ---- candidate keys ----------------------------------------------------------
// A candidate is one u64: (order-preserving bits of the fp32 score) << 32 | slot.
// Unsigned compare of keys == lexicographic (score, slot) compare.
__host__ __device__ __forceinline__ uint32_t f32_orderable(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } cv; cv.f = f; uint32_t b = cv.u;
#endif
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_orderable(uint32_t o) {
    uint32_t b = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t slot) {
    return ((uint64_t)f32_orderable(score) << 32) | slot;
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) {
    return f32_from_orderable((uint32_t)(k >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_slot(uint64_t k) { return (uint32_t)k; }

// fp64 -> order-preserving u64 (for the exact full-sort plan)
__host__ __device__ __forceinline__ uint64_t f64_orderable(double d) {
#ifdef __CUDA_ARCH__
    uint64_t b = (uint64_t)__double_as_longlong(d);
#else
    union { double d; uint64_t u; } cv; cv.d = d; uint64_t b = cv.u;
#endif
    return b ^ ((b >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
}

// ---- counter-based synthetic corpus -------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// 24-bit grid in [-1,1): exact in fp32 and fp64
__host__ __device__ __forceinline__ float synth_value(uint64_t seed, uint64_t idx) {
    int32_t m = (int32_t)(mix64(seed ^ idx) >> 40) - 8388608;
    return (float)m * (1.0f / 8388608.0f);
}

// ---- streaming 128-bit loads -------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ int dp4a_su(int a_s8x4, uint32_t b_u8x4, int c) {
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s8x4), "r"(b_u8x4), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- programmatic dependent launch (the kernel chain of one search) -----------------------------
// A kernel launched with launch_chained() may be scheduled while its predecessor in the stream still
// runs; it must call pdl_wait() before it touches anything the predecessor wrote (or still reads), and
// a predecessor lets its dependent in with pdl_launch_dependents().  The launch latency of the next link
// then overlaps the tail of the previous one.  EVDB_PDL=0 launches every link the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chained(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         int cluster_x, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster_x > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cluster_x; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, fn, args...);
}

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline uint64_t round_up64(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }
static inline int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace evdb
