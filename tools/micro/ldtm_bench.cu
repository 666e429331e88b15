// Microbenchmark: tcgen05.ld (LDTM) throughput per SM as a function of warps, vector width and
// loads in flight.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench ldtm_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t (&v)[X]);
template <>
__device__ __forceinline__ void ldtm<32>(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ldtm<16>(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// WARPS warps (4..16, multiple of 4) stream the 512 TMEM columns ITER times, INFL loads in flight per warp
template <int X, int INFL>
__global__ void __launch_bounds__(512) k(int warps, int iters, unsigned long long *out, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < warps) {
        // warps sharing a lane quarter split the columns
        const int share = warps / 4, part = warp / 4;
        const int cols = 512 / share, c0 = part * cols;
        uint32_t v[INFL][X];
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int c = 0; c < cols; c += X * INFL) {
#pragma unroll
                for (int f = 0; f < INFL; ++f) ldtm<X>(base + c0 + c + f * X, v[f]);
                ldwait();
#pragma unroll
                for (int f = 0; f < INFL; ++f)
#pragma unroll
                    for (int j = 0; j < X; ++j) acc ^= v[f][j];
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    if (lane == 0 && warp < warps) out[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int X, int INFL>
void run(int warps, unsigned long long *d_out, uint32_t *d_sink) {
    const int iters = 200;
    k<X, INFL><<<148, 512>>>(warps, iters, d_out, d_sink);
    cudaDeviceSynchronize();
    unsigned long long h[148 * 16];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < 148; ++b) for (int w = 0; w < warps; ++w) if ((double)h[b * 16 + w] > mx) mx = (double)h[b * 16 + w];
    const double bytes = 128.0 * 512 * 4 * iters;  // whole TMEM per iteration
    printf("x%-3d inflight=%d warps=%2d : %.1f B/clk/SM (%.0f cycles)  err=%s\n", X, INFL, warps, bytes / mx, mx,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    unsigned long long *d_out; uint32_t *d_sink;
    cudaMalloc(&d_out, 148 * 16 * 8); cudaMalloc(&d_sink, 4096);
    for (int warps : {4, 8, 16}) {
        run<16, 1>(warps, d_out, d_sink); run<16, 2>(warps, d_out, d_sink); run<16, 4>(warps, d_out, d_sink);
        run<32, 1>(warps, d_out, d_sink); run<32, 2>(warps, d_out, d_sink); run<32, 4>(warps, d_out, d_sink);
    }
    return 0;
}
