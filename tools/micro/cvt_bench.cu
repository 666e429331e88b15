// Microbenchmark: throughput of F2F.F64.F32 (float -> double) against the integer widening used by the
// exact re-rank, and of DMUL, per SM as a function of resident warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_bench cvt_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double widen_int(unsigned u) {
    const unsigned hi = ((u >> 3) & 0x0FFFFFFFu) + 0x38000000u + (u & 0x80000000u);
    return __hiloint2double((int)hi, (int)(u << 29));
}

template <int MODE>
__global__ void k(const float *in, double *out, long long *cyc, int iters) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = in[threadIdx.x * 8 + i];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double d;
            if (MODE == 0) d = (double)x[i];                       // F2F.F64.F32
            else if (MODE == 1) d = widen_int(__float_as_uint(x[i]));
            else d = __dmul_rn(acc[i], 1.0000001);                  // DMUL only
            if (MODE == 2) acc[i] = d;
            else acc[i] = __longlong_as_double(__double_as_longlong(acc[i]) ^ __double_as_longlong(d));
            x[i] = __uint_as_float(__float_as_uint(x[i]) + 0x10u);
        }
    }
    long long t1 = clock64();
    double r = 0;
    for (int i = 0; i < 8; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float *d_in; double *d_out; long long *d_cyc;
    cudaMalloc(&d_in, 4 * 8 * 1024); cudaMalloc(&d_out, 8 * 1024); cudaMalloc(&d_cyc, 64);
    cudaMemset(d_in, 0x3f, 4 * 8 * 1024);
    const int iters = 2048;
    const char *names[3] = {"F2F.F64.F32", "integer widen", "DMUL"};
    for (int mode = 0; mode < 3; ++mode)
        for (int w : {1, 4, 8, 16, 32}) {
            if (mode == 0) k<0><<<1, w * 32>>>(d_in, d_out, d_cyc, iters);
            else if (mode == 1) k<1><<<1, w * 32>>>(d_in, d_out, d_cyc, iters);
            else k<2><<<1, w * 32>>>(d_in, d_out, d_cyc, iters);
            cudaDeviceSynchronize();
            long long c;
            cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-14s warps/SM=%2d : %.2f cycles per warp-level conversion on the SM\n", names[mode], w,
                   (double)c / iters / 8 / w);
        }
    return 0;
}
