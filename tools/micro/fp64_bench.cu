// Microbenchmark: dependent-chain latency and issue interval of the non-tensor fp64 unit (DADD/DMUL)
// on one SM, as a function of warps per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_bench fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(double *out, long long *cyc, int iters) {
    double s[CHAINS];
    for (int c = 0; c < CHAINS; ++c) s[c] = threadIdx.x * 1e-9 + c;
    const double inc = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) s[c] = __dadd_rn(s[c], inc);
    }
    long long t1 = clock64();
    double r = 0;
    for (int c = 0; c < CHAINS; ++c) r += s[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
void run(int warps, double *d_out, long long *d_cyc) {
    const int iters = 4096;
    k<CHAINS><<<1, warps * 32>>>(d_out, d_cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM=%2d independent chains/thread=%d : %.1f cycles per dependent DADD step, %.1f cycles per warp-instruction on the SM\n",
           warps, CHAINS, (double)c / iters, (double)c / iters / (CHAINS * warps));
}

int main() {
    double *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 8 * 2048); cudaMalloc(&d_cyc, 8 * 16);
    for (int w : {1, 2, 4, 8, 16, 32}) run<1>(w, d_out, d_cyc);
    for (int w : {1, 4}) run<4>(w, d_out, d_cyc);
    for (int w : {1, 4}) run<8>(w, d_out, d_cyc);
    return 0;
}
