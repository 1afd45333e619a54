// dev microbenchmark: FP64 issue rate per SM as a function of warps per SM and independent chains per warp
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, long long* cyc, double a, double b) {
    double p[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) p[c] = a + threadIdx.x + c;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
#pragma unroll
            for (int c = 0; c < CH; ++c) p[c] = fma(p[c], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += p[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH>
void run(int warps, double* out, long long* cyc) {
    k<CH><<<1, 32 * warps>>>(out, cyc, 1.000001, 0.999999);
    cudaDeviceSynchronize();
    k<CH><<<1, 32 * warps>>>(out, cyc, 1.000001, 0.999999);
    cudaDeviceSynchronize();
    const double n = 64.0 * 16 * CH;
    printf("warps/SM %2d chains %d: %.2f cyc per DFMA per warp, %.1f DFMA lanes/clk/SM\n", warps, CH, cyc[0] / n, 32.0 * warps * n / cyc[0]);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 8);
    for (int w : {1, 2, 4, 8, 16, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
    return 0;
}
