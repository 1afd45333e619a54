// dev microbenchmark: dependent-issue latency of FP64 instructions on the target (one warp, clock64 around an unrolled chain)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 512; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    double y = x;
#pragma unroll
    for (int i = 0; i < 512; ++i) y = y + a;
    long long t2 = clock64();
    double z = y;
#pragma unroll
    for (int i = 0; i < 64; ++i) z = a / z;
    long long t3 = clock64();
    double w = z;
#pragma unroll
    for (int i = 0; i < 64; ++i) w = sqrt(w + a);
    long long t4 = clock64();
    // 4 independent chains
    double p0 = w, p1 = w + 1, p2 = w + 2, p3 = w + 3;
#pragma unroll
    for (int i = 0; i < 256; ++i) { p0 = fma(p0, b, a); p1 = fma(p1, b, a); p2 = fma(p2, b, a); p3 = fma(p3, b, a); }
    long long t5 = clock64();
    out[threadIdx.x] = p0 + p1 + p2 + p3;
    if (threadIdx.x == 0) { cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = (t4 - t3); cyc[4] = t5 - t4; }
}
__global__ void lds_lat(double* out, long long* cyc, int stride) {
    __shared__ double s[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (double)((i + stride) & 1023);
    __syncwarp();
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) idx = (int)s[idx];
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int r = 0; r < 2; ++r) { lat<<<1, 32>>>(out, cyc, 1.000001, 0.999999); lds_lat<<<1, 32>>>(out, cyc, 7); cudaDeviceSynchronize(); }
    printf("DFMA dep %.1f cyc  DADD dep %.1f  DDIV dep %.1f  DSQRT(+add) dep %.1f  4-chain DFMA per-instr %.2f  LDS->cvt->LDS %.1f\n",
           cyc[0] / 512.0, cyc[1] / 512.0, cyc[2] / 64.0, cyc[3] / 64.0, cyc[4] / 1024.0, cyc[5] / 256.0);
    return 0;
}
