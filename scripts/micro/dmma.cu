// dev micro-benchmark: FP64 tensor-core MMA (mma.sync m8n8k4 / m16n8k8 .f64) on this GPU - dependent latency and
// throughput per SM with 1..16 warps per SM and 1..4 independent accumulator chains per warp; DFMA beside it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <int CH>
__global__ void k884(double* out, int iters, double a, double b) {
    double c[CH][2];
    for (int q = 0; q < CH; ++q) { c[q][0] = threadIdx.x + q; c[q][1] = q; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < CH; ++q) mma884(c[q][0], c[q][1], a, b);
    }
    double s = 0; for (int q = 0; q < CH; ++q) s += c[q][0] + c[q][1];
    if (s == 1.2345) out[0] = s;
}
template <int CH>
__global__ void k1688(double* out, int iters, double a, double b) {
    double c[CH][4], av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
    for (int q = 0; q < CH; ++q) for (int e = 0; e < 4; ++e) c[q][e] = threadIdx.x + q + e;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < CH; ++q) mma1688(c[q], av, bv);
    }
    double s = 0; for (int q = 0; q < CH; ++q) for (int e = 0; e < 4; ++e) s += c[q][e];
    if (s == 1.2345) out[0] = s;
}
template <int CH>
__global__ void kfma(double* out, int iters, double a, double b) {
    double c[CH];
    for (int q = 0; q < CH; ++q) c[q] = threadIdx.x + q;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < CH; ++q) c[q] = fma(c[q], a, b);
    }
    double s = 0; for (int q = 0; q < CH; ++q) s += c[q];
    if (s == 1.2345) out[0] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    double* out; cudaMalloc(&out, 8);
    const int iters = 2000;
    printf("%s, %d SMs, %.3f GHz\n", p.name, nsm, ghz);
    for (int wps : {1, 2, 4, 8, 16, 32}) {
        const int threads = 32 * (wps > 8 ? 8 : wps), blocks = nsm * (wps > 8 ? wps / 8 : 1);
#define RUN(K, CH, FMAS, NAME) { float ms = timeit([&] { K<CH><<<blocks, threads>>>(out, iters, 0.999, 1e-9); }); \
        double n = (double)iters * 8 * CH; double cyc = ms * 1e-3 * ghz * 1e9 / n; \
        printf("%-8s warps/SM %2d chains %d: %7.2f cycles per instr per warp, %7.1f FMA/clk/SM, %6.2f TFMA/s\n", NAME, wps, CH, cyc, \
               (double)FMAS * n * wps / (ms * 1e-3 * ghz * 1e9), (double)FMAS * n * wps * nsm / (ms * 1e-3) * 1e-12); }
        RUN(k884, 1, 256, "m8n8k4") RUN(k884, 2, 256, "m8n8k4") RUN(k884, 4, 256, "m8n8k4")
        RUN(k1688, 1, 1024, "m16n8k8") RUN(k1688, 2, 1024, "m16n8k8")
        RUN(kfma, 1, 32, "dfma") RUN(kfma, 4, 32, "dfma") RUN(kfma, 8, 32, "dfma")
    }
    return 0;
}
