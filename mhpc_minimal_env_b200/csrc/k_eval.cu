// Standalone batched evaluation of the light generated functions (parity tests of the device code
// generator against the reference's compiled C) and the FP64 FMA peak probe used as the roofline
// denominator of the linearisation kernels.
#include <string.h>

#include "kernels.h"
#include "model.cuh"

namespace mhpc {

enum { EV_DYN_BS = 0, EV_DYN_FS, EV_DYN_FL, EV_IMP_F, EV_IMP_B, EV_FB, EV_FB_PAR, EV_TC1, EV_TC2, EV_JF, EV_JB, EV_LINK0, EV_LINK1, EV_LINK2, EV_LINK3, EV_LINK4 };

__global__ void k_eval(int fn, int n, const double* in0, const double* in1, const double* in2, const double* in3,
                       double* o0, double* o1, double* o2, double* o3) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (fn <= EV_DYN_FL) {
        double x[14], u[4], xd[14], y[4];
        for (int e = 0; e < 14; ++e) x[e] = in0[(size_t)t * 14 + e];
        for (int e = 0; e < 4; ++e) u[e] = in1[(size_t)t * 4 + e];
        if (fn == EV_DYN_BS) dyn_bs(x, u, xd, y); else if (fn == EV_DYN_FS) dyn_fs(x, u, xd, y); else dyn_fl(x, u, xd, y);
        for (int e = 0; e < 14; ++e) o0[(size_t)t * 14 + e] = xd[e];
        for (int e = 0; e < 4; ++e) o1[(size_t)t * 4 + e] = y[e];
    } else if (fn <= EV_IMP_B) {
        double x[14], xp[14], y[4];
        for (int e = 0; e < 14; ++e) x[e] = in0[(size_t)t * 14 + e];
        if (fn == EV_IMP_F) imp_f(x, xp, y); else imp_b(x, xp, y);
        for (int e = 0; e < 14; ++e) o0[(size_t)t * 14 + e] = xp[e];
        for (int e = 0; e < 4; ++e) o1[(size_t)t * 4 + e] = y[e];
    } else if (fn == EV_FB || fn == EV_FB_PAR) {
        double x[6], u[4], fh[4], ct[2];
        for (int e = 0; e < 6; ++e) x[e] = in0[(size_t)t * 6 + e];
        for (int e = 0; e < 4; ++e) { u[e] = in1[(size_t)t * 4 + e]; fh[e] = in2[(size_t)t * 4 + e]; }
        for (int e = 0; e < 2; ++e) ct[e] = in3[(size_t)t * 2 + e];
        if (fn == EV_FB) {
            double xd[6];
            fb_dyn(x, u, fh, ct, xd);
            for (int e = 0; e < 6; ++e) o0[(size_t)t * 6 + e] = xd[e];
        } else {
            double a2[2], b8[8];
            fb_dyn_par(x, u, fh, ct, a2, b8);
            double* A = o0 + (size_t)t * 36;
            double* B = o1 + (size_t)t * 24;
            for (int e = 0; e < 36; ++e) A[e] = 0.0;
            for (int e = 0; e < 24; ++e) B[e] = 0.0;
            A[5 + 6 * 0] = a2[0]; A[5 + 6 * 1] = a2[1]; A[0 + 6 * 3] = 1.0; A[1 + 6 * 4] = 1.0; A[2 + 6 * 5] = 1.0;
            const int rows[8] = {3, 5, 4, 5, 3, 5, 4, 5}, cols[8] = {0, 0, 1, 1, 2, 2, 3, 3};
            for (int q = 0; q < 8; ++q) B[rows[q] + 6 * cols[q]] = b8[q];
        }
    } else if (fn == EV_TC1 || fn == EV_TC2) {
        double x[14], h, hx[14], hxx[196];
        for (int e = 0; e < 14; ++e) { x[e] = in0[(size_t)t * 14 + e]; hx[e] = 0.0; }
        for (int e = 0; e < 196; ++e) hxx[e] = 0.0;
        if (fn == EV_TC1) tconstr_fl1(x, &h, hx, hxx); else tconstr_fl2(x, &h, hx, hxx);
        o0[t] = h;
        for (int e = 0; e < 14; ++e) o1[(size_t)t * 14 + e] = hx[e];
        for (int e = 0; e < 196; ++e) o2[(size_t)t * 196 + e] = hxx[e];
    } else if (fn >= EV_LINK0) {
        double x[14], pt[2], J[14], Jd[14];
        for (int e = 0; e < 14; ++e) { x[e] = in0[(size_t)t * 14 + e]; J[e] = 0.0; Jd[e] = 0.0; }
        pt[0] = in1[(size_t)t * 2]; pt[1] = in1[(size_t)t * 2 + 1];
        link_jacobian(fn - EV_LINK0, x, pt, J, Jd);
        for (int e = 0; e < 14; ++e) { o0[(size_t)t * 14 + e] = J[e]; o1[(size_t)t * 14 + e] = Jd[e]; }
    } else {
        double x[14], J[14];
        for (int e = 0; e < 14; ++e) { x[e] = in0[(size_t)t * 14 + e]; J[e] = 0.0; }
        if (fn == EV_JF) jacob_f(x, J); else jacob_b(x, J);
        for (int e = 0; e < 14; ++e) o0[(size_t)t * 14 + e] = J[e];
    }
}

int eval_fn_id(const char* name) {
    const char* names[] = {"Dyn_BS", "Dyn_FS", "Dyn_FL", "Imp_F", "Imp_B", "FBDynamics", "FBDynamics_par",
                           "WB_FL1_terminal_constr", "WB_FL2_terminal_constr", "Jacob_F", "Jacob_B",
                           "Link0Jacobian", "Link1Jacobian", "Link2Jacobian", "Link3Jacobian", "Link4Jacobian"};
    for (int i = 0; i < 16; ++i)
        if (!strcmp(name, names[i])) return i;
    return -1;
}

int launch_eval(int fn, int n, const double* in0, const double* in1, const double* in2, const double* in3,
                double* o0, double* o1, double* o2, double* o3, cudaStream_t s) {
    k_eval<<<(n + 63) / 64, 64, 0, s>>>(fn, n, in0, in1, in2, in3, o0, o1, o2, o3);
    return 0;
}

// ---- FP64 FMA peak probe: 8 independent dependent-chains per thread ----------------------------
__global__ void __launch_bounds__(256) k_dfma_probe(double* out, int iters, double a, double b) {
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
        }
    }
    const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 123.456) out[0] = s;
}

double measure_fp64_peak_tflops(int device) {
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double* out;
    cudaMalloc(&out, sizeof(double));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma_probe<<<blocks, threads>>>(out, 64, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_dfma_probe<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    return flops / (best * 1e-3) / 1e12;
}

}  // namespace mhpc
