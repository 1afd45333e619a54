// Impact-aware Riccati backward sweep (MultiPhaseDDP.cpp:100-127, 300-341; SinglePhase.cpp:183-216;
// MHPC_CompoundTypes.h:116-144) with the regularisation retry loop of MultiPhaseDDP::solve
// (MultiPhaseDDP.cpp:209-227) folded in.
//
// A half-warp owns one instance, lane r owns row r of the value Hessian H (registers).  Time is
// sequential; every warp streams the knot records of its two instances into warp-private shared
// memory with 16-byte cp.async, one knot ahead of the math (double buffered, synchronised with
// __syncwarp only: warps never wait for each other, there is no CTA barrier in the sweep).
// The structure of the reference's matrices is used instead of dense 14x14 products:
//   A = [I7  dt*I7 ; 0(7x2) Ab(7x12)],  B = [0 ; Bb],  C/D = two stance rows,  lxx/luu diagonal,
// which removes only exact-zero / exact-one terms of the dense sums.
#include <cuda_pipeline.h>

#include "kernels.h"
#include "model.cuh"

namespace mhpc {

constexpr int BWS_THREADS = 128;             // 4 warps = 8 instances per CTA
constexpr int BWS_INST = BWS_THREADS / 16;
constexpr int MQ_LD = 15;
// per-instance scratch (doubles)
constexpr int S_G = 0;        // 16
constexpr int S_MQ = 16;      // 14 x 15 : M = H*A, later Qxx for the symmetrisation
constexpr int S_N = 226;      // 7 x 4  : rows 7..13 of H*B   (RB: 6 x 4)
constexpr int S_QUX = 254;    // 4 x 14
constexpr int S_QUU = 310;    // 16
constexpr int S_QU = 326;     // 4
// the three 4x4 work matrices of the value update live in the M/Qxx area, which is dead once Qxx has been symmetrised
constexpr int S_ADJ = S_MQ;         // 16 adjugate of Quu
constexpr int S_INV = S_MQ + 16;    // 16 Quu^-1
constexpr int S_QI = S_MQ + 32;     // 16 symmetrised Quu^-1
constexpr int SCR = 331;
#ifndef BWS_MINBLOCKS
#define BWS_MINBLOCKS 4
#endif
constexpr size_t BWS_SMEM = (size_t)(2 * REC + SCR + 1) * BWS_INST * sizeof(double);

// Eigen::LDLT<>::isPositive() restated (diagonal pivoting on the largest |a_ii|, sign of D tracked;
// Eigen/src/Cholesky/LDLT.h); call site SinglePhase.cpp:202-205.
__device__ __forceinline__ bool ldlt4_is_positive(const double q[4][4]) {
    // work on a full symmetric copy of the LOWER triangle (the only part Eigen reads) so that the symmetric pivot
    // swaps are plain row + column swaps with compile-time indices (everything stays in registers)
    double a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) a[i][jj] = (i >= jj) ? q[i][jj] : q[jj][i];
    int sign = 0;  // 0 zero, 1 pos, 2 neg, 3 indefinite
    bool result_known = false, result = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int pv = k;
        double big = fabs(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i)
            if (fabs(a[i][i]) > big) { big = fabs(a[i][i]); pv = i; }
#pragma unroll
        for (int pp = k + 1; pp < 4; ++pp) {
            if (pv == pp) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { const double t = a[k][c]; a[k][c] = a[pp][c]; a[pp][c] = t; }
#pragma unroll
                for (int r = 0; r < 4; ++r) { const double t = a[r][k]; a[r][k] = a[r][pp]; a[r][pp] = t; }
            }
        }
        if (k > 0) {
            double temp[3];
#pragma unroll
            for (int jj = 0; jj < k; ++jj) temp[jj] = a[jj][jj] * a[k][jj];
            double sacc = a[k][0] * temp[0];
#pragma unroll
            for (int jj = 1; jj < k; ++jj) sacc += a[k][jj] * temp[jj];
            a[k][k] -= sacc;
#pragma unroll
            for (int i = k + 1; i < 4; ++i) {
                double t = a[i][0] * temp[0];
#pragma unroll
                for (int jj = 1; jj < k; ++jj) t += a[i][jj] * temp[jj];
                a[i][k] -= t;
            }
        }
        const double akk = a[k][k];
        const bool valid = fabs(akk) > 0.0;
        if (k == 0 && !valid && !result_known) { result_known = true; result = true; }
        if (valid) {
            // one reciprocal per pivot instead of a division per entry (<= 1 ulp on L; only the SIGNS of the pivots are used)
            const double rk = 1.0 / akk;
#pragma unroll
            for (int i = k + 1; i < 4; ++i) a[i][k] *= rk;
        }
        if (sign == 1) { if (akk < 0) sign = 3; }
        else if (sign == 2) { if (akk > 0) sign = 3; }
        else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = 2; }
    }
    return result_known ? result : (sign == 1 || sign == 0);
}

// value-function update shared by the WB and RB steps (MHPC_CompoundTypes.h:128-144).  Called by ALL lanes of the warp
// (`on` = this half-warp takes part).  Quu^-1 is the closed-form cofactor inverse (what a fixed-size Matrix4d::inverse()
// computes): lane l forms ONE adjugate entry and does ONE division, the 16 entries are shared through scratch.
// In: Quu/Qu/Qux in scratch, this lane's Qux column qux[4], Qx value qx, symmetrised Qxx row in Hn[].
// Out: Hn[] = H row, G entry, K column, du, dV of the step.
template <int NX>
__device__ __forceinline__ bool value_update(bool on, int l, double* scr, const double qux[4], double qx, double Hn[14],
                                             double quu_shift, double kcol[4], double du[4], double& Gout, double& dVstep) {
    bool pd = true;
    if (on) {
        double tst[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) tst[r][c] = scr[S_QUU + r + 4 * c] - (r == c ? 1.0 * quu_shift : 0.0);
        pd = ldlt4_is_positive(tst);
        // adjugate entry r[jj][i] from the cofactor of (i, jj)
        const int i = l & 3, jj = l >> 2;
        const int i1 = (i + 1) & 3, i2 = (i + 2) & 3, i3 = (i + 3) & 3, j1 = (jj + 1) & 3, j2 = (jj + 2) & 3, j3 = (jj + 3) & 3;
#define QM(r, c) scr[S_QUU + (r) + 4 * (c)]
        const double cf = QM(i1, j1) * (QM(i2, j2) * QM(i3, j3) - QM(i2, j3) * QM(i3, j2)) +
                          QM(i2, j1) * (QM(i3, j2) * QM(i1, j3) - QM(i3, j3) * QM(i1, j2)) +
                          QM(i3, j1) * (QM(i1, j2) * QM(i2, j3) - QM(i1, j3) * QM(i2, j2));
        scr[S_ADJ + jj + 4 * i] = ((i + jj) & 1) ? -cf : cf;
    }
    __syncwarp();
    if (on) {
        double det = QM(0, 0) * scr[S_ADJ + 0 + 4 * 0];
#pragma unroll
        for (int i = 1; i < 4; ++i) det += QM(i, 0) * scr[S_ADJ + 0 + 4 * i];
        // adj / det; a zero numerator (every off-diagonal entry in the rigid-body phases) would send the FP64 division
        // through its slow path (quotient below the fast path's range check): 0 * det has the same value and sign
        const double adj = scr[S_ADJ + l];
        scr[S_INV + l] = (adj == 0.0) ? adj * det : adj / det;
#undef QM
    }
    __syncwarp();
    if (on) {
        const int r = l & 3, c = l >> 2;
        scr[S_QI + l] = (scr[S_INV + r + 4 * c] + scr[S_INV + c + 4 * r]) / 2;
    }
    __syncwarp();
    if (!on || !pd) return !on ? true : false;
    double Qu[4], T[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) Qu[c] = scr[S_QU + c];
#pragma unroll
    for (int c = 0; c < 4; ++c) {   // Qi(r, c) at S_QI + r + 4c
        double sacc = (-scr[S_QI + c + 4 * 0]) * qux[0], t = qux[0] * scr[S_QI + 0 + 4 * c], e = (-scr[S_QI + c + 4 * 0]) * Qu[0];
#pragma unroll
        for (int cc = 1; cc < 4; ++cc) {
            sacc += (-scr[S_QI + c + 4 * cc]) * qux[cc];
            t += qux[cc] * scr[S_QI + cc + 4 * c];
            e += (-scr[S_QI + c + 4 * cc]) * Qu[cc];
        }
        kcol[c] = sacc; T[c] = t; du[c] = e;
    }
    double g = T[0] * Qu[0];
#pragma unroll
    for (int c = 1; c < 4; ++c) g += T[c] * Qu[c];
    Gout = qx - g;
    if (l < NX) {
#pragma unroll
        for (int jc = 0; jc < NX; ++jc) {
            double sacc = T[0] * scr[S_QUX + jc];
#pragma unroll
            for (int c = 1; c < 4; ++c) sacc += T[c] * scr[S_QUX + c * 14 + jc];
            Hn[jc] = Hn[jc] - sacc;
        }
    }
    double dv = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {   // dV = ((-Qu^T) Quu^-1) Qu with the un-symmetrised inverse
        double sacc = (-Qu[0]) * scr[S_INV + 0 + 4 * c];
#pragma unroll
        for (int cc = 1; cc < 4; ++cc) sacc += (-Qu[cc]) * scr[S_INV + cc + 4 * c];
        dv = (c == 0) ? sacc * Qu[0] : dv + sacc * Qu[c];
    }
    dVstep = dv;
    return true;
}

__device__ __forceinline__ int knot_rec_len(bool wb, bool term) {
    return wb ? (term ? WB_TERM_LEN : WB_STEP_LEN) : (term ? RB_TERM_LEN : RB_STEP_LEN);
}

__global__ void __launch_bounds__(BWS_THREADS, BWS_MINBLOCKS) k_bws(Dev d, Consts c, Opt opt, int mask_mode, int single_attempt) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int hw = tid >> 4, l = tid & 15;          // half-warp = instance slot in the CTA, lane within
    const int i = blockIdx.x * BWS_INST + hw;
    double* bufs = smem + (size_t)hw * 2 * REC;                        // [2 stages][REC] of my instance
    double* scr = smem + (size_t)BWS_INST * 2 * REC + (size_t)hw * SCR;
    const size_t es = d.npad;
    const bool valid = i < d.n;
    const bool enabled = valid && inst_enabled(d, i, mask_mode);
    const int np = valid ? d.n_phases[i] : 0, nwb = valid ? d.n_wb[i] : 0, nk = enabled ? d.n_knots[i] : 0;
    const double dtw = valid ? d.dt_wb[i] : 0.0, dtf = valid ? d.dt_fb[i] : 0.0;
    double reg = enabled ? d.sd[(size_t)SD_REG * es + i] : 0.0;
    int n_bws = 1;
    bool done = !enabled;
    const int ktb = max(nk, __shfl_xor_sync(0xffffffffu, nk, 16));     // knots this warp walks
    const double* lin_i = d.lin + (size_t)(valid ? i : 0) * d.KT * REC;
    // lane-constant parts of the output addresses (SoA over instances): only `kk * stride` is added per knot
    double* const gv_lane = d.Gv + (size_t)l * es + (valid ? i : 0);
    double* const kf_lane = d.Kfb + (size_t)(4 * l) * es + (valid ? i : 0);
    double* const du_lane = d.dufb + (size_t)l * es + (valid ? i : 0);
    const size_t gv_knot = 14 * es, kf_knot = (size_t)KDIM * es, du_knot = 4 * es;

    // per-instance schedule walker (phase p, in-phase knot k; knot slot = p*KSTRIDE + k), used both by the prefetcher
    // (one knot ahead) and by the math; it visits the nk used knots from the last one backwards
    struct Walk { int p, N, mode, k; };
    auto walk_load = [&](Walk& w) {
        w.N = d.ph_N[ph_idx(d, w.p, i)];
        w.mode = d.ph_mode[ph_idx(d, w.p, i)];
    };
    auto walk_init = [&](Walk& w) {
        w.p = np - 1; w.N = 0; w.mode = 1; w.k = 0;
        if (enabled && np > 0) { walk_load(w); w.k = w.N - 1; }
    };
    auto walk_step = [&](Walk& w) {   // to the previous knot
        if (--w.k < 0 && w.p > 0) { --w.p; walk_load(w); w.k = w.N - 1; }
    };
    auto prefetch = [&](Walk& w, int it_next, double* dst) {
        if (enabled && it_next < nk) {
            const int len = knot_rec_len(w.p < nwb, w.k == w.N - 1);
            const double* src = lin_i + (size_t)(w.p * KSTRIDE + w.k) * REC;
            for (int q = l; 2 * q < len; q += 16) __pipeline_memcpy_async(dst + 2 * q, src + 2 * q, 16);
            walk_step(w);
        }
        __pipeline_commit();
    };

    while (true) {
        // ---- one attempt -----------------------------------------------------------------
        double Hn[14];
#pragma unroll
        for (int q = 0; q < 14; ++q) Hn[q] = 0.0;
        scr[S_G + l] = 0.0;
        double dV = 0.0;
        bool failed = false;
        Walk wk, wp;
        walk_init(wk); walk_init(wp);
        int& p = wk.p; int& Nn = wk.N; int& mode = wk.mode;
        __syncwarp();
        if (ktb > 0) prefetch(wp, 0, bufs);
        for (int it = 0; it < ktb; ++it) {
            double* buf = bufs + (size_t)(it & 1) * REC;
            double* nbuf = bufs + (size_t)((it & 1) ^ 1) * REC;
            __pipeline_wait_prior(0);
            __syncwarp();
            const int alive = (!done && !failed) ? 1 : 0;
            if (!__any_sync(0xffffffffu, alive)) break;
            prefetch(wp, it + 1, nbuf);

            // ---- which kind of knot is this for my instance? ----
            const int kk = wk.p * KSTRIDE + wk.k;   // knot slot of this step
            const int off = wk.p * KSTRIDE;
            int ty = 0;  // 0 idle, 1 WB step, 2 RB step, 3 WB terminal, 4 RB terminal
            if (alive && it < nk) {
                const bool wb = p < nwb, term = (kk - off == Nn - 1);
                ty = wb ? (term ? 3 : 1) : (term ? 4 : 2);
            }
            const int ty_o = __shfl_xor_sync(0xffffffffu, ty, 16);
#define RBUF(e) buf[(e)]

            // ============ terminal knots: value reset across the phase boundary ============
            if (ty == 3 || ty_o == 3 || ty == 4 || ty_o == 4) {
                const bool on = (ty == 3 || ty == 4);
                const bool last = on && (p == np - 1);
                const bool wbp = (ty == 3);
                const bool next_rb = on && !last && (p + 1 >= nwb);
                // (a) WB <- RB: embed the 6-dim value into 14 dims through P^T (.) P  (MultiPhaseDDP.cpp:330-331)
                const bool embed = wbp && next_rb;
                double g_old = 0.0;
                if (embed) {
                    if (l < 6) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) scr[S_MQ + l * MQ_LD + q] = Hn[q];
                    }
                    const int pl = (l < 3) ? l : (l >= 7 && l < 10) ? l - 4 : -1;
                    g_old = (pl >= 0) ? scr[S_G + pl] : 0.0;
                }
                __syncwarp();
                if (embed) {
                    const int pl = (l < 3) ? l : (l >= 7 && l < 10) ? l - 4 : -1;
#pragma unroll
                    for (int q = 0; q < 14; ++q) {
                        const int pq = (q < 3) ? q : (q >= 7 && q < 10) ? q - 4 : -1;
                        Hn[q] = (pl >= 0 && pq >= 0) ? scr[S_MQ + pl * MQ_LD + pq] : 0.0;
                    }
                }
                __syncwarp();
                if (embed && l < 14) scr[S_G + l] = g_old;
                __syncwarp();
                // (b) WB phase end: G = Px^T G', H = Px^T H' Px  (MultiPhaseDDP.cpp:320-326), Px = [I7 0; 0 Pb]
                const bool px = wbp && !last;
                if (px && l < 14) {
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) {
                        double acc = (jc < 7) ? Hn[jc] : 0.0;
                        if (jc >= 2) {
#pragma unroll
                            for (int s = 0; s < 7; ++s) acc += Hn[7 + s] * RBUF(R_AB(s, jc - 2));
                        }
                        scr[S_MQ + l * MQ_LD + jc] = acc;
                    }
                }
                __syncwarp();
                double g_new = 0.0;
                if (px && l < 14) {
                    double pb[7];
#pragma unroll
                    for (int s = 0; s < 7; ++s) pb[s] = (l >= 2) ? RBUF(R_AB(s, l - 2)) : 0.0;
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) {
                        double acc = (l < 7) ? scr[S_MQ + l * MQ_LD + jc] : 0.0;
                        if (l >= 2) {
#pragma unroll
                            for (int s = 0; s < 7; ++s) acc += pb[s] * scr[S_MQ + (7 + s) * MQ_LD + jc];
                        }
                        Hn[jc] = acc;
                    }
                    g_new = (l < 7) ? scr[S_G + l] : 0.0;
                    if (l >= 2) {
#pragma unroll
                        for (int s = 0; s < 7; ++s) g_new += pb[s] * scr[S_G + 7 + s];
                    }
                }
                __syncwarp();
                if (px && l < 14) scr[S_G + l] = g_new;
                __syncwarp();
                // (c) terminal cost: G = Phix + Gnext, H = Phixx + Hnext  (SinglePhase.cpp:190-191)
                if (on) {
                    if (last) {
#pragma unroll
                        for (int q = 0; q < 14; ++q) Hn[q] = 0.0;
                    }
                    const int nx = wbp ? 14 : 6;
                    if (l < nx) {
                        double gv = last ? 0.0 : scr[S_G + l];
                        if (wbp) {
#pragma unroll
                            for (int jc = 0; jc < 14; ++jc) {
                                double ph = (jc == l) ? RBUF(R_PHIXXD + l) : 0.0;
                                if (l >= 1 && l <= 6 && jc >= 1 && jc <= 6) ph = ph + RBUF(R_PHIXXB + (l - 1) + 6 * (jc - 1));
                                Hn[jc] = ph + Hn[jc];
                            }
                            gv = RBUF(R_PHIX + l) + gv;
                        } else {
#pragma unroll
                            for (int jc = 0; jc < 6; ++jc) Hn[jc] = ((jc == l) ? c.rbQf[l] : 0.0) + Hn[jc];
                            gv = RBUF(R_RPHIX + l) + gv;
                        }
                        scr[S_G + l] = gv;
                        gv_lane[(size_t)kk * gv_knot] = gv;
                    }
                }
                __syncwarp();
            }

            // ============ whole-body step ============
            if (ty == 1 || ty_o == 1) {
                const bool on = (ty == 1);
                const double dt = dtw;
                const int r0 = (mode == 1) ? 2 : 0;
                if (on && l < 14) {
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) {   // M = H A
                        double acc;
                        if (jc < 2) acc = Hn[jc];
                        else {
                            acc = (jc < 7) ? Hn[jc] : Hn[jc - 7] * dt;
#pragma unroll
                            for (int s = 0; s < 7; ++s) acc += Hn[7 + s] * RBUF(R_AB(s, jc - 2));
                        }
                        scr[S_MQ + l * MQ_LD + jc] = acc;
                    }
                    if (l >= 7) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {   // N = H B (rows 7..13 are all B^T N needs)
                            double acc = Hn[7] * RBUF(R_BB(0, cc));
#pragma unroll
                            for (int s = 1; s < 7; ++s) acc += Hn[7 + s] * RBUF(R_BB(s, cc));
                            scr[S_N + (l - 7) * 4 + cc] = acc;
                        }
                    }
                }
                __syncwarp();
                double qux[4] = {0, 0, 0, 0}, qx = 0.0;
                if (on) {
                    const double L00 = RBUF(R_LYY), L01 = RBUF(R_LYY + 1), L11 = RBUF(R_LYY + 2);
                    const double ly0 = RBUF(R_LY), ly1 = RBUF(R_LY + 1);
                    if (l < 14) {
                        double ab[7];
#pragma unroll
                        for (int s = 0; s < 7; ++s) ab[s] = (l >= 2) ? RBUF(R_AB(s, l - 2)) : 0.0;
                        const double c0 = (l >= 2) ? RBUF(R_CB(0, l - 2)) : 0.0, c1 = (l >= 2) ? RBUF(R_CB(1, l - 2)) : 0.0;
                        const double w0 = c0 * L00 + c1 * L01, w1 = c0 * L01 + c1 * L11;   // row l of C^T lyy
                        const double lxxd = (l >= 3 && l < 7) ? RBUF(R_LXXQ + (l - 3)) : (2 * dt) * c.wbQ[l];
#pragma unroll
                        for (int jc = 0; jc < 14; ++jc) {   // Qxx row l = (lxx + C^T lyy C) + A^T M
                            double q;
                            if (l < 2) q = scr[S_MQ + l * MQ_LD + jc];
                            else {
                                q = (l < 7) ? scr[S_MQ + l * MQ_LD + jc] : scr[S_MQ + (l - 7) * MQ_LD + jc] * dt;
#pragma unroll
                                for (int s = 0; s < 7; ++s) q += ab[s] * scr[S_MQ + (7 + s) * MQ_LD + jc];
                            }
                            double lc = (jc == l) ? lxxd : 0.0;
                            if (l >= 2 && jc >= 2) lc = lc + (w0 * RBUF(R_CB(0, jc - 2)) + w1 * RBUF(R_CB(1, jc - 2)));
                            Hn[jc] = lc + q;
                            if (jc == l) Hn[jc] += 1.0 * reg;
                        }
                        // Qx = (lx + A^T G) + C^T ly
                        double ag;
                        if (l < 2) ag = scr[S_G + l];
                        else {
                            ag = (l < 7) ? scr[S_G + l] : scr[S_G + l - 7] * dt;
#pragma unroll
                            for (int s = 0; s < 7; ++s) ag += ab[s] * scr[S_G + 7 + s];
                        }
                        qx = (RBUF(R_LX + l) + ag) + (c0 * ly0 + c1 * ly1);
                        // Qux column l = (0 + D^T lyy C) + B^T M
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            double acc = RBUF(R_BB(0, cc)) * scr[S_MQ + 7 * MQ_LD + l];
#pragma unroll
                            for (int s = 1; s < 7; ++s) acc += RBUF(R_BB(s, cc)) * scr[S_MQ + (7 + s) * MQ_LD + l];
                            const double d0 = RBUF(R_DB(0, cc)), d1 = RBUF(R_DB(1, cc));
                            const double v0 = d0 * L00 + d1 * L01, v1 = d0 * L01 + d1 * L11;
                            qux[cc] = (0.0 + (v0 * c0 + v1 * c1)) + acc;
                        }
                    }
                    {   // Quu entry (row l&3, col l>>2) = (luu + D^T lyy D) + B^T N
                        const int rr = l & 3, cc = l >> 2;
                        double acc = RBUF(R_BB(0, rr)) * scr[S_N + 0 * 4 + cc];
#pragma unroll
                        for (int s = 1; s < 7; ++s) acc += RBUF(R_BB(s, rr)) * scr[S_N + s * 4 + cc];
                        const double d0 = RBUF(R_DB(0, rr)), d1 = RBUF(R_DB(1, rr));
                        const double v0 = d0 * L00 + d1 * L01, v1 = d0 * L01 + d1 * L11;
                        double q = ((rr == cc ? RBUF(R_LUU + rr) : 0.0) + (v0 * RBUF(R_DB(0, cc)) + v1 * RBUF(R_DB(1, cc)))) + acc;
                        if (rr == cc) q += 1.0 * reg;
                        scr[S_QUU + rr + 4 * cc] = q;
                    }
                    if (l < 4) {   // Qu = (lu + B^T G) + D^T ly
                        double acc = RBUF(R_BB(0, l)) * scr[S_G + 7];
#pragma unroll
                        for (int s = 1; s < 7; ++s) acc += RBUF(R_BB(s, l)) * scr[S_G + 7 + s];
                        scr[S_QU + l] = (RBUF(R_LU + l) + acc) + (RBUF(R_DB(0, l)) * ly0 + RBUF(R_DB(1, l)) * ly1);
                    }
                    (void)r0;
                }
                __syncwarp();   // everyone is done reading M
                if (on && l < 14) {
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) scr[S_MQ + l * MQ_LD + jc] = Hn[jc];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_QUX + cc * 14 + l] = qux[cc];
                }
                __syncwarp();
                if (on && l < 14) {
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) Hn[jc] = (Hn[jc] + scr[S_MQ + jc * MQ_LD + l]) / 2;   // Qxx = (Qxx + Qxx^T)/2
                }
                double kcol[4], du[4], gout = 0.0, dvs = 0.0;
                bool ok = true;
                __syncwarp();
                ok = value_update<14>(on, l, scr, qux, qx, Hn, opt.quu_shift, kcol, du, gout, dvs);
                __syncwarp();
                if (on) {
                    if (!ok) failed = true;
                    else {
                        dV += dvs;
                        if (l < 14) {
                            scr[S_G + l] = gout;
                            gv_lane[(size_t)kk * gv_knot] = gout;
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) kf_lane[(size_t)kk * kf_knot + (size_t)cc * es] = kcol[cc];
                        }
                        if (l < 4) du_lane[(size_t)kk * du_knot] = du[l];
                    }
                }
                __syncwarp();
            }

            // ============ rigid-body step ============
            // A = I + dt*[0 I3; 0 0] + e5 (a0, a1, 0, 0, 0, 0), B has two non-zeros per column (rows 3 + (c&1) and 5):
            // only exact-zero terms of the reference's dense 6x6 products are dropped, the remaining terms keep the
            // ascending inner-index order.
            if (ty == 2 || ty_o == 2) {
                const bool on = (ty == 2);
                const double dt = dtf;
                const double a0 = on ? RBUF(R_RA) : 0.0, a1 = on ? RBUF(R_RA + 1) : 0.0;
                double bt[4], b5[4];   // B(3 + (c&1), c), B(5, c)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) { bt[cc] = on ? RBUF(R_RB + 2 * cc) : 0.0; b5[cc] = on ? RBUF(R_RB + 2 * cc + 1) : 0.0; }
                if (on && l < 6) {
                    // M = H A (row l)
                    scr[S_MQ + l * MQ_LD + 0] = Hn[0] + Hn[5] * a0;
                    scr[S_MQ + l * MQ_LD + 1] = Hn[1] + Hn[5] * a1;
                    scr[S_MQ + l * MQ_LD + 2] = Hn[2];
                    scr[S_MQ + l * MQ_LD + 3] = Hn[0] * dt + Hn[3];
                    scr[S_MQ + l * MQ_LD + 4] = Hn[1] * dt + Hn[4];
                    scr[S_MQ + l * MQ_LD + 5] = Hn[2] * dt + Hn[5];
                    // N = H B (row l)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_N + l * 4 + cc] = Hn[3 + (cc & 1)] * bt[cc] + Hn[5] * b5[cc];
                }
                __syncwarp();
                double qux[4] = {0, 0, 0, 0}, qx = 0.0;
                if (on) {
                    if (l < 6) {
                        const double acol = (l == 0) ? a0 : a1;
#pragma unroll
                        for (int jc = 0; jc < 6; ++jc) {   // Qxx row l = lxx + A^T M
                            double q;
                            if (l < 2) q = scr[S_MQ + l * MQ_LD + jc] + acol * scr[S_MQ + 5 * MQ_LD + jc];
                            else if (l == 2) q = scr[S_MQ + 2 * MQ_LD + jc];
                            else q = dt * scr[S_MQ + (l - 3) * MQ_LD + jc] + scr[S_MQ + l * MQ_LD + jc];
                            Hn[jc] = ((jc == l) ? (2 * dt) * c.rbQ[l] : 0.0) + q;
                            if (jc == l) Hn[jc] += 1.0 * reg;
                        }
                        double ag;
                        if (l < 2) ag = scr[S_G + l] + acol * scr[S_G + 5];
                        else if (l == 2) ag = scr[S_G + 2];
                        else ag = dt * scr[S_G + l - 3] + scr[S_G + l];
                        qx = RBUF(R_RLX + l) + ag;
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc)   // Qux column l = B^T M
                            qux[cc] = bt[cc] * scr[S_MQ + (3 + (cc & 1)) * MQ_LD + l] + b5[cc] * scr[S_MQ + 5 * MQ_LD + l];
                    }
                    {
                        const int rr = l & 3, cc = l >> 2;
                        const double acc = bt[rr] * scr[S_N + (3 + (rr & 1)) * 4 + cc] + b5[rr] * scr[S_N + 5 * 4 + cc];
                        double q = ((rr == cc) ? (2 * dt) * c.rbR[mode - 1][rr] : 0.0) + acc;
                        if (rr == cc) q += 1.0 * reg;
                        scr[S_QUU + rr + 4 * cc] = q;
                    }
                    if (l < 4) scr[S_QU + l] = RBUF(R_RLU + l) + (bt[l] * scr[S_G + 3 + (l & 1)] + b5[l] * scr[S_G + 5]);
                }
                __syncwarp();
                if (on && l < 6) {
#pragma unroll
                    for (int jc = 0; jc < 6; ++jc) scr[S_MQ + l * MQ_LD + jc] = Hn[jc];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_QUX + cc * 14 + l] = qux[cc];
                }
                __syncwarp();
                if (on && l < 6) {
#pragma unroll
                    for (int jc = 0; jc < 6; ++jc) Hn[jc] = (Hn[jc] + scr[S_MQ + jc * MQ_LD + l]) / 2;
                }
                double kcol[4], du[4], gout = 0.0, dvs = 0.0;
                __syncwarp();
                bool ok = value_update<6>(on, l, scr, qux, qx, Hn, opt.quu_shift, kcol, du, gout, dvs);
                if (on) {
                    if (!ok) failed = true;
                    else {
                        dV += dvs;
                        if (l < 6) {
                            scr[S_G + l] = gout;
                            gv_lane[(size_t)kk * gv_knot] = gout;
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) kf_lane[(size_t)kk * kf_knot + (size_t)cc * es] = kcol[cc];
                        }
                        if (l < 4) du_lane[(size_t)kk * du_knot] = du[l];
                    }
                }
                __syncwarp();
            }
#undef RBUF
            if (enabled && it < nk) walk_step(wk);
        }
        __pipeline_wait_prior(0);
        __syncwarp();

        // ---- outcome of the attempt (MultiPhaseDDP.cpp:209-241) ----
        int retry = 0;
        if (!done) {
            if (single_attempt) {
                if (l == 0) { d.si[(size_t)SI_BWS_OK * es + i] = failed ? 0 : 1; if (!failed) d.sd[(size_t)SD_DV * es + i] = dV; }
                done = true;
            } else if (failed) {
                reg = fmax(reg * opt.update_regularization, 1e-03);
                ++n_bws;
                if (reg > 1000) {
                    if (l == 0) { d.si[(size_t)SI_STATUS * es + i] = STATUS_REG_ABORT; d.si[(size_t)SI_DDP_ACTIVE * es + i] = 0; d.si[(size_t)SI_N_BWS * es + i] = n_bws; }
                    done = true;
                } else retry = 1;
            } else {
                if (l == 0) {
                    d.sd[(size_t)SD_DV * es + i] = dV;
                    d.sd[(size_t)SD_REG_USED * es + i] = reg;
                    double r2 = reg / 20;
                    if (r2 < 1e-06) r2 = 0;
                    d.sd[(size_t)SD_REG * es + i] = r2;
                    d.si[(size_t)SI_N_BWS * es + i] = n_bws;
                    d.si[(size_t)SI_BWS_OK * es + i] = 1;
                }
                done = true;
            }
        }
        if (!__any_sync(0xffffffffu, retry)) break;
    }
}

void launch_bws(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int single_attempt, cudaStream_t s) {
    cudaFuncSetAttribute(k_bws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWS_SMEM);
    k_bws<<<(d.n + BWS_INST - 1) / BWS_INST, BWS_THREADS, BWS_SMEM, s>>>(d, c, opt, mask_mode, single_attempt);
}

}  // namespace mhpc
