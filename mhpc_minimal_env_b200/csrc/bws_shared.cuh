// Pieces shared by the two lane mappings of the backward sweep (k_bws.cu: 8 lanes per instance, k_bws16.cu: 16 lanes per
// instance): per-instance scratch layout, the Eigen LDLT sign test, the value-function update and the per-row pieces of the
// phase-boundary value reset.  Every stored entry is produced by the same operation sequence whichever lane computes it.
#pragma once
#include "kernels.h"
#include "model.cuh"

namespace mhpc {

constexpr int MQ_LD = 15;
// per-instance scratch (doubles)
constexpr int S_G = 0;        // 16
constexpr int S_MQ = 16;      // 14 x 15 : rows 7..13 of M = H*A (at rows 0..6), later Qxx for the symmetrisation
constexpr int S_N = 226;      // 7 x 4  : rows 7..13 of H*B   (RB: 6 x 4)
constexpr int S_QUX = 254;    // 4 x 14
constexpr int S_QUU = 310;    // 16
constexpr int S_QU = 326;     // 4
// the three 4x4 work matrices of the value update live in the M/Qxx area, which is dead once Qxx has been symmetrised
constexpr int S_ADJ = S_MQ;         // 16 adjugate of Quu
constexpr int S_INV = S_MQ + 16;    // 16 Quu^-1
constexpr int S_QI = S_MQ + 32;     // 16 symmetrised Quu^-1
constexpr int SCR = 330;
// doubles per instance in shared memory: two record stages + scratch; = 8 (mod 16) so that the two instances of a
// half-warp sit in disjoint banks for every access pattern below
constexpr int BWS_REGION = 2 * REC + SCR + 14;
static_assert(BWS_REGION % 16 == 8 && SCR >= S_QU + 4, "bank layout of the per-instance shared-memory region");

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
#ifdef BWS_EXACT_DIV
            const double rk = 1.0 / akk;
#else
            const double rk = sx_rcp(akk);
#endif
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
// (`on` = this instance takes part; `rows` = this lane owns rows).  Quu^-1 is the closed-form cofactor inverse (what a
// fixed-size Matrix4d::inverse() computes): every lane forms TWO adjugate entries and does TWO divisions, the 16 entries
// are shared through scratch.
// In: Quu/Qu/Qux in scratch; per owned row (A = row m, B = row m+7 when TWO): Qux column qux[4], Qx value qx,
// symmetrised Qxx row in H[].  Out: H[] = new H row, G entry, K column per row; du; dV of the step.
template <int NX, bool TWO, int LPI>
__device__ __forceinline__ bool value_update(bool on, bool rows, int m, double* scr, const double quxa[4], const double quxb[4],
                                             double qxa, double qxb, double Ha[14], double Hb[14], double quu_shift,
                                             double ka[4], double kb[4], double du[4], double& Ga, double& Gb, double& dVstep) {
    bool pd = true;
#define QM(r, c) scr[S_QUU + (r) + 4 * (c)]
    if (on) {
        double tst[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) tst[r][c] = QM(r, c) - (r == c ? 1.0 * quu_shift : 0.0);
        pd = ldlt4_is_positive(tst);
        // adjugate entry r[jj][i] from the cofactor of (i, jj)
#pragma unroll
        for (int h = 0; h < 16 / LPI; ++h) {
            const int l = m + LPI * h;
            const int i = l & 3, jj = l >> 2;
            const int i1 = (i + 1) & 3, i2 = (i + 2) & 3, i3 = (i + 3) & 3, j1 = (jj + 1) & 3, j2 = (jj + 2) & 3, j3 = (jj + 3) & 3;
            const double cf = QM(i1, j1) * (QM(i2, j2) * QM(i3, j3) - QM(i2, j3) * QM(i3, j2)) +
                              QM(i2, j1) * (QM(i3, j2) * QM(i1, j3) - QM(i3, j3) * QM(i1, j2)) +
                              QM(i3, j1) * (QM(i1, j2) * QM(i2, j3) - QM(i1, j3) * QM(i2, j2));
            scr[S_ADJ + jj + 4 * i] = ((i + jj) & 1) ? -cf : cf;
        }
    }
    __syncwarp();
    if (on) {
        double det = QM(0, 0) * scr[S_ADJ + 0 + 4 * 0];
#pragma unroll
        for (int i = 1; i < 4; ++i) det += QM(i, 0) * scr[S_ADJ + 0 + 4 * i];
        // adj / det as adj * (1/det) with the Newton-refined hardware reciprocal (model.cuh: sx_rcp, <= 2 ulp): an IEEE
        // division is a ~127-cycle dependent sequence on the critical path of every knot.  -DBWS_EXACT_DIV restores it (a
        // zero numerator - every off-diagonal entry in the rigid-body phases - then takes adj * det, which has the same
        // value and sign, instead of the division's slow path).
        const double rdet = sx_rcp(det);
#pragma unroll
        for (int h = 0; h < 16 / LPI; ++h) {
            const int l = m + LPI * h;
            const double adj = scr[S_ADJ + l];
#ifdef BWS_EXACT_DIV
            scr[S_INV + l] = (adj == 0.0) ? adj * det : adj / det;
#else
            scr[S_INV + l] = adj * rdet;
#endif
        }
    }
#undef QM
    __syncwarp();
    if (on) {
#pragma unroll
        for (int h = 0; h < 16 / LPI; ++h) {
            const int l = m + LPI * h;
            const int r = l & 3, c = l >> 2;
            scr[S_QI + l] = (scr[S_INV + r + 4 * c] + scr[S_INV + c + 4 * r]) / 2;
        }
    }
    __syncwarp();
    if (!on || !pd) return !on ? true : false;
    double Qu[4], Ta[4], Tb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) Qu[c] = scr[S_QU + c];
    {   // Qi(r, c) at S_QI + r + 4c; the four columns c are independent chains
        double e4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double n0 = -scr[S_QI + c + 4 * 0], p0 = scr[S_QI + 0 + 4 * c];
            ka[c] = n0 * quxa[0]; Ta[c] = quxa[0] * p0; kb[c] = n0 * quxb[0]; Tb[c] = quxb[0] * p0; e4[c] = n0 * Qu[0];
        }
#pragma unroll
        for (int cc = 1; cc < 4; ++cc) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double nn = -scr[S_QI + c + 4 * cc], pp = scr[S_QI + cc + 4 * c];
                ka[c] += nn * quxa[cc];
                Ta[c] += quxa[cc] * pp;
                if (TWO) { kb[c] += nn * quxb[cc]; Tb[c] += quxb[cc] * pp; }
                e4[c] += nn * Qu[cc];
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) du[c] = e4[c];
    }
    double ga = Ta[0] * Qu[0], gb = Tb[0] * Qu[0];
#pragma unroll
    for (int c = 1; c < 4; ++c) { ga += Ta[c] * Qu[c]; gb += Tb[c] * Qu[c]; }
    Ga = qxa - ga;
    Gb = qxb - gb;
    if (rows) {
#pragma unroll
        for (int j0 = 0; j0 < NX; j0 += 7) {   // seven columns (up to fourteen chains) at a time
            double sa[7], sb[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                if (j0 + k < NX) { const double u0 = scr[S_QUX + j0 + k]; sa[k] = Ta[0] * u0; sb[k] = Tb[0] * u0; }
            }
#pragma unroll
            for (int c = 1; c < 4; ++c) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    if (j0 + k < NX) {
                        const double uc = scr[S_QUX + c * 14 + j0 + k];
                        sa[k] += Ta[c] * uc;
                        if (TWO) sb[k] += Tb[c] * uc;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                if (j0 + k < NX) {
                    Ha[j0 + k] = Ha[j0 + k] - sa[k];
                    if (TWO) Hb[j0 + k] = Hb[j0 + k] - sb[k];
                }
            }
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

// ---- per-row pieces of the phase-boundary value reset (called once per owned row; rare: one knot per phase) ----------
// WB index -> RB index of the projection P (MultiPhaseDDP.cpp:330-331): (x, z, pitch) and their rates
__device__ __forceinline__ int wb_to_rb(int l) { return (l < 3) ? l : (l >= 7 && l < 10) ? l - 4 : -1; }

__device__ __forceinline__ void embed_row(int l, const double* scr, double H[14]) {   // H <- row l of P^T H6 P
    const int pl = wb_to_rb(l);
#pragma unroll
    for (int q = 0; q < 14; ++q) {
        const int pq = wb_to_rb(q);
        H[q] = (pl >= 0 && pq >= 0) ? scr[S_MQ + pl * MQ_LD + pq] : 0.0;
    }
}
__device__ __forceinline__ void px_right_row(int l, const double* buf, double* scr, const double H[14]) {   // row l of H' Px
#pragma unroll
    for (int jc = 0; jc < 14; ++jc) {
        double acc = (jc < 7) ? H[jc] : 0.0;
        if (jc >= 2) {
#pragma unroll
            for (int s = 0; s < 7; ++s) acc += H[7 + s] * buf[R_AB(s, jc - 2)];
        }
        scr[S_MQ + l * MQ_LD + jc] = acc;
    }
}
__device__ __forceinline__ double px_left_row(int l, const double* buf, const double* scr, double H[14]) {   // row l of Px^T (.), returns G entry
    double pb[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) pb[s] = (l >= 2) ? buf[R_AB(s, l - 2)] : 0.0;
#pragma unroll
    for (int jc = 0; jc < 14; ++jc) {
        double acc = (l < 7) ? scr[S_MQ + l * MQ_LD + jc] : 0.0;
        if (l >= 2) {
#pragma unroll
            for (int s = 0; s < 7; ++s) acc += pb[s] * scr[S_MQ + (7 + s) * MQ_LD + jc];
        }
        H[jc] = acc;
    }
    double g_new = (l < 7) ? scr[S_G + l] : 0.0;
    if (l >= 2) {
#pragma unroll
        for (int s = 0; s < 7; ++s) g_new += pb[s] * scr[S_G + 7 + s];
    }
    return g_new;
}
__device__ __forceinline__ double wb_terminal_row(int l, const double* buf, double gnext, double H[14]) {   // Phixx + H, returns Phix + G
#pragma unroll
    for (int jc = 0; jc < 14; ++jc) {
        double ph = (jc == l) ? buf[R_PHIXXD + l] : 0.0;
        if (l >= 1 && l <= 6 && jc >= 1 && jc <= 6) ph = ph + buf[R_PHIXXB + (l - 1) + 6 * (jc - 1)];
        H[jc] = ph + H[jc];
    }
    return buf[R_PHIX + l] + gnext;
}

}  // namespace mhpc
