// TEST INFRASTRUCTURE (oracle) - not part of the product.
//
// Restatement of the two Eigen3 algorithms that carry hot-path arithmetic in the
// reference and are NOT under /root/reference (Eigen3 is an un-vendored, un-pinned
// system dependency: Readme.md:2, Common/header/MHPC_CPPTypes.h:12-16,75-76):
//
//  * `Quu.inverse()` for a fixed 4x4 matrix  (call sites
//    Common/header/MHPC_CompoundTypes.h:133,141).  Eigen's published generic algorithm
//    (Eigen/src/LU/InverseImpl.h, compute_inverse_size4, scalar path): adjugate from
//    3x3 cofactors written as sums of three "a*(b*c-d*e)" terms, divided by the
//    determinant obtained as col(0) . row(0) of the adjugate.  (Eigen also ships SSE/AVX
//    block-partitioned kernels for 4x4; which one a given build uses depends on the
//    Eigen version and flags.  Results agree to rounding; parity is therefore toleranced.)
//
//  * `Eigen::LDLT<MatrixXd>::compute(M).isPositive()`  (call site
//    HSDDPSolver/source/SinglePhase.cpp:202-205).  Eigen's published algorithm
//    (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked, 3.3/3.4): symmetric
//    diagonal pivoting on the largest |diagonal| entry, in-place LDL^T on the lower
//    triangle, and a sign tracker: start ZeroSign; a positive pivot moves Zero ->
//    PositiveSemiDef, a negative pivot moves Zero -> NegativeSemiDef, and a pivot of the
//    opposite sign makes it Indefinite.  isPositive() == (sign is PositiveSemiDef or Zero).
#pragma once
#include "smallmat.h"

namespace orc {

inline double det3_helper(const Mat<4, 4>& m, int i1, int i2, int i3, int j1, int j2, int j3) {
    return m(i1, j1) * (m(i2, j2) * m(i3, j3) - m(i2, j3) * m(i3, j2));
}
inline double cofactor4(const Mat<4, 4>& m, int i, int j) {
    const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4;
    const int j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
    return det3_helper(m, i1, i2, i3, j1, j2, j3) + det3_helper(m, i2, i3, i1, j1, j2, j3) +
           det3_helper(m, i3, i1, i2, j1, j2, j3);
}
inline Mat<4, 4> inverse4(const Mat<4, 4>& m) {
    Mat<4, 4> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const double c = cofactor4(m, i, j);
            r(j, i) = ((i + j) & 1) ? -c : c;
        }
    double det = m(0, 0) * r(0, 0);
    for (int i = 1; i < 4; ++i) det += m(i, 0) * r(0, i);
    for (int i = 0; i < 16; ++i) r.d[i] /= det;
    return r;
}

// Returns true iff the pivoted LDL^T sign tracker ends PositiveSemiDef or ZeroSign.
template <int N>
inline bool ldlt_is_positive(Mat<N, N> a) {
    enum { Zero, Pos, Neg, Indef } sign = Zero;
    bool found_zero_pivot = false;
    double temp[N];
    for (int k = 0; k < N; ++k) {
        int p = k;
        double big = std::fabs(a(k, k));
        for (int i = k + 1; i < N; ++i)
            if (std::fabs(a(i, i)) > big) { big = std::fabs(a(i, i)); p = i; }
        if (p != k) {  // symmetric swap acting on the lower triangle
            const int s = N - p - 1;
            for (int j = 0; j < k; ++j) { double t = a(k, j); a(k, j) = a(p, j); a(p, j) = t; }
            for (int i = 0; i < s; ++i) { double t = a(p + 1 + i, k); a(p + 1 + i, k) = a(p + 1 + i, p); a(p + 1 + i, p) = t; }
            { double t = a(k, k); a(k, k) = a(p, p); a(p, p) = t; }
            for (int i = k + 1; i < p; ++i) { double t = a(i, k); a(i, k) = a(p, i); a(p, i) = t; }
        }
        const int rs = N - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = a(j, j) * a(k, j);
            double s = a(k, 0) * temp[0];
            for (int j = 1; j < k; ++j) s += a(k, j) * temp[j];
            a(k, k) -= s;
            for (int i = 0; i < rs; ++i) {
                double t = a(k + 1 + i, 0) * temp[0];
                for (int j = 1; j < k; ++j) t += a(k + 1 + i, j) * temp[j];
                a(k + 1 + i, k) -= t;
            }
        }
        const double akk = a(k, k);
        const bool pivot_is_valid = std::fabs(akk) > 0.0;
        if (k == 0 && !pivot_is_valid) return true;  // whole diagonal is zero: ZeroSign
        if (rs > 0 && pivot_is_valid)
            for (int i = 0; i < rs; ++i) a(k + 1 + i, k) /= akk;
        if (found_zero_pivot && pivot_is_valid) { /* factorisation flagged failed; sign keeps tracking */ }
        else if (!pivot_is_valid) found_zero_pivot = true;
        if (sign == Pos) { if (akk < 0) sign = Indef; }
        else if (sign == Neg) { if (akk > 0) sign = Indef; }
        else if (sign == Zero) { if (akk > 0) sign = Pos; else if (akk < 0) sign = Neg; }
    }
    return sign == Pos || sign == Zero;
}

}  // namespace orc
