// TEST INFRASTRUCTURE (oracle) - not part of the product. See oracle/README.md.
//
// Minimal fixed-size column-major matrix helper used by the CPU restatement of the
// reference solver.  The reference uses Eigen3 (system dependency, version unpinned,
// absent from this image: /root/reference/Common/header/MHPC_CPPTypes.h:12-16).  The
// operations below restate the *mathematical* meaning of the Eigen expressions at the
// reference call sites with a fixed, documented evaluation order:
//   * products associate left-to-right exactly as written in the reference source;
//   * every inner product sums its terms in ascending index order, starting from the
//     first term (no zero seed), one rounding per operation unless the compiler
//     contracts a*b+c (the reference is built with -O3 -march=native, i.e.
//     -ffp-contract=fast; see oracle/Makefile for the two builds).
// Storage is column-major like Eigen's default so that the CasADi CCS -> dense scatter
// (row + nrow*col, CasadiGen/source/CasadiGen.cpp:62) lands where Eigen would put it.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>

namespace orc {

template <int R, int C>
struct Mat {
    double d[R * C > 0 ? R * C : 1];
    static constexpr int rows = R, cols = C;
    double& operator()(int r, int c) { return d[r + R * c]; }
    double operator()(int r, int c) const { return d[r + R * c]; }
    double& operator[](int i) { return d[i]; }
    double operator[](int i) const { return d[i]; }
    void setZero() { std::memset(d, 0, sizeof(d)); }
    void setIdentity() {
        setZero();
        for (int i = 0; i < (R < C ? R : C); ++i) (*this)(i, i) = 1.0;
    }
    double* data() { return d; }
    const double* data() const { return d; }
};
template <int N>
using Vec = Mat<N, 1>;

template <int R, int C>
inline Mat<R, C> zeros() { Mat<R, C> m; m.setZero(); return m; }

template <int R, int C>
inline Mat<R, C> operator+(const Mat<R, C>& a, const Mat<R, C>& b) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] + b.d[i];
    return o;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& a, const Mat<R, C>& b) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] - b.d[i];
    return o;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& a) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.d[i] = -a.d[i];
    return o;
}
template <int R, int C>
inline Mat<R, C> operator*(const Mat<R, C>& a, double s) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] * s;
    return o;
}
template <int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C>& a) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.d[i] = s * a.d[i];
    return o;
}
template <int R, int C>
inline Mat<R, C> operator/(const Mat<R, C>& a, double s) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] / s;
    return o;
}
template <int R, int C>
inline Mat<R, C>& operator+=(Mat<R, C>& a, const Mat<R, C>& b) {
    for (int i = 0; i < R * C; ++i) a.d[i] += b.d[i];
    return a;
}

// a * b, inner index ascending
template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
    Mat<R, C> o;
    for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r) {
            double s = a(r, 0) * b(0, c);
            for (int k = 1; k < K; ++k) s += a(r, k) * b(k, c);
            o(r, c) = s;
        }
    return o;
}

template <int R, int C>
inline Mat<C, R> T(const Mat<R, C>& a) {
    Mat<C, R> o;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) o(c, r) = a(r, c);
    return o;
}

// a^T * b without forming the transpose (same term order as T(a) * b)
template <int K, int R, int C>
inline Mat<R, C> tmul(const Mat<K, R>& a, const Mat<K, C>& b) {
    Mat<R, C> o;
    for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r) {
            double s = a(0, r) * b(0, c);
            for (int k = 1; k < K; ++k) s += a(k, r) * b(k, c);
            o(r, c) = s;
        }
    return o;
}

template <int N>
inline double dot(const Vec<N>& a, const Vec<N>& b) {
    double s = a[0] * b[0];
    for (int i = 1; i < N; ++i) s += a[i] * b[i];
    return s;
}

}  // namespace orc
