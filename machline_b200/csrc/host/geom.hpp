// Small fixed-size vector helpers for the host-side setup code.
// Arithmetic order follows the reference's common/math.f90 (cross :80, inner :96, outer :120,
// det3 :137) so that rounding matches a gfortran -O2 -fdefault-real-8 build (no FMA contraction:
// this directory is compiled with -ffp-contract=off).
#pragma once
#include <array>
#include <cmath>
#include <vector>
#include <cstddef>

namespace mlh {

using V3 = std::array<double, 3>;
using M33 = std::array<std::array<double, 3>, 3>;  // [row][col]
typedef __float128 quad;

inline V3 operator+(const V3& a, const V3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline V3 operator-(const V3& a, const V3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline V3 operator-(const V3& a) { return {-a[0], -a[1], -a[2]}; }
inline V3 operator*(double s, const V3& a) { return {s * a[0], s * a[1], s * a[2]}; }
inline V3 operator*(const V3& a, double s) { return {a[0] * s, a[1] * s, a[2] * s}; }
inline V3 operator/(const V3& a, double s) { return {a[0] / s, a[1] / s, a[2] / s}; }

// math.f90:96
inline double inner(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// math.f90:80
inline V3 cross(const V3& a, const V3& b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}

// gfortran's NORM2 intrinsic (libgfortran norm2_r8 and the inline expansion both use the scaled
// sum-of-squares recurrence below, not sqrt(sum(x**2))).
template <typename T>
inline T norm2_gf(const T* x, int n) {
    T result = 0, scale = 1;
    for (int i = 0; i < n; ++i) {
        if (x[i] != 0) {
            T ax = x[i] < 0 ? -x[i] : x[i];
            if (scale < ax) {
                T val = scale / ax;
                result = 1 + result * val * val;
                scale = ax;
            } else {
                T val = ax / scale;
                result += val * val;
            }
        }
    }
    return scale * std::sqrt(result);
}
quad sqrt_quad(quad x);  // geom.cpp (libquadmath)
template <>
inline quad norm2_gf<quad>(const quad* x, int n) {
    quad result = 0, scale = 1;
    for (int i = 0; i < n; ++i) {
        if (x[i] != 0) {
            quad ax = x[i] < 0 ? -x[i] : x[i];
            if (scale < ax) {
                quad val = scale / ax;
                result = 1 + result * val * val;
                scale = ax;
            } else {
                quad val = ax / scale;
                result += val * val;
            }
        }
    }
    return scale * sqrt_quad(result);
}
inline double norm2(const V3& a) { return norm2_gf<double>(a.data(), 3); }
inline double norm2_2(double a, double b) {
    double v[2] = {a, b};
    return norm2_gf<double>(v, 2);
}
// math.f90:71 dist = norm2(a-b)
inline double dist(const V3& a, const V3& b) { return norm2(a - b); }

// helpers.f90:41-54 (plane = 1..3, index of the component normal to the plane)
inline V3 mirror_across_plane(const V3& v, int plane) {
    V3 m = v;
    m[plane - 1] = -v[plane - 1];
    return m;
}

// matmul(A(3,3), v(3)) -- gfortran accumulates over the contracted index in ascending order.
inline V3 matvec(const M33& A, const V3& v) {
    V3 c;
    for (int i = 0; i < 3; ++i) c[i] = A[i][0] * v[0] + A[i][1] * v[1] + A[i][2] * v[2];
    return c;
}
inline M33 transpose(const M33& A) {
    M33 T;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i][j] = A[j][i];
    return T;
}
inline M33 matmul(const M33& A, const M33& B) {
    M33 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
    return C;
}
// math.f90:137
inline double det3(const M33& a) {
    double c = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]);
    c = c - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]);
    c = c + a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]);
    return c;
}

// linalg.f90:9-115 matinv (Gauss-Jordan with row scaling and partial pivoting on a permutation
// vector).  a and ai are n x n, row-major here ([i*n+j] == Fortran a(i,j)).
void matinv(int n, const double* a, double* ai);
inline M33 matinv3(const M33& A) {
    double a[9], ai[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[3 * i + j] = A[i][j];
    matinv(3, a, ai);
    M33 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = ai[3 * i + j];
    return R;
}

inline double sign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }

}  // namespace mlh
