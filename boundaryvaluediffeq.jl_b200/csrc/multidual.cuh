// multidual.cuh — forward-mode dual number with M tangents: ONE evaluation of the collocation residual of an
// interval on DualN<2n> gives Phi_i, the stages and every column of [L_i R_i] at once, so elementary functions
// (sin, exp, ...) of the right-hand side are evaluated once per stage instead of once per column.  Used by the
// warp-per-trajectory ensemble kernel (ensemble_warp.cuh) where n is tiny (2n = 4 tangents live in registers).
// Same operator set as Dual (dual.cuh), so a functor templated on T needs no change.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "dual.cuh"

namespace mirk {

template <int M> struct DualN {
    double v;
    double d[M];
    __host__ __device__ __forceinline__ DualN() : v(0.0) {
#pragma unroll
        for (int j = 0; j < M; j++) d[j] = 0.0;
    }
    __host__ __device__ __forceinline__ DualN(double v_) : v(v_) {
#pragma unroll
        for (int j = 0; j < M; j++) d[j] = 0.0;
    }
    // value with the unit tangent e_seed
    __host__ __device__ __forceinline__ static DualN seed(double v_, int seed_) {
        DualN r(v_);
#pragma unroll
        for (int j = 0; j < M; j++) r.d[j] = j == seed_ ? 1.0 : 0.0;
        return r;
    }
};

#define MIRK_MD template <int M> __host__ __device__ __forceinline__
// r = f(a.v) with tangents fa * a.d
template <int M> __host__ __device__ __forceinline__ DualN<M> md_chain(const DualN<M>& a, double val, double fa) {
    DualN<M> r;
    r.v = val;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = fa * a.d[j];
    return r;
}
MIRK_MD DualN<M> operator+(const DualN<M>& a, const DualN<M>& b) {
    DualN<M> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = a.d[j] + b.d[j];
    return r;
}
MIRK_MD DualN<M> operator-(const DualN<M>& a, const DualN<M>& b) {
    DualN<M> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = a.d[j] - b.d[j];
    return r;
}
MIRK_MD DualN<M> operator-(const DualN<M>& a) {
    DualN<M> r;
    r.v = -a.v;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = -a.d[j];
    return r;
}
MIRK_MD DualN<M> operator*(const DualN<M>& a, const DualN<M>& b) {
    DualN<M> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = a.d[j] * b.v + a.v * b.d[j];
    return r;
}
MIRK_MD DualN<M> operator/(const DualN<M>& a, const DualN<M>& b) {
    DualN<M> r;
    const double q = a.v / b.v;
    r.v = q;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = (a.d[j] - q * b.d[j]) / b.v;
    return r;
}
MIRK_MD DualN<M> operator+(const DualN<M>& a, double b) { DualN<M> r = a; r.v = a.v + b; return r; }
MIRK_MD DualN<M> operator+(double a, const DualN<M>& b) { DualN<M> r = b; r.v = a + b.v; return r; }
MIRK_MD DualN<M> operator-(const DualN<M>& a, double b) { DualN<M> r = a; r.v = a.v - b; return r; }
MIRK_MD DualN<M> operator-(double a, const DualN<M>& b) { DualN<M> r = -b; r.v = a - b.v; return r; }
MIRK_MD DualN<M> operator*(const DualN<M>& a, double b) {
    DualN<M> r;
    r.v = a.v * b;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = a.d[j] * b;
    return r;
}
MIRK_MD DualN<M> operator*(double a, const DualN<M>& b) { return b * a; }
MIRK_MD DualN<M> operator/(const DualN<M>& a, double b) {
    DualN<M> r;
    r.v = a.v / b;
#pragma unroll
    for (int j = 0; j < M; j++) r.d[j] = a.d[j] / b;
    return r;
}
MIRK_MD DualN<M> operator/(double a, const DualN<M>& b) {
    const double q = a / b.v;
    return md_chain(b, q, -q / b.v);
}
MIRK_MD DualN<M>& operator+=(DualN<M>& a, const DualN<M>& b) { a = a + b; return a; }
MIRK_MD DualN<M>& operator-=(DualN<M>& a, const DualN<M>& b) { a = a - b; return a; }
MIRK_MD DualN<M>& operator*=(DualN<M>& a, const DualN<M>& b) { a = a * b; return a; }
MIRK_MD DualN<M>& operator+=(DualN<M>& a, double b) { a.v += b; return a; }
MIRK_MD DualN<M>& operator*=(DualN<M>& a, double b) { a = a * b; return a; }

namespace fn {
MIRK_MD DualN<M> sin(const DualN<M>& a) {
    double s, c;
#ifdef __CUDA_ARCH__
    MIRK_EL_SINCOS(a.v, &s, &c);
#else
    s = ::sin(a.v); c = ::cos(a.v);
#endif
    return md_chain(a, s, c);
}
MIRK_MD DualN<M> cos(const DualN<M>& a) {
    double s, c;
#ifdef __CUDA_ARCH__
    MIRK_EL_SINCOS(a.v, &s, &c);
#else
    s = ::sin(a.v); c = ::cos(a.v);
#endif
    return md_chain(a, c, -s);
}
#ifdef __CUDA_ARCH__
MIRK_MD DualN<M> exp(const DualN<M>& a) { const double e = MIRK_EL_EXP(a.v); return md_chain(a, e, e); }
MIRK_MD DualN<M> log(const DualN<M>& a) { return md_chain(a, MIRK_EL_LOG(a.v), 1.0 / a.v); }
#else
MIRK_MD DualN<M> exp(const DualN<M>& a) { const double e = ::exp(a.v); return md_chain(a, e, e); }
MIRK_MD DualN<M> log(const DualN<M>& a) { return md_chain(a, ::log(a.v), 1.0 / a.v); }
#endif
MIRK_MD DualN<M> sqrt(const DualN<M>& a) { const double r = ::sqrt(a.v); return md_chain(a, r, 0.5 / r); }
MIRK_MD DualN<M> tanh(const DualN<M>& a) { const double t = ::tanh(a.v); return md_chain(a, t, 1.0 - t * t); }
MIRK_MD DualN<M> square(const DualN<M>& a) { return md_chain(a, a.v * a.v, 2.0 * a.v); }
MIRK_MD double value(const DualN<M>& a) { return a.v; }
}  // namespace fn
#undef MIRK_MD

}  // namespace mirk
