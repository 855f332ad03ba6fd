// dual.cuh — forward-mode dual numbers (value + ONE directional derivative) for the per-interval
// Jacobian kernels.  The reference differentiates the whole collocation loss with coloured
// ForwardDiff sweeps (lib/BoundaryValueDiffEqMIRK/src/mirk.jl:810-838); here every thread carries
// one seed direction through the s stages of ONE interval, so a column of [L_i R_i] costs one
// dual evaluation of Phi_i and never touches the other N-2 intervals.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mirk {

// functor trait (problems.cuh): does P declare `static constexpr bool bc_uses_derivative = true`?
template <class P, class = void> struct BcUsesDerivative { static constexpr bool value = false; };
template <class P> struct BcUsesDerivative<P, decltype((void)P::bc_uses_derivative)> { static constexpr bool value = P::bc_uses_derivative; };


struct Dual {
    double v, d;
    __host__ __device__ __forceinline__ Dual() : v(0.0), d(0.0) {}
    __host__ __device__ __forceinline__ Dual(double v_) : v(v_), d(0.0) {}
    __host__ __device__ __forceinline__ Dual(double v_, double d_) : v(v_), d(d_) {}
};

#define MIRK_OP __host__ __device__ __forceinline__
MIRK_OP Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
MIRK_OP Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
MIRK_OP Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
MIRK_OP Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
MIRK_OP Dual operator/(Dual a, Dual b) {
    const double q = a.v / b.v;
    return Dual(q, (a.d - q * b.d) / b.v);
}
MIRK_OP Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
MIRK_OP Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
MIRK_OP Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
MIRK_OP Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
MIRK_OP Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
MIRK_OP Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
MIRK_OP Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }
MIRK_OP Dual operator/(double a, Dual b) {
    const double q = a / b.v;
    return Dual(q, -q * b.d / b.v);
}
MIRK_OP Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
MIRK_OP Dual& operator-=(Dual& a, Dual b) { a.v -= b.v; a.d -= b.d; return a; }
MIRK_OP Dual& operator*=(Dual& a, Dual b) { a = a * b; return a; }
MIRK_OP Dual& operator+=(Dual& a, double b) { a.v += b; return a; }
MIRK_OP Dual& operator*=(Dual& a, double b) { a.v *= b; a.d *= b; return a; }

// elementary functions usable on both scalar types from templated RHS code: `using namespace mirk::fn;`
// MIRK_OUTLINE_ELEMENTARY (defined by translation units whose kernels are bounded by CODE SIZE — the whole-solve
// ensemble kernels): sin / cos / sincos / exp / log become out-of-line device functions, one copy per kernel instead
// of one (~100-400 instructions with its slow path) per call site.
#if defined(MIRK_OUTLINE_ELEMENTARY) && defined(__CUDACC__)
static __device__ __noinline__ double ol_sin(double x) { return ::sin(x); }
static __device__ __noinline__ double ol_cos(double x) { return ::cos(x); }
static __device__ __noinline__ double ol_exp(double x) { return ::exp(x); }
static __device__ __noinline__ double ol_log(double x) { return ::log(x); }
static __device__ __noinline__ void ol_sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
#define MIRK_EL_SIN(x) ol_sin(x)
#define MIRK_EL_COS(x) ol_cos(x)
#define MIRK_EL_EXP(x) ol_exp(x)
#define MIRK_EL_LOG(x) ol_log(x)
#define MIRK_EL_SINCOS(x, s, c) ol_sincos(x, s, c)
#else
#define MIRK_EL_SIN(x) ::sin(x)
#define MIRK_EL_COS(x) ::cos(x)
#define MIRK_EL_EXP(x) ::exp(x)
#define MIRK_EL_LOG(x) ::log(x)
#define MIRK_EL_SINCOS(x, s, c) ::sincos(x, s, c)
#endif

namespace fn {
#ifdef __CUDA_ARCH__
MIRK_OP double sin(double x) { return MIRK_EL_SIN(x); }
MIRK_OP double cos(double x) { return MIRK_EL_COS(x); }
MIRK_OP double exp(double x) { return MIRK_EL_EXP(x); }
MIRK_OP double log(double x) { return MIRK_EL_LOG(x); }
#else
MIRK_OP double sin(double x) { return ::sin(x); }
MIRK_OP double cos(double x) { return ::cos(x); }
MIRK_OP double exp(double x) { return ::exp(x); }
MIRK_OP double log(double x) { return ::log(x); }
#endif
MIRK_OP double sqrt(double x) { return ::sqrt(x); }
MIRK_OP double tanh(double x) { return ::tanh(x); }
MIRK_OP double square(double x) { return x * x; }
MIRK_OP double value(double x) { return x; }
MIRK_OP Dual sin(Dual a) {
    double s, c;
#ifdef __CUDA_ARCH__
    MIRK_EL_SINCOS(a.v, &s, &c);
#else
    s = ::sin(a.v); c = ::cos(a.v);
#endif
    return Dual(s, c * a.d);
}
MIRK_OP Dual cos(Dual a) {
    double s, c;
#ifdef __CUDA_ARCH__
    MIRK_EL_SINCOS(a.v, &s, &c);
#else
    s = ::sin(a.v); c = ::cos(a.v);
#endif
    return Dual(c, -s * a.d);
}
#ifdef __CUDA_ARCH__
MIRK_OP Dual exp(Dual a) { const double e = MIRK_EL_EXP(a.v); return Dual(e, e * a.d); }
MIRK_OP Dual log(Dual a) { return Dual(MIRK_EL_LOG(a.v), a.d / a.v); }
#else
MIRK_OP Dual exp(Dual a) { const double e = ::exp(a.v); return Dual(e, e * a.d); }
MIRK_OP Dual log(Dual a) { return Dual(::log(a.v), a.d / a.v); }
#endif
MIRK_OP Dual sqrt(Dual a) { const double r = ::sqrt(a.v); return Dual(r, a.d / (2.0 * r)); }
MIRK_OP Dual tanh(Dual a) { const double t = ::tanh(a.v); return Dual(t, (1.0 - t * t) * a.d); }
MIRK_OP Dual square(Dual a) { return Dual(a.v * a.v, 2.0 * a.v * a.d); }
MIRK_OP double value(Dual a) { return a.v; }
}  // namespace fn

}  // namespace mirk
