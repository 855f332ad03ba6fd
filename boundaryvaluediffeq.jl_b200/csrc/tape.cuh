// tape.cuh — taped forward-mode differentiation for the per-interval Jacobian kernel.
//
// The reference obtains [L_i R_i] from coloured ForwardDiff sweeps of the collocation loss
// (lib/BoundaryValueDiffEqMIRK/src/mirk.jl:810-838).  A plain dual sweep per column recomputes the VALUE
// part of every stage — and with it every sin/cos/exp of the right-hand side — once per column: at
// n = 16 that is 32x the same 40 sincos per interval, >80 % of the instructions of the old kernel.
// Here the values are computed ONCE per interval (thread per interval, type RecVal) and every
// elementary-function result is appended to a small shared-memory tape; the 2n tangent sweeps (lane per
// column, type TapeDual) then replay the tape instead of calling the function again.  A user functor
// templated on T needs no change: `using namespace mirk::fn;` resolves sin/cos/exp/... on both types.
//
// Tape layout: entry e of the interval in slot s (0..31 within the warp) is the double2 at
// tape[e * 32 + s]; a warp records 32 intervals at once without bank conflicts and replays one
// interval with broadcast reads.  Calls beyond CAP entries are simply recomputed (correct, slower).
// The per-thread cursor lives in shared memory; after inlining and unrolling the compiler forwards it,
// so tape offsets become compile-time constants.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mirk {

#ifdef __CUDACC__

constexpr int kTapeSlots = 32;     // intervals recorded per CTA (by the lanes of its first warp)
constexpr int kTapeThreads = 128;  // most threads a tape kernel runs per CTA (each has its own cursor)

template <int CAP> struct Tape {
    static_assert(CAP >= 1, "tape needs at least one entry");
    __device__ __forceinline__ static double2* buf() {
        __shared__ __align__(16) double2 b[CAP * kTapeSlots];
        return b;
    }
    __device__ __forceinline__ static unsigned* cursor() {
        __shared__ unsigned c[kTapeThreads];
        return c + threadIdx.x;
    }
    __device__ __forceinline__ static unsigned* slot() {
        __shared__ unsigned s[kTapeThreads];
        return s + threadIdx.x;
    }
    __device__ __forceinline__ static void begin(int slot_) { *cursor() = 0u; *slot() = (unsigned)slot_; }
    // next entry, or nullptr when the tape is full
    __device__ __forceinline__ static double2* next() {
        const unsigned i = *cursor();
        *cursor() = i + 1u;
        return i < (unsigned)CAP ? buf() + i * kTapeSlots + *slot() : nullptr;
    }
};

// value-only scalar that records elementary-function results
template <int CAP> struct RecVal {
    double v;
    __host__ __device__ __forceinline__ RecVal() : v(0.0) {}
    __host__ __device__ __forceinline__ RecVal(double v_) : v(v_) {}
};
// value + one tangent; elementary functions come from the tape
template <int CAP> struct TapeDual {
    double v, d;
    __host__ __device__ __forceinline__ TapeDual() : v(0.0), d(0.0) {}
    __host__ __device__ __forceinline__ TapeDual(double v_) : v(v_), d(0.0) {}
    __host__ __device__ __forceinline__ TapeDual(double v_, double d_) : v(v_), d(d_) {}
};

#define MIRK_TOP template <int C> __host__ __device__ __forceinline__
// RecVal arithmetic
MIRK_TOP RecVal<C> operator+(RecVal<C> a, RecVal<C> b) { return RecVal<C>(a.v + b.v); }
MIRK_TOP RecVal<C> operator-(RecVal<C> a, RecVal<C> b) { return RecVal<C>(a.v - b.v); }
MIRK_TOP RecVal<C> operator*(RecVal<C> a, RecVal<C> b) { return RecVal<C>(a.v * b.v); }
MIRK_TOP RecVal<C> operator/(RecVal<C> a, RecVal<C> b) { return RecVal<C>(a.v / b.v); }
MIRK_TOP RecVal<C> operator-(RecVal<C> a) { return RecVal<C>(-a.v); }
MIRK_TOP RecVal<C> operator+(RecVal<C> a, double b) { return RecVal<C>(a.v + b); }
MIRK_TOP RecVal<C> operator+(double a, RecVal<C> b) { return RecVal<C>(a + b.v); }
MIRK_TOP RecVal<C> operator-(RecVal<C> a, double b) { return RecVal<C>(a.v - b); }
MIRK_TOP RecVal<C> operator-(double a, RecVal<C> b) { return RecVal<C>(a - b.v); }
MIRK_TOP RecVal<C> operator*(RecVal<C> a, double b) { return RecVal<C>(a.v * b); }
MIRK_TOP RecVal<C> operator*(double a, RecVal<C> b) { return RecVal<C>(a * b.v); }
MIRK_TOP RecVal<C> operator/(RecVal<C> a, double b) { return RecVal<C>(a.v / b); }
MIRK_TOP RecVal<C> operator/(double a, RecVal<C> b) { return RecVal<C>(a / b.v); }
MIRK_TOP RecVal<C>& operator+=(RecVal<C>& a, RecVal<C> b) { a.v += b.v; return a; }
MIRK_TOP RecVal<C>& operator-=(RecVal<C>& a, RecVal<C> b) { a.v -= b.v; return a; }
MIRK_TOP RecVal<C>& operator*=(RecVal<C>& a, RecVal<C> b) { a.v *= b.v; return a; }
MIRK_TOP RecVal<C>& operator+=(RecVal<C>& a, double b) { a.v += b; return a; }
MIRK_TOP RecVal<C>& operator*=(RecVal<C>& a, double b) { a.v *= b; return a; }
// TapeDual arithmetic (same rules as Dual in dual.cuh)
MIRK_TOP TapeDual<C> operator+(TapeDual<C> a, TapeDual<C> b) { return TapeDual<C>(a.v + b.v, a.d + b.d); }
MIRK_TOP TapeDual<C> operator-(TapeDual<C> a, TapeDual<C> b) { return TapeDual<C>(a.v - b.v, a.d - b.d); }
MIRK_TOP TapeDual<C> operator-(TapeDual<C> a) { return TapeDual<C>(-a.v, -a.d); }
MIRK_TOP TapeDual<C> operator*(TapeDual<C> a, TapeDual<C> b) { return TapeDual<C>(a.v * b.v, a.d * b.v + a.v * b.d); }
MIRK_TOP TapeDual<C> operator/(TapeDual<C> a, TapeDual<C> b) {
    const double q = a.v / b.v;
    return TapeDual<C>(q, (a.d - q * b.d) / b.v);
}
MIRK_TOP TapeDual<C> operator+(TapeDual<C> a, double b) { return TapeDual<C>(a.v + b, a.d); }
MIRK_TOP TapeDual<C> operator+(double a, TapeDual<C> b) { return TapeDual<C>(a + b.v, b.d); }
MIRK_TOP TapeDual<C> operator-(TapeDual<C> a, double b) { return TapeDual<C>(a.v - b, a.d); }
MIRK_TOP TapeDual<C> operator-(double a, TapeDual<C> b) { return TapeDual<C>(a - b.v, -b.d); }
MIRK_TOP TapeDual<C> operator*(TapeDual<C> a, double b) { return TapeDual<C>(a.v * b, a.d * b); }
MIRK_TOP TapeDual<C> operator*(double a, TapeDual<C> b) { return TapeDual<C>(a * b.v, a * b.d); }
MIRK_TOP TapeDual<C> operator/(TapeDual<C> a, double b) { return TapeDual<C>(a.v / b, a.d / b); }
MIRK_TOP TapeDual<C> operator/(double a, TapeDual<C> b) {
    const double q = a / b.v;
    return TapeDual<C>(q, -q * b.d / b.v);
}
MIRK_TOP TapeDual<C>& operator+=(TapeDual<C>& a, TapeDual<C> b) { a.v += b.v; a.d += b.d; return a; }
MIRK_TOP TapeDual<C>& operator-=(TapeDual<C>& a, TapeDual<C> b) { a.v -= b.v; a.d -= b.d; return a; }
MIRK_TOP TapeDual<C>& operator*=(TapeDual<C>& a, TapeDual<C> b) { a = a * b; return a; }
MIRK_TOP TapeDual<C>& operator+=(TapeDual<C>& a, double b) { a.v += b; return a; }
MIRK_TOP TapeDual<C>& operator*=(TapeDual<C>& a, double b) { a.v *= b; a.d *= b; return a; }

// record / replay of one entry: x = first result, y = second result (sincos) or unused
// (the tape exists on the device only; the host bodies just keep __host__ __device__ functors compilable)
template <int C> __host__ __device__ __forceinline__ void tape_put(double x, double y) {
#ifdef __CUDA_ARCH__
    double2* e = Tape<C>::next();
    if (e) *e = make_double2(x, y);
#else
    (void)x; (void)y;
#endif
}
// returns false when the call lies beyond the tape and has to be recomputed
template <int C> __host__ __device__ __forceinline__ bool tape_get(double& x, double& y) {
#ifdef __CUDA_ARCH__
    const double2* e = Tape<C>::next();
    if (!e) return false;
    const double2 t = *e;
    x = t.x; y = t.y;
    return true;
#else
    (void)x; (void)y;
    return false;
#endif
}
// One out-of-line copy of sincos per kernel: the recording pass of a pendulum chain calls it s x n/2 times per
// interval, and inlined (~100 instructions each) that is a 70 KB straight-line pass every warp streams once
// through the instruction cache (measured as `no_instruction` stalls); as a call it is ~1 KB of hot code.
#ifdef __CUDACC__
static __device__ __noinline__ void sincos_outlined(double a, double* s, double* c) { ::sincos(a, s, c); }
#endif
__host__ __device__ __forceinline__ void sincos_hd(double a, double& s, double& c) {
#ifdef __CUDA_ARCH__
    sincos_outlined(a, &s, &c);
#else
    s = ::sin(a); c = ::cos(a);
#endif
}

namespace fn {
MIRK_TOP RecVal<C> sin(RecVal<C> a) { double s, c; sincos_hd(a.v, s, c); tape_put<C>(s, c); return RecVal<C>(s); }
MIRK_TOP RecVal<C> cos(RecVal<C> a) { double s, c; sincos_hd(a.v, s, c); tape_put<C>(s, c); return RecVal<C>(c); }
MIRK_TOP RecVal<C> exp(RecVal<C> a) { const double e = ::exp(a.v); tape_put<C>(e, 0.0); return RecVal<C>(e); }
MIRK_TOP RecVal<C> log(RecVal<C> a) { const double l = ::log(a.v); tape_put<C>(l, 0.0); return RecVal<C>(l); }
MIRK_TOP RecVal<C> sqrt(RecVal<C> a) { const double r = ::sqrt(a.v); tape_put<C>(r, 0.0); return RecVal<C>(r); }
MIRK_TOP RecVal<C> tanh(RecVal<C> a) { const double t = ::tanh(a.v); tape_put<C>(t, 0.0); return RecVal<C>(t); }
MIRK_TOP RecVal<C> square(RecVal<C> a) { return RecVal<C>(a.v * a.v); }
MIRK_TOP double value(RecVal<C> a) { return a.v; }

MIRK_TOP TapeDual<C> sin(TapeDual<C> a) {
    double s, c;
    if (!tape_get<C>(s, c)) sincos_hd(a.v, s, c);
    return TapeDual<C>(s, c * a.d);
}
MIRK_TOP TapeDual<C> cos(TapeDual<C> a) {
    double s, c;
    if (!tape_get<C>(s, c)) sincos_hd(a.v, s, c);
    return TapeDual<C>(c, -s * a.d);
}
MIRK_TOP TapeDual<C> exp(TapeDual<C> a) {
    double e, u;
    if (!tape_get<C>(e, u)) e = ::exp(a.v);
    return TapeDual<C>(e, e * a.d);
}
MIRK_TOP TapeDual<C> log(TapeDual<C> a) {
    double l, u;
    if (!tape_get<C>(l, u)) l = ::log(a.v);
    return TapeDual<C>(l, a.d / a.v);
}
MIRK_TOP TapeDual<C> sqrt(TapeDual<C> a) {
    double r, u;
    if (!tape_get<C>(r, u)) r = ::sqrt(a.v);
    return TapeDual<C>(r, a.d / (2.0 * r));
}
MIRK_TOP TapeDual<C> tanh(TapeDual<C> a) {
    double t, u;
    if (!tape_get<C>(t, u)) t = ::tanh(a.v);
    return TapeDual<C>(t, (1.0 - t * t) * a.d);
}
MIRK_TOP TapeDual<C> square(TapeDual<C> a) { return TapeDual<C>(a.v * a.v, 2.0 * a.v * a.d); }
MIRK_TOP double value(TapeDual<C> a) { return a.v; }
}  // namespace fn
#undef MIRK_TOP

#endif  // __CUDACC__

}  // namespace mirk
