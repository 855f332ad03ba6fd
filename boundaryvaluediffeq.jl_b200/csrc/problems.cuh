// problems.cuh — device-function contract for BVPs and the built-in registry.
//
// The reference takes arbitrary Julia closures f!(du,u,p,t) and bc!(res,sol,p,t)
// (lib/BoundaryValueDiffEqMIRK/src/mirk.jl:71-116); a CUDA backend cannot call those, so the
// boundary is a *device functor*: a struct with
//     static constexpr int n, np, n_bc, n_bca, problem_type (0 Standard, 1 TwoPoint), max_bc_pts
//     template<class T> static void f(T* du, const T* u, const double* p, double t)
//     static int  bc_times(double* times, const double* p, double t0, double t1)
//     template<class T> static void bc(T* res, const T* U /* m×n */, const double* p)
//   optional (singular BVPs, prob.singular_term):  static constexpr bool has_singular_term = true;
//     template<class T> static void singular(T* out, const T* u, const double* p)     out = S u
//   optional (bc! reads sol(t, Val{1}), MIRK/src/interpolation.jl:277-292):  static constexpr bool bc_uses_derivative = true;
//     U[k] (m x n) is then followed by dU[k] = sol'(times[k]) (m x n more entries: U[m*n + k*n + c]).  As in the
//     reference the derivative is built from the Float64 stage buffers: it is a constant of the boundary Jacobian
// `T` is double for residuals; for Jacobians it is mirk::Dual (plain forward mode) or the pair mirk::RecVal /
// mirk::TapeDual of the taped kernel (tape.cuh) — one templated source serves all, like ForwardDiff on the
// Julia side.  Write elementary functions unqualified after `using namespace mirk::fn;` (sin, cos, exp, log,
// sqrt, tanh, square, value) so they resolve on every scalar type; arithmetic with double constants is defined
// on both sides of + - * /.  bc reads the solution only through U[k] = sol(times[k]),
// which covers every access style the reference's tests use.  TwoPoint problems
// have times = {t0, t1}, rows [0,n_bca) depend on U[0] only and the rest on U[1] only.
//
// Built-ins mirror oracle/mirk_problems.c id for id (the oracle is written independently, with
// hand-derived analytic Jacobians).
#pragma once
#include "dual.cuh"

namespace mirk {
namespace problems {

#define MIRK_PF template <class T> __host__ __device__ __forceinline__ static void
// the same with argument types left open: `du` only needs `du[k] = value`, `u` only `u[k]` — raw pointers, or the
// proxies of stagejac.cuh (seeds generated on the fly, tangents stored straight into the Jacobian).  Functors written
// this way must treat du as WRITE-ONLY.
#define MIRK_PFX template <class T, class DU = T*, class U = const T*> __host__ __device__ __forceinline__ static void
#define MIRK_PT __host__ __device__ __forceinline__ static int

constexpr double kPi = 3.14159265358979323846;

MIRK_PT ends_times(double* tm, double t0, double t1) { tm[0] = t0; tm[1] = t1; return 2; }

// 0: simple pendulum (benchmark/simple_pendulum.jl:5-19,32), p = [g/L]
struct Pendulum {
    static constexpr int n = 2, np = 1, n_bc = 2, n_bca = 0, problem_type = 0, max_bc_pts = 2;
    MIRK_PF f(T* du, const T* u, const double* p, double) {
        using namespace fn;
        du[0] = u[1];
        du[1] = -p[0] * sin(u[0]);
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) {
        tm[0] = (t0 + t1) / 2; tm[1] = t1; return 2;
    }
    MIRK_PF bc(T* r, const T* U, const double*) {
        r[0] = U[0] + kPi / 2;
        r[1] = U[2] - kPi / 2;
    }
};

// 1: u'' = -k u, two scalar conditions (mirk_basic_tests.jl:16-33, ensemble_tests.jl:10-18),
//    p = [k, ta, va, tb, vb, ca, cb]
struct Linear2 {
    static constexpr int n = 2, np = 7, n_bc = 2, n_bca = 0, problem_type = 0, max_bc_pts = 2;
    MIRK_PF f(T* du, const T* u, const double* p, double) {
        du[0] = u[1];
        du[1] = -p[0] * u[0];
    }
    MIRK_PT bc_times(double* tm, const double* p, double, double) { tm[0] = p[1]; tm[1] = p[3]; return 2; }
    MIRK_PF bc(T* r, const T* U, const double* p) {
        r[0] = U[0 + (int)p[5]] - p[2];
        r[1] = U[2 + (int)p[6]] - p[4];
    }
};

// 2: same ODE as a TwoPointBVProblem (mirk_basic_tests.jl:42-47), p = [k, va, vb]
struct Linear2TP {
    static constexpr int n = 2, np = 3, n_bc = 2, n_bca = 1, problem_type = 1, max_bc_pts = 2;
    MIRK_PF f(T* du, const T* u, const double* p, double) {
        du[0] = u[1];
        du[1] = -p[0] * u[0];
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) { return ends_times(tm, t0, t1); }
    MIRK_PF bc(T* r, const T* U, const double* p) {
        r[0] = U[0] - p[1];
        r[1] = U[2] - p[2];
    }
};

// 3: swirling flow III (mirk_basic_tests.jl:315-344), p = [eps]
struct Swirling {
    static constexpr int n = 6, np = 1, n_bc = 6, n_bca = 0, problem_type = 0, max_bc_pts = 2;
    MIRK_PF f(T* du, const T* u, const double* p, double) {
        const double e = p[0];
        du[0] = u[1];
        du[1] = (u[0] * u[3] - u[2] * u[1]) / e;
        du[2] = u[3];
        du[3] = u[4];
        du[4] = u[5];
        du[5] = (-u[2] * u[5] - u[0] * u[1]) / e;
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) { return ends_times(tm, t0, t1); }
    MIRK_PF bc(T* r, const T* U, const double*) {
        r[0] = U[0] + 1.0; r[1] = U[2]; r[2] = U[3];
        r[3] = U[6] - 1.0; r[4] = U[8]; r[5] = U[9];
    }
};

// 4: Lotka-Volterra, both conditions at t0 (mirk_basic_tests.jl:438-455), p = [a,b,c,d]
struct Lotka {
    static constexpr int n = 2, np = 4, n_bc = 2, n_bca = 0, problem_type = 0, max_bc_pts = 1;
    MIRK_PF f(T* du, const T* u, const double* p, double) {
        du[0] = p[0] * u[0] - p[1] * u[0] * u[1];
        du[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double) { tm[0] = t0; return 1; }
    MIRK_PF bc(T* r, const T* U, const double*) {
        r[0] = U[0] - 1.0;
        r[1] = U[1] - 2.0;
    }
};

// 5: geodesic on a torus (mirk_basic_tests.jl:726-756), p = [R, r, a1_1, a1_2, a2_1, a2_2]
struct Torus {
    static constexpr int n = 4, np = 6, n_bc = 4, n_bca = 0, problem_type = 0, max_bc_pts = 2;
    MIRK_PF f(T* du, const T* u, const double* p, double) {
        using namespace fn;
        const double R = p[0], r = p[1];
        const T s = sin(u[0]), c = cos(u[0]);
        const T Rt = R + r * c;
        du[0] = u[2];
        du[1] = u[3];
        du[2] = -(u[3] * u[3]) * Rt * s / r;
        du[3] = 2.0 * r * s / Rt * u[2] * u[3];
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) { return ends_times(tm, t0, t1); }
    MIRK_PF bc(T* r, const T* U, const double* p) {
        r[0] = U[0] - p[2]; r[1] = U[1] - p[3];
        r[2] = U[4] - p[4]; r[3] = U[5] - p[5];
    }
};

// 6: boundary layer (test/misc/adaptivity_tests.jl:7-17), p = [eps]
struct Layer {
    static constexpr int n = 2, np = 1, n_bc = 2, n_bca = 0, problem_type = 0, max_bc_pts = 2;
    MIRK_PF f(T* du, const T* u, const double* p, double t) {
        du[0] = u[1];
        du[1] = -t / p[0] * u[1] - kPi * kPi * ::cos(kPi * t) - kPi * t / p[0] * ::sin(kPi * t);
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) { return ends_times(tm, t0, t1); }
    MIRK_PF bc(T* r, const T* U, const double*) {
        r[0] = U[0] + 2.0;
        r[1] = U[2];
    }
};

// 7/8: chain of NP torsionally coupled pendula, two-point,
//      u = [th_1..th_NP, om_1..om_NP], p = [g, kappa, a_1..a_NP, b_1..b_NP]
template <int NP> struct Chain {
    static constexpr int tape_calls = NP;  // one sin per pendulum and f evaluation (sizes the Jacobian tape)
    static constexpr int n = 2 * NP, np = 2 + 2 * NP, n_bc = 2 * NP, n_bca = NP, problem_type = 1,
                         max_bc_pts = 2;
    MIRK_PF f(T* du, const T* u, const double* p, double) {
        using namespace fn;
        const double g = p[0], kap = p[1];
#pragma unroll
        for (int k = 0; k < NP; k++) {
            du[k] = u[NP + k];
            T acc = -g * sin(u[k]) - (2.0 * kap) * u[k];
            if (k > 0) acc = acc + kap * u[k - 1];
            if (k < NP - 1) acc = acc + kap * u[k + 1];
            du[NP + k] = acc;
        }
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) { return ends_times(tm, t0, t1); }
    MIRK_PF bc(T* r, const T* U, const double* p) {
#pragma unroll
        for (int k = 0; k < NP; k++) {
            r[k] = U[k] - p[2 + k];
            r[NP + k] = U[n + k] - p[2 + NP + k];
        }
    }
};

// 9: 2-D Bratu by the method of lines, M interior lines, two-point,
//    u = [u_1..u_M, v_1..v_M], p = [lambda]
template <int M> struct BratuMOL {
    static constexpr int n = 2 * M, np = 1, n_bc = 2 * M, n_bca = M, problem_type = 1, max_bc_pts = 2;
    MIRK_PFX f(DU du, U u, const double* p, double) {
        using namespace fn;
        constexpr double dz = 1.0 / (M + 1), idz2 = 1.0 / (dz * dz);
#pragma unroll 4
        for (int j = 0; j < M; j++) {
            du[j] = u[M + j];
            T lap = -2.0 * u[j];
            if (j > 0) lap = lap + u[j - 1];
            if (j < M - 1) lap = lap + u[j + 1];
            du[M + j] = -(lap * idz2) - p[0] * exp(u[j]);
        }
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) { return ends_times(tm, t0, t1); }
    MIRK_PF bc(T* r, const T* U, const double*) {
#pragma unroll 4
        for (int j = 0; j < M; j++) { r[j] = U[j]; r[M + j] = U[n + j]; }
    }
};

// 10: Lane-Emden equation of index 1 as a SINGULAR BVP  y' = S y / t + f(t, y)  (lib/BoundaryValueDiffEqMIRK/test/Core/
//     singular_bvp_tests.jl:15-61): y'' + (2/t) y' + y = 0, y(0) = 1, y(1) = sin(1), exact solution sin(t)/t.
//     `singular` returns S u (here S = [0 0; 0 -2]); the collocation adds it, divided by t, to every DISCRETE stage
//     with t > 0 (CORE/src/utils.jl:932-941, MIRK/src/collocation.jl:65) — not to the interpolation stages.
struct LaneEmden {
    static constexpr int n = 2, np = 0, n_bc = 2, n_bca = 1, problem_type = 1, max_bc_pts = 2;
    static constexpr bool has_singular_term = true;
    MIRK_PF f(T* du, const T* u, const double*, double) {
        du[0] = u[1];
        du[1] = -u[0];
    }
    MIRK_PF singular(T* out, const T* u, const double*) {
        out[0] = 0.0 * u[0];
        out[1] = -2.0 * u[1];
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) { return ends_times(tm, t0, t1); }
    MIRK_PF bc(T* r, const T* U, const double*) {
        r[0] = U[0] - 1.0;
        r[1] = U[2] - 0.84147098480789650665;  // sin(1)
    }
};
// 11: u'' = -u with a boundary condition on the DERIVATIVE of the interpolant (sol(t, Val{1}) inside bc!,
//     MIRK/src/interpolation.jl:277-292; the reference's tests reach it through maxsol / minsol, mirk_basic_tests.jl:453-479):
//     u1(t0) = 0,  u1(t1) - 1 + alpha (u1'(tm) - c) = 0,  tm = (t0 + t1) / 2;  p = [alpha, c].
//     On [0, pi/2] with c = cos(pi/4) the solution is (sin t, cos t).
struct RobinSine {
    static constexpr int n = 2, np = 2, n_bc = 2, n_bca = 0, problem_type = 0, max_bc_pts = 3;
    static constexpr bool bc_uses_derivative = true;
    MIRK_PF f(T* du, const T* u, const double*, double) {
        du[0] = u[1];
        du[1] = -u[0];
    }
    MIRK_PT bc_times(double* tm, const double*, double t0, double t1) {
        tm[0] = t0; tm[1] = (t0 + t1) / 2; tm[2] = t1; return 3;
    }
    MIRK_PF bc(T* r, const T* U, const double* p) {
        const T* dU = U + 3 * n;
        r[0] = U[0];
        r[1] = U[4] - 1.0 + p[0] * (dU[2] - p[1]);
    }
};


enum BuiltinId {
    kPendulum = 0, kLinear2 = 1, kLinear2TP = 2, kSwirling = 3, kLotka = 4, kTorus = 5, kLayer = 6,
    kChain8 = 7, kChain16 = 8, kBratu64 = 9, kLaneEmden = 10, kRobinSine = 11, kNumBuiltin = 12
};

}  // namespace problems
}  // namespace mirk
