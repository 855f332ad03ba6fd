// kernels.cuh — problem-templated device kernels of the MIRK Newton step.
//
//   k_residual   Phi_i for every interval + discrete stages K     (MIRK/src/collocation.jl:44-102)
//   k_bc         boundary rows, their Jacobian blocks, pinned nodes (MIRK/src/mirk.jl:471-534,565-571;
//                                                                   interpolation.jl:214-242, 300-336)
//   k_jac_blocks [L_i R_i] by one dual-number sweep per column      (MIRK/src/mirk.jl:810-838)
//   k_defect     interpolation stages + defect samples              (MIRK/src/adaptivity.jl:370-415)
//   k_reinterp   new guess on a new mesh                            (MIRK/src/adaptivity.jl:6-13,590-621)
//   k_interp     sol(t) / sol'(t) dense output                      (MIRK/src/interpolation.jl:98-204)
//
// HBM layout (all FP64, node-major like the reference's flat unknown vector, CORE/src/utils.jl:59-66):
//   mesh[N]  y[N][n]  Kd[N-1][s][n]  Ki[N-1][s*-s][n]  resid[L + (N-1) n]  errors[N-1][n]
//   Lb/Rb[N-1][n][n] row-major blocks of the almost-block-diagonal Jacobian.
#pragma once
// loops over the state dimension unroll fully up to n = 32 and by 4 beyond (n = 128 would explode)
#include "dual.cuh"
#include "tableau.cuh"
#include "tape.cuh"

namespace mirk {

__host__ __device__ constexpr int unroll_for(int n) { return n <= 32 ? 64 : 4; }

// ---- small helpers -----------------------------------------------------------------------------

// |x| as an integer that orders like the double (NaN > Inf > finite), for atomicMax reductions
__device__ __forceinline__ unsigned long long abs_bits(double x) {
    return (unsigned long long)__double_as_longlong(fabs(x));
}

__device__ __forceinline__ void block_max_to_global(unsigned long long m, unsigned long long* out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0 && m != 0ull) atomicMax(out, m);
}

// clamp(searchsortedfirst(mesh,t)-1, 1, N-1) as a 0-based interval (CORE/src/utils.jl:119-121)
__host__ __device__ __forceinline__ int interval_of(const double* mesh, int N, double t) {
    int lo = 0, hi = N;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (mesh[mid] < t) lo = mid + 1; else hi = mid;
    }
    int j = lo;
    if (j < 1) j = 1;
    if (j > N - 1) j = N - 1;
    return j - 1;
}

// singular BVPs y' = S y / t + f(t, y): the functor says so with `has_singular_term` and provides S u
template <class P, class = void> struct HasSingular { static constexpr bool value = false; };
template <class P> struct HasSingular<P, decltype((void)P::has_singular_term)> { static constexpr bool value = P::has_singular_term; };

// ---- one interval: stages and residual (Appendix A.1) -----------------------------------------
// K is [s][n] of T.  Works for T = double (residual) and T = Dual (one Jacobian column).
template <class P, int ORDER, class T>
__device__ __forceinline__ void phi_interval(const T* __restrict__ yi, const T* __restrict__ yi1,
                                             double h, double ti, const double* __restrict__ p,
                                             T (*K)[P::n], T* __restrict__ phi) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n;
    T tmp[n];
#pragma unroll (unroll_for(n))
    for (int r = 0; r < TB::s; r++) {
        const double vr = TB::v(r);
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) {
            if (vr == 0.0) tmp[k] = yi[k];
            else if (vr == 1.0) tmp[k] = yi1[k];
            else tmp[k] = (1.0 - vr) * yi[k] + vr * yi1[k];
        }
#pragma unroll (unroll_for(n))
        for (int j = 0; j < r; j++) {
            const double xrj = TB::x(r, j);
            if (xrj != 0.0) {
                const double hx = h * xrj;
#pragma unroll (unroll_for(n))
                for (int k = 0; k < n; k++) tmp[k] = tmp[k] + hx * K[j][k];
            }
        }
        P::template f<T>(K[r], tmp, p, ti + TB::c(r) * h);
        if constexpr (HasSingular<P>::value) {
            // __add_singular_term!: K_r += S tmp / t for t > 0 (CORE/src/utils.jl:932-941)
            const double tt = ti + TB::c(r) * h;
            if (tt > 0.0) {
                T sv[n];
                P::template singular<T>(sv, tmp, p);
                const double it = 1.0 / tt;
#pragma unroll (unroll_for(n))
                for (int k = 0; k < n; k++) K[r][k] = K[r][k] + it * sv[k];
            }
        }
    }
#pragma unroll (unroll_for(n))
    for (int k = 0; k < n; k++) {
        T acc = TB::b(0) * K[0][k];
#pragma unroll (unroll_for(n))
        for (int r = 1; r < TB::s; r++) acc = acc + TB::b(r) * K[r][k];
        phi[k] = yi1[k] - yi[k] - h * acc;
    }
}

// interpolation stages of one interval (Appendix A.3); Kd row pointer [s][n], writes KI [si][n]
template <class P, int ORDER>
__device__ __forceinline__ void interp_stages_interval(const double* __restrict__ yi,
                                                       const double* __restrict__ yi1, double h,
                                                       double ti, const double* __restrict__ p,
                                                       const double* __restrict__ K,
                                                       double* __restrict__ KI) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n;
    double tmp[n], out[n];
#pragma unroll (unroll_for(n))
    for (int r = 0; r < TB::si; r++) {
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) tmp[k] = 0.0;
#pragma unroll (unroll_for(n))
        for (int j = 0; j < TB::s; j++) {
            const double xs = TB::x_star(r, j);
            if (xs != 0.0) {
#pragma unroll (unroll_for(n))
                for (int k = 0; k < n; k++) tmp[k] += xs * K[j * n + k];
            }
        }
#pragma unroll (unroll_for(n))
        for (int j = 0; j < r; j++) {
            const double xs = TB::x_star(r, TB::s + j);
            if (xs != 0.0) {
#pragma unroll (unroll_for(n))
                for (int k = 0; k < n; k++) tmp[k] += xs * KI[j * n + k];
            }
        }
        const double vs = TB::v_star(r);
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) tmp[k] = tmp[k] * h + (1.0 - vs) * yi[k] + vs * yi1[k];
        P::template f<double>(out, tmp, p, ti + TB::c_star(r) * h);
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) KI[r * n + k] = out[k];
    }
}

// ---- K1: collocation residual -----------------------------------------------------------------
// one thread per interval; writes Kd, Phi (into resid at `phi_off`) and |Phi|_inf
template <class P, int ORDER>
__global__ void __launch_bounds__(128)
k_residual(int N, const double* __restrict__ mesh, const double* __restrict__ y,
           const double* __restrict__ p, double* __restrict__ Kd, double* __restrict__ phi_out,
           unsigned long long* __restrict__ norm_bits) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long m = 0ull;
    if (i < N - 1) {
        double yi[n], yi1[n], K[TB::s][n], phi[n];
        const double* yp = y + (size_t)i * n;
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) { yi[k] = yp[k]; yi1[k] = yp[n + k]; }
        const double ti = mesh[i], h = mesh[i + 1] - ti;
        phi_interval<P, ORDER, double>(yi, yi1, h, ti, p, K, phi);
        double* Ko = Kd + (size_t)i * TB::s * n;
#pragma unroll (unroll_for(n))
        for (int r = 0; r < TB::s; r++)
#pragma unroll (unroll_for(n))
            for (int k = 0; k < n; k++) Ko[r * n + k] = K[r][k];
        double* po = phi_out + (size_t)i * n;
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) {
            po[k] = phi[k];
            const unsigned long long b = abs_bits(phi[k]);
            m = b > m ? b : m;
        }
    }
    block_max_to_global(m, norm_bits);
}

// ---- boundary rows ------------------------------------------------------------------------------
// One block.  Thread 0 evaluates the solution at the BC times (continuous extension for interior
// times, end-point short-circuit as EvalSol does), all threads then differentiate bc with one dual
// seed each.  Reference pattern for the Jacobian: the derivative of an interior
// evaluation lands on the LEFT node of its interval only (interpolation.jl:227,239).
// Outputs: resid BC rows, bc_nodes[m], Bc[m][L][n], *m_out, |bc|_inf into norm_bits.
template <class P, int ORDER>
__global__ void __launch_bounds__(256)
k_bc(int N, const double* __restrict__ mesh, const double* __restrict__ y,
     const double* __restrict__ p, const double* __restrict__ Kd, double* __restrict__ Ki,
     double* __restrict__ resid, int* __restrict__ bc_nodes, double* __restrict__ Bc,
     int* __restrict__ m_out, unsigned long long* __restrict__ norm_bits, int want_jac) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n, L = P::n_bc;
    // functors that read sol(t, Val{1}) get dU[k] = sol'(times[k]) behind U (problems.cuh: bc_uses_derivative)
    constexpr bool kDeriv = BcUsesDerivative<P>::value;
    constexpr int UW = kDeriv ? 2 : 1;
    __shared__ double U[UW * P::max_bc_pts * n];
    __shared__ int s_m;
    if (threadIdx.x == 0) {
        double tm[P::max_bc_pts];
        int m;
        if (P::problem_type == 1) { m = 2; tm[0] = mesh[0]; tm[1] = mesh[N - 1]; }
        else m = P::bc_times(tm, p, mesh[0], mesh[N - 1]);
        s_m = m;
        for (int k = 0; k < m; k++) {
            const double t = tm[k];
            if (t == mesh[0]) {
                bc_nodes[k] = 0;
                for (int c = 0; c < n; c++) U[k * n + c] = y[c];
            } else if (t == mesh[N - 1]) {
                bc_nodes[k] = N - 1;
                for (int c = 0; c < n; c++) U[k * n + c] = y[(size_t)(N - 1) * n + c];
            } else {
                const int i = interval_of(mesh, N, t);
                bc_nodes[k] = i;
                const double ti = mesh[i], h = mesh[i + 1] - ti, tau = (t - ti) / h;
                double yi[n], yi1[n];
                for (int c = 0; c < n; c++) { yi[c] = y[(size_t)i * n + c]; yi1[c] = y[(size_t)(i + 1) * n + c]; }
                const double* K = Kd + (size_t)i * TB::s * n;
                double* KI = Ki + (size_t)i * TB::si * n;
                interp_stages_interval<P, ORDER>(yi, yi1, h, ti, p, K, KI);
                double w[TB::s_star], wp[TB::s_star];
                TB::weights(tau, w, wp);
                for (int c = 0; c < n; c++) {
                    double z = 0.0;
                    for (int r = 0; r < TB::s; r++) z += K[r * n + c] * w[r];
                    for (int r = 0; r < TB::si; r++) z += KI[r * n + c] * w[TB::s + r];
                    U[k * n + c] = z * h + yi[c];
                }
            }
        }
        if (kDeriv) {
            // sol(t, Val{1}) of EvalSol (MIRK/src/interpolation.jl:277-292): no end-point short cut; interval(t), the
            // derivative weights on the discrete and interpolation stages of that interval (plain Float64: a constant
            // of the boundary Jacobian below)
            for (int k = 0; k < m; k++) {
                const double t = tm[k];
                const int i = interval_of(mesh, N, t);
                const double ti = mesh[i], h = mesh[i + 1] - ti, tau = (t - ti) / h;
                double yi[n], yi1[n];
                for (int c = 0; c < n; c++) { yi[c] = y[(size_t)i * n + c]; yi1[c] = y[(size_t)(i + 1) * n + c]; }
                const double* K = Kd + (size_t)i * TB::s * n;
                double* KI = Ki + (size_t)i * TB::si * n;
                interp_stages_interval<P, ORDER>(yi, yi1, h, ti, p, K, KI);
                double w[TB::s_star], wp[TB::s_star];
                TB::weights(tau, w, wp);
                for (int c = 0; c < n; c++) {
                    double z = 0.0;
                    for (int r = 0; r < TB::s; r++) z += K[r * n + c] * wp[r];
                    for (int r = 0; r < TB::si; r++) z += KI[r * n + c] * wp[TB::s + r];
                    U[(m + k) * n + c] = z;
                }
            }
        }
        *m_out = m;
    }
    __syncthreads();
    const int m = s_m;
    const int La = P::problem_type == 1 ? P::n_bca : L;
    const size_t tail_off = (size_t)La + (size_t)(N - 1) * n;   // two-point: bc_b rows go last
    if (threadIdx.x == 0) {
        double Uv[UW * P::max_bc_pts * n], r[L];
        for (int e = 0; e < UW * m * n; e++) Uv[e] = U[e];
        P::template bc<double>(r, Uv, p);
        unsigned long long mb = 0ull;
        for (int q = 0; q < L; q++) {
            if (q < La) resid[q] = r[q]; else resid[tail_off + (q - La)] = r[q];
            const unsigned long long b = abs_bits(r[q]);
            mb = b > mb ? b : mb;
        }
        if (mb) atomicMax(norm_bits, mb);
    }
    if (want_jac) {
        for (int d = threadIdx.x; d < m * n; d += blockDim.x) {
            Dual Ud[UW * P::max_bc_pts * n], r[L];
            for (int e = 0; e < UW * m * n; e++) Ud[e] = Dual(U[e], e == d ? 1.0 : 0.0);
            P::template bc<Dual>(r, Ud, p);
            const int k = d / n, c = d % n;
            for (int q = 0; q < L; q++) Bc[((size_t)k * L + q) * n + c] = r[q].d;
        }
    }
}

// ---- K2: Jacobian blocks -----------------------------------------------------------------------
// one thread per (interval, column d of [L_i R_i]); d < n seeds y_i, d >= n seeds y_{i+1}.
template <class P, int ORDER>
__global__ void __launch_bounds__(128)
k_jac_blocks(int N, const double* __restrict__ mesh, const double* __restrict__ y,
             const double* __restrict__ p, double* __restrict__ Lb, double* __restrict__ Rb) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid / (2 * n)), d = (int)(gid % (2 * n));
    if (i >= N - 1) return;
    Dual yi[n], yi1[n], K[TB::s][n], phi[n];
    const double* yp = y + (size_t)i * n;
#pragma unroll (unroll_for(n))
    for (int k = 0; k < n; k++) {
        yi[k] = Dual(yp[k], k == d ? 1.0 : 0.0);
        yi1[k] = Dual(yp[n + k], (n + k) == d ? 1.0 : 0.0);
    }
    const double ti = mesh[i], h = mesh[i + 1] - ti;
    phi_interval<P, ORDER, Dual>(yi, yi1, h, ti, p, K, phi);
    double* out = (d < n ? Lb : Rb) + (size_t)i * n * n + (d < n ? d : d - n);
#pragma unroll (unroll_for(n))
    for (int k = 0; k < n; k++) out[k * n] = phi[k].d;
}

// ---- K1+K2 fused: the dual sweep of column 0 carries the residual for free ------------------------
// Same thread mapping as k_jac_blocks; the thread of column d = 0 also writes the discrete stages,
// Phi_i and |Phi|_inf, so a Newton iteration needs no separate residual pass over the mesh.
template <class P, int ORDER>
__global__ void __launch_bounds__(128)
k_resjac(int N, const double* __restrict__ mesh, const double* __restrict__ y, const double* __restrict__ p,
         double* __restrict__ Kd, double* __restrict__ phi_out, unsigned long long* __restrict__ norm_bits,
         double* __restrict__ Lb, double* __restrict__ Rb) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid / (2 * n)), d = (int)(gid % (2 * n));
    unsigned long long m = 0ull;
    if (i < N - 1) {
        Dual yi[n], yi1[n], K[TB::s][n], phi[n];
        const double* yp = y + (size_t)i * n;
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) {
            yi[k] = Dual(yp[k], k == d ? 1.0 : 0.0);
            yi1[k] = Dual(yp[n + k], (n + k) == d ? 1.0 : 0.0);
        }
        const double ti = mesh[i], h = mesh[i + 1] - ti;
        phi_interval<P, ORDER, Dual>(yi, yi1, h, ti, p, K, phi);
        double* out = (d < n ? Lb : Rb) + (size_t)i * n * n + (d < n ? d : d - n);
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) out[k * n] = phi[k].d;
        if (d == 0) {
            double* Ko = Kd + (size_t)i * TB::s * n;
#pragma unroll (unroll_for(n))
            for (int r = 0; r < TB::s; r++)
#pragma unroll (unroll_for(n))
                for (int k = 0; k < n; k++) Ko[r * n + k] = K[r][k].v;
            double* po = phi_out + (size_t)i * n;
#pragma unroll (unroll_for(n))
            for (int k = 0; k < n; k++) {
                po[k] = phi[k].v;
                const unsigned long long b = abs_bits(phi[k].v);
                m = b > m ? b : m;
            }
        }
    }
    block_max_to_global(m, norm_bits);
}

// ---- K1+K2 taped: values once per interval, tangents replay the tape (tape.cuh) ------------------------
// A CTA of 1..4 warps takes `ipw` <= 32 consecutive intervals.
//   pass 1  first warp, lane = interval: stages, Phi_i, |Phi|_inf in plain FP64; every sin/cos/exp result goes
//           to the tape (one pass-1 per CTA: with several warps its cost is shared by all of them)
//   pass 2  all warps, thread = (interval, column d of [L_i R_i]): tangent sweep with the elementary functions
//           read back from the tape.  For right-hand sides whose only non-linearities are elementary functions of the
//           state (the pendulum chains) the whole value computation of this pass is dead code.
// The optional `P::tape_calls` (elementary-function calls per f evaluation) sizes the tape; default n.
template <class P, class = void> struct TapeCalls { static constexpr int value = P::n; };
template <class P> struct TapeCalls<P, decltype((void)P::tape_calls)> { static constexpr int value = P::tape_calls; };
template <class P, int ORDER> __host__ __device__ constexpr int tape_cap() {
    const int want = Tableau<ORDER>::s * TapeCalls<P>::value;
    return want < 1 ? 1 : (want > 92 ? 92 : want);  // 92 entries x 32 slots x 16 B stays under 48 KB static
}

template <class P, int ORDER>
#ifndef MIRK_TAPE_MINB
#define MIRK_TAPE_MINB 1
#endif
__global__ void __launch_bounds__(kTapeThreads, MIRK_TAPE_MINB)
k_resjac_tape(int N, int ipw, const double* __restrict__ mesh, const double* __restrict__ y,
              const double* __restrict__ p, double* __restrict__ Kd, double* __restrict__ phi_out,
              unsigned long long* __restrict__ norm_bits, double* __restrict__ Lb, double* __restrict__ Rb) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n, cols = 2 * n, CAP = tape_cap<P, ORDER>();
    using TP = Tape<CAP>;
    const int lane = threadIdx.x;  // pass 1 runs on the first warp only
    const int i0 = blockIdx.x * ipw;
    unsigned long long m = 0ull;
    {
        const int i = i0 + lane;
        if (lane < ipw && lane < 32 && i < N - 1) {
            using V = RecVal<CAP>;
            TP::begin(lane);
            V yi[n], yi1[n], K[TB::s][n], phi[n];
            const double* yp = y + (size_t)i * n;
#pragma unroll (unroll_for(n))
            for (int k = 0; k < n; k++) { yi[k] = V(yp[k]); yi1[k] = V(yp[n + k]); }
            const double ti = mesh[i], h = mesh[i + 1] - ti;
            phi_interval<P, ORDER, V>(yi, yi1, h, ti, p, K, phi);
            double* Ko = Kd + (size_t)i * TB::s * n;
#pragma unroll (unroll_for(n))
            for (int r = 0; r < TB::s; r++)
#pragma unroll (unroll_for(n))
                for (int k = 0; k < n; k++) Ko[r * n + k] = K[r][k].v;
            double* po = phi_out + (size_t)i * n;
#pragma unroll (unroll_for(n))
            for (int k = 0; k < n; k++) {
                po[k] = phi[k].v;
                const unsigned long long b = abs_bits(phi[k].v);
                m = b > m ? b : m;
            }
        }
    }
    block_max_to_global(m, norm_bits);
    __syncthreads();
    int here = N - 1 - i0;
    if (here > ipw) here = ipw;
    const int items = here * cols;
    for (int e = threadIdx.x; e < items; e += blockDim.x) {
        using D = TapeDual<CAP>;
        const int li = e / cols, d = e % cols, i = i0 + li;
        TP::begin(li);
        D yi[n], yi1[n], K[TB::s][n], phi[n];
        const double* yp = y + (size_t)i * n;
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) {
            yi[k] = D(yp[k], k == d ? 1.0 : 0.0);
            yi1[k] = D(yp[n + k], (n + k) == d ? 1.0 : 0.0);
        }
        const double ti = mesh[i], h = mesh[i + 1] - ti;
        phi_interval<P, ORDER, D>(yi, yi1, h, ti, p, K, phi);
        double* out = (d < n ? Lb : Rb) + (size_t)i * n * n + (d < n ? d : d - n);
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) out[k * n] = phi[k].d;
    }
}

// ---- K4: defect estimate (Appendix A.5) ---------------------------------------------------------
// one thread per interval: interpolation stages, two samples, errors[i][:], est[i] = |errors_i|_inf
template <class P, int ORDER>
__global__ void __launch_bounds__(128)
k_defect(int N, const double* __restrict__ mesh, const double* __restrict__ y,
         const double* __restrict__ p, const double* __restrict__ Kd, double* __restrict__ Ki,
         double* __restrict__ errors, double* __restrict__ est,
         unsigned long long* __restrict__ defect_bits) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mb = 0ull;
    if (i < N - 1) {
        double yi[n], yi1[n];
        const double* yp = y + (size_t)i * n;
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) { yi[k] = yp[k]; yi1[k] = yp[n + k]; }
        const double ti = mesh[i], h = mesh[i + 1] - ti;
        const double* K = Kd + (size_t)i * TB::s * n;
        double* KI = Ki + (size_t)i * TB::si * n;
        interp_stages_interval<P, ORDER>(yi, yi1, h, ti, p, K, KI);
        double d1[n], d2[n];
        double e1 = 0.0, e2 = 0.0;
#pragma unroll (unroll_for(n))
        for (int smp = 0; smp < 2; smp++) {
            const double tau = smp ? (1.0 - TB::tau_star()) : TB::tau_star();
            double w[TB::s_star], wp[TB::s_star], z[n], zp[n], g[n];
            TB::weights(tau, w, wp);
#pragma unroll (unroll_for(n))
            for (int k = 0; k < n; k++) {
                double a = 0.0, b = 0.0;
#pragma unroll (unroll_for(n))
                for (int r = 0; r < TB::s; r++) { a += K[r * n + k] * w[r]; b += K[r * n + k] * wp[r]; }
#pragma unroll (unroll_for(n))
                for (int r = 0; r < TB::si; r++) { a += KI[r * n + k] * w[TB::s + r]; b += KI[r * n + k] * wp[TB::s + r]; }
                z[k] = a * h + yi[k];
                zp[k] = b;
            }
            P::template f<double>(g, z, p, ti + tau * h);
            double e = 0.0;
#pragma unroll (unroll_for(n))
            for (int k = 0; k < n; k++) {
                const double dd = (zp[k] - g[k]) / (fabs(g[k]) + 1.0);
                if (smp) d2[k] = dd; else d1[k] = dd;
                if (fabs(dd) > e) e = fabs(dd);
            }
            if (smp) e2 = e; else e1 = e;
        }
        const bool first = e1 > e2;
        double em = 0.0;
#pragma unroll (unroll_for(n))
        for (int k = 0; k < n; k++) {
            const double v = first ? d1[k] : d2[k];
            errors[(size_t)i * n + k] = v;
            const unsigned long long b = abs_bits(v);
            mb = b > mb ? b : mb;
            if (fabs(v) > em) em = fabs(v);
        }
        est[i] = em;
    }
    block_max_to_global(mb, defect_bits);
}

// ---- interpolation stages for every interval (needed for dense output of non-adaptive runs) ----
template <class P, int ORDER>
__global__ void __launch_bounds__(128)
k_interp_setup(int N, const double* __restrict__ mesh, const double* __restrict__ y,
               const double* __restrict__ p, const double* __restrict__ Kd, double* __restrict__ Ki) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N - 1) return;
    double yi[n], yi1[n];
    const double* yp = y + (size_t)i * n;
#pragma unroll (unroll_for(n))
    for (int k = 0; k < n; k++) { yi[k] = yp[k]; yi1[k] = yp[n + k]; }
    const double ti = mesh[i], h = mesh[i + 1] - ti;
    interp_stages_interval<P, ORDER>(yi, yi1, h, ti, p, Kd + (size_t)i * TB::s * n,
                                     Ki + (size_t)i * TB::si * n);
}

// ---- continuous extension at arbitrary times (Appendix A.4) -------------------------------------
// out[j][:] = u(ts[j]) (deriv = 0) or u'(ts[j]) (deriv = 1); one thread per (time, component).
// mode 1 (re-interpolation onto a new mesh) also returns the old interval index per time.
template <int ORDER>
__global__ void k_interp(int n, int N, const double* __restrict__ mesh, const double* __restrict__ y,
                         const double* __restrict__ Kd, const double* __restrict__ Ki, int nt,
                         const double* __restrict__ ts, int deriv, int add_base,
                         double* __restrict__ out, int* __restrict__ iold) {
    using TB = Tableau<ORDER>;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = (int)(gid / n), k = (int)(gid % n);
    if (j >= nt) return;
    const double t = ts[j];
    const int i = interval_of(mesh, N, t);
    const double ti = mesh[i], h = mesh[i + 1] - ti, tau = (t - ti) / h;
    double w[TB::s_star], wp[TB::s_star];
    TB::weights(tau, w, wp);
    const double* ww = deriv ? wp : w;
    const double* K = Kd + (size_t)i * TB::s * n;
    const double* KI = Ki + (size_t)i * TB::si * n;
    double z = 0.0;
#pragma unroll
    for (int r = 0; r < TB::s; r++) z += K[r * n + k] * ww[r];
#pragma unroll
    for (int r = 0; r < TB::si; r++) z += KI[r * n + k] * ww[TB::s + r];
    if (!deriv) { z *= h; if (add_base) z += y[(size_t)i * n + k]; }
    out[(size_t)j * n + k] = z;
    if (iold && k == 0) iold[j] = i;
}

}  // namespace mirk
