// stagejac.cuh — Jacobian blocks [L_i R_i] for LARGE state dimension (n = 64, 128: BASELINE config C4) by stage-wise
// Jacobians and dense FP64 tensor-core chain-rule products:
//
//     J_r = df/dy(Y_r)                                   one dual evaluation of f per column, seeds generated on the fly
//     A_r = dK_r/dy_i     = J_r [ (1 - v_r) I + h sum_{j<r} x_rj A_j ]
//     B_r = dK_r/dy_{i+1} = J_r [      v_r  I + h sum_{j<r} x_rj B_j ]      one n x n x 2n DMMA product per implicit stage
//     L_i = -I - h sum_r b_r A_r ,   R_i = I - h sum_r b_r B_r
//
// This replaces the per-column dual sweep of the whole interval (k_resjac_tape), whose 2n sweeps each carry n-long
// tangent vectors for every stage: at n = 128 those live in local memory (15 ms per C4 step in round 2's first
// measurement, half the step).  What the reference does here is coloured ForwardDiff of the collocation loss
// (lib/BoundaryValueDiffEqMIRK/src/mirk.jl:810-838); the result is the same exact block Jacobian.
//
//   k_stage_jac   CTA per interval, thread d = column d.  Stage arguments Y_r are rebuilt from the discrete stages the
//                 residual pass stored; f is called with PROXY arguments: the input proxy yields Dual(Y_r[k], k == d)
//                 from shared memory and the output proxy stores the tangent of du[k] straight to J_r[k][d] in global
//                 memory (coalesced over d) — no n-long per-thread arrays.  Functors whose f only accepts raw pointers
//                 fall back to per-thread arrays (correct, slower).
//   k_chain_gemm  CTA per interval, 8 warps.  For every implicit stage: S_r = J_r * M_r (n x 2n) with the B operand
//                 M_r = [(1 - v_r) I + h sum x_rj A_j | v_r I + h sum x_rj B_j] assembled while its tiles are staged to
//                 shared memory; 16-deep K chunks, register-staged double buffering, DMMA m8n8k4 with the fragment
//                 strides of abd_block.cuh.  The epilogue of the last implicit stage combines all stages into L_i, R_i.
#pragma once
#include <type_traits>

#include "dual.cuh"
#include "tableau.cuh"

namespace mirk {

__device__ __forceinline__ void dmma_acc(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// ---- proxies ------------------------------------------------------------------------------------------
struct SeedIn {  // u[k] = Dual(Y[k], k == d)
    const double* y;
    int d;
    __device__ __forceinline__ Dual operator[](int k) const { return Dual(y[k], k == d ? 1.0 : 0.0); }
};
struct ColOut {  // du[k] = x  stores x.d to J[k * n + d]
    double* col;
    int n;
    struct Ref {
        double* p;
        __device__ __forceinline__ void operator=(const Dual& x) const { *p = x.d; }
    };
    __device__ __forceinline__ Ref operator[](int k) const { return Ref{col + (size_t)k * n}; }
};
template <class P, class = void> struct AcceptsProxies : std::false_type {};
template <class P>
struct AcceptsProxies<P, decltype(P::template f<Dual>(std::declval<ColOut>(), std::declval<SeedIn>(), (const double*)nullptr, 0.0))>
    : std::true_type {};

template <int ORDER> struct StageShape {
    using TB = Tableau<ORDER>;
    __host__ __device__ static constexpr bool implicit_stage(int r) {
        for (int j = 0; j < r; j++)
            if (TB::x(r, j) != 0.0) return true;
        return false;
    }
    __host__ __device__ static constexpr int last_implicit() {
        int l = -1;
        for (int r = 0; r < TB::s; r++)
            if (implicit_stage(r)) l = r;
        return l;
    }
    // slot of stage r's product S_r in the scratch (only implicit stages before the last one are stored)
    __host__ __device__ static constexpr int slot(int r) {
        int k = 0;
        for (int j = 0; j < r; j++)
            if (implicit_stage(j)) k++;
        return k;
    }
    __host__ __device__ static constexpr int stored() { return last_implicit() < 0 ? 0 : slot(last_implicit()); }
};

template <class P, int ORDER> __host__ __device__ constexpr size_t stagejac_doubles_per_interval() {
    return (size_t)Tableau<ORDER>::s * P::n * P::n + (size_t)StageShape<ORDER>::stored() * 2 * P::n * P::n;
}

// ---- k_stage_jac ----------------------------------------------------------------------------------------
template <class P, int ORDER>
__global__ void __launch_bounds__(P::n)
k_stage_jac(int i_first, int N, const double* __restrict__ mesh, const double* __restrict__ y,
            const double* __restrict__ p, const double* __restrict__ Kd, double* __restrict__ scratch) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n, s = TB::s;
    __shared__ double Y[s][n];
    const int i = i_first + blockIdx.x, d = threadIdx.x;
    if (i >= N - 1) return;
    const double ti = mesh[i], h = mesh[i + 1] - ti;
    {
        const double yi = y[(size_t)i * n + d], yi1 = y[(size_t)(i + 1) * n + d];
        const double* K = Kd + (size_t)i * s * n;
#pragma unroll
        for (int r = 0; r < s; r++) {
            const double vr = TB::v(r);
            double a = vr == 0.0 ? yi : vr == 1.0 ? yi1 : (1.0 - vr) * yi + vr * yi1;
#pragma unroll
            for (int j = 0; j < r; j++)
                if (TB::x(r, j) != 0.0) a = a + (h * TB::x(r, j)) * K[j * n + d];
            Y[r][d] = a;
        }
    }
    __syncthreads();
    double* J = scratch + (size_t)blockIdx.x * stagejac_doubles_per_interval<P, ORDER>();
#pragma unroll 1
    for (int r = 0; r < s; r++) {
        const double tr = ti + TB::c(r) * h;
        if constexpr (AcceptsProxies<P>::value) {
            P::template f<Dual>(ColOut{J + (size_t)r * n * n + d, n}, SeedIn{Y[r], d}, p, tr);
        } else {
            Dual u[n], du[n];
#pragma unroll 4
            for (int k = 0; k < n; k++) u[k] = Dual(Y[r][k], k == d ? 1.0 : 0.0);
            P::template f<Dual>(du, u, p, tr);
#pragma unroll 4
            for (int k = 0; k < n; k++) J[(size_t)r * n * n + (size_t)k * n + d] = du[k].d;
        }
    }
}

// ---- k_chain_gemm -----------------------------------------------------------------------------------------
__device__ __forceinline__ double sj_lds(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

template <int n, int ORDER> struct ChainGemm {
    using TB = Tableau<ORDER>;
    using SH = StageShape<ORDER>;
    static constexpr int s = TB::s, RB = n / 32, CG = 8 / RB, PW = (2 * n / CG) < 64 ? (2 * n / CG) : 64,
                         NPASS = 2 * n / (CG * PW), TJ = PW / 8, KC = 16, AS = n + 4, BW = CG * PW, BS = BW + 4;
    static constexpr size_t nn = (size_t)n * n, per = (size_t)s * nn + (size_t)SH::stored() * 2 * nn;
    static constexpr int smem_doubles = 2 * KC * AS + 2 * KC * BS;

    // A_j[k][c] / B_j[k][c] of stage j (part = 0: A, 1: B): explicit stages from J_j, implicit ones from the scratch
    template <int j> __device__ __forceinline__ static double stage_elem(const double* J, const double* Sst, int part, int k, int c) {
        if constexpr (SH::implicit_stage(j)) {
            return Sst[(size_t)SH::slot(j) * 2 * nn + (size_t)k * 2 * n + part * n + c];
        } else {
            const double w = part ? TB::v(j) : 1.0 - TB::v(j);
            return w == 0.0 ? 0.0 : w * J[(size_t)j * nn + (size_t)k * n + c];
        }
    }
    // element (k, cfull) of M_R = [(1 - v_R) I + h sum_j x_Rj A_j | v_R I + h sum_j x_Rj B_j]
    template <int R, int j> __device__ __forceinline__ static void m_accum(const double* J, const double* Sst, double h, int part, int k,
                                                                            int c, double& m) {
        if constexpr (j < R) {
            if constexpr (TB::x(R, j) != 0.0) m = fma(h * TB::x(R, j), stage_elem<j>(J, Sst, part, k, c), m);
            m_accum<R, j + 1>(J, Sst, h, part, k, c, m);
        }
    }
    template <int R, int q> __device__ __forceinline__ static void lr_accum(const double* J, const double* Sst, int part, int k, int c,
                                                                             double& s0, double& s1) {
        if constexpr (q < s) {
            if constexpr (q != R) {
                s0 = fma(TB::b(q), stage_elem<q>(J, Sst, part, k, c), s0);
                s1 = fma(TB::b(q), stage_elem<q>(J, Sst, part, k, c + 1), s1);
            }
            lr_accum<R, q + 1>(J, Sst, part, k, c, s0, s1);
        }
    }

    template <int R>
    __device__ __forceinline__ static void do_stage(const double* J, double* Sst, double h, double* As, double* Bs, double* Lout,
                                                    double* Rout) {
        if constexpr (SH::implicit_stage(R)) {
            constexpr bool last = R == SH::last_implicit();
            const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
            const int rb = warp % RB, cg = warp / RB;
            const unsigned sA = (unsigned)__cvta_generic_to_shared(As), sB = (unsigned)__cvta_generic_to_shared(Bs);
            const double* Jr = J + (size_t)R * nn;
            constexpr int A_PER = n * KC / 256, B_PER = KC * BW / 256, BT = BW / B_PER;  // B: BT threads per k row
            const int a_row = tid % n, a_k0 = (tid / n) * A_PER;
            const int b_k = tid / BT, b_c0 = tid % BT;
#pragma unroll 1
            for (int ps = 0; ps < NPASS; ps++) {
                const int colbase = ps * BW;  // first of the BW columns of [A | B] this pass produces
                double acc[4][TJ][2];
#pragma unroll
                for (int tr = 0; tr < 4; tr++)
#pragma unroll
                    for (int j = 0; j < TJ; j++) { acc[tr][j][0] = 0.0; acc[tr][j][1] = 0.0; }
                double ra[A_PER], rbv[B_PER];
                auto load_chunk = [&](int kk) {
#pragma unroll
                    for (int e = 0; e < A_PER; e++) ra[e] = Jr[(size_t)a_row * n + kk + a_k0 + e];
                    const int k = kk + b_k;
#pragma unroll
                    for (int e = 0; e < B_PER; e++) {
                        const int cfull = colbase + b_c0 + BT * e, part = cfull / n, c = cfull % n;
                        double m = (k == c) ? (part ? TB::v(R) : 1.0 - TB::v(R)) : 0.0;
                        m_accum<R, 0>(J, Sst, h, part, k, c, m);
                        rbv[e] = m;
                    }
                };
                auto store_chunk = [&](int buf) {
#pragma unroll
                    for (int e = 0; e < A_PER; e++) As[buf * KC * AS + (a_k0 + e) * AS + a_row] = ra[e];
#pragma unroll
                    for (int e = 0; e < B_PER; e++) Bs[buf * KC * BS + b_k * BS + b_c0 + BT * e] = rbv[e];
                };
                load_chunk(0);
                store_chunk(0);
                __syncthreads();
                constexpr int NCH = n / KC;
#pragma unroll 1
                for (int ch = 0; ch < NCH; ch++) {
                    const int buf = ch & 1;
                    if (ch + 1 < NCH) load_chunk((ch + 1) * KC);
                    const unsigned a0 = sA + 8u * (unsigned)(buf * KC * AS), b0 = sB + 8u * (unsigned)(buf * KC * BS);
#pragma unroll
                    for (int ks = 0; ks < KC / 4; ks++) {
                        double a[4];
#pragma unroll
                        for (int tr = 0; tr < 4; tr++) a[tr] = sj_lds(a0 + 8u * (unsigned)((4 * ks + t) * AS + 32 * rb + 8 * tr + g));
#pragma unroll
                        for (int j = 0; j < TJ; j++) {
                            const double b = sj_lds(b0 + 8u * (unsigned)((4 * ks + t) * BS + cg * PW + 8 * j + g));
#pragma unroll
                            for (int tr = 0; tr < 4; tr++) dmma_acc(acc[tr][j], a[tr], b);
                        }
                    }
                    if (ch + 1 < NCH) store_chunk(buf ^ 1);
                    __syncthreads();
                }
                // epilogue: rows 32 rb + 8 tr + g, columns colbase + cg PW + 8 j + 2t, +1 of [A_R | B_R]
#pragma unroll
                for (int tr = 0; tr < 4; tr++) {
                    const int row = 32 * rb + 8 * tr + g;
#pragma unroll
                    for (int j = 0; j < TJ; j++) {
                        const int cfull = colbase + cg * PW + 8 * j + 2 * t, part = cfull / n, c = cfull % n;
                        if constexpr (!last) {
                            *reinterpret_cast<double2*>(Sst + (size_t)SH::slot(R) * 2 * nn + (size_t)row * 2 * n + cfull) =
                                make_double2(acc[tr][j][0], acc[tr][j][1]);
                        } else {
                            // L = -I - h sum_q b_q A_q ,  R = I - h sum_q b_q B_q
                            double s0 = TB::b(R) * acc[tr][j][0], s1 = TB::b(R) * acc[tr][j][1];
                            lr_accum<R, 0>(J, Sst, part, row, c, s0, s1);
                            const double sg = part ? 1.0 : -1.0;
                            const double o0 = (row == c ? sg : 0.0) - h * s0, o1 = (row == c + 1 ? sg : 0.0) - h * s1;
                            *reinterpret_cast<double2*>((part ? Rout : Lout) + (size_t)row * n + c) = make_double2(o0, o1);
                        }
                    }
                }
            }
            __syncthreads();  // S_R is complete (global writes + block barrier) before a later stage reads it
        }
    }
    template <int R> __device__ __forceinline__ static void all_stages(const double* J, double* Sst, double h, double* As, double* Bs,
                                                                       double* Lout, double* Rout) {
        if constexpr (R < s) {
            do_stage<R>(J, Sst, h, As, Bs, Lout, Rout);
            all_stages<R + 1>(J, Sst, h, As, Bs, Lout, Rout);
        }
    }
};

template <int n, int ORDER>
__global__ void __launch_bounds__(256, 1)
k_chain_gemm(int i_first, int N, const double* __restrict__ mesh, double* __restrict__ scratch, double* __restrict__ Lb,
             double* __restrict__ Rb) {
    using CGm = ChainGemm<n, ORDER>;
    static_assert(n == 64 || n == 128, "dense chain-rule products: n = 64 or 128");
    static_assert(CGm::SH::last_implicit() >= 0, "the tableau needs at least one implicit stage");
    extern __shared__ __align__(16) double cg_smem[];
    const int i = i_first + blockIdx.x;
    if (i >= N - 1) return;
    const double h = mesh[i + 1] - mesh[i];
    double* base = scratch + (size_t)blockIdx.x * CGm::per;
    CGm::template all_stages<0>(base, base + (size_t)CGm::s * CGm::nn, h, cg_smem, cg_smem + 2 * CGm::KC * CGm::AS,
                                Lb + (size_t)i * CGm::nn, Rb + (size_t)i * CGm::nn);
}

}  // namespace mirk
