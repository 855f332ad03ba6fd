// ensemble_warp.cuh — K6 rebuilt around ON-CHIP state: ONE WARP PER TRAJECTORY, the whole adaptive solve of a
// small BVP (n = 2 states: the pendulum sweep of BASELINE config C3 and the reference's 2-state test problems) with
// mesh, iterate, stages and elimination factors in SHARED MEMORY.  HBM sees a trajectory's parameters on the way in
// and its final mesh / solution / outcome on the way out (~1.5 KB), nothing in between.
//
// Same adaptive loop as the single-problem driver (mirk_solve in mirk_b200.cu; reference:
// lib/BoundaryValueDiffEqMIRK/src/mirk.jl:286-388) and the thread-per-trajectory kernel of ensemble.cuh, other
// mapping:
//   * the 32 lanes share the MESH: lane l owns a contiguous chunk of intervals (chunks never straddle a node the
//     boundary conditions touch).  One evaluation of the collocation residual on DualN<2n> (multidual.cuh) per
//     interval yields Phi_i, the stages and all of [L_i R_i]; the lane merges its chunk into one relation right
//     away (row-pivoted elimination of the stacked blocks, as everywhere in this library);
//   * the per-lane relations are collapsed by a SHUFFLE TREE over the lanes of a segment (log2 depth, relation =
//     2n^2 + n doubles through __shfl_down), factors of every eliminated node go to shared memory;
//   * closing system on the boundary nodes: lane per row (warp_dense_solve32); back substitution walks the tree
//     top down, then every lane its own chunk; the update y -= delta is fused;
//   * defect estimate, the s_hat powers and the re-interpolation are parallel over intervals / new nodes; the two
//     order-dependent pieces (the sums that decide the new mesh size through round(), the equidistribution sweep) are
//     done by one lane in the reference's order.
//   Warps fetch trajectories from a global work counter (trajectories differ in iterations and mesh growth).
//   A trajectory whose mesh outgrows the on-chip capacity is reported in an overflow list; the host re-runs only
//   those through the HBM-slab kernel of ensemble.cuh with a capacity of max_num_subintervals + 1 nodes.
#pragma once
#include "warp_dense.cuh"
#include "ensemble.cuh"
#include "multidual.cuh"

namespace mirk {

constexpr int MIRK_ENS_OVERFLOW = -100;  // internal: the mesh outgrew the on-chip capacity

struct EnsWarpArgs {
    EnsArgs a;                          // shared scalar options and the per-trajectory result arrays
    int NCs;                            // on-chip node capacity per warp
    unsigned long long* counter;        // next trajectory to fetch
    unsigned long long* overflow_count;
    long long* overflow_list;           // trajectories to re-run with the HBM-slab kernel
    double* out_mesh;                   // [ntraj][NCs]
    double* out_y;                      // [ntraj][NCs][n]
    double* y_first;                    // [ntraj][n]
    const long long* idx;               // optional: work item t is global trajectory idx[t] (re-runs of overflowed ones)
    long long nwork;                    // work items (a.ntraj when idx is null)
};

template <class P, int ORDER> struct EnsWarpLayout {
    using TB = Tableau<ORDER>;
    static constexpr int n = P::n, s = TB::s, si = TB::si;
    static constexpr int FAC = 2 * n * n + n;                       // TL | TR | rt of an eliminated node
    static constexpr int SCR = FAC > si * n ? FAC : si * n;          // factors during Newton, interpolation stages after it
    static constexpr int QMAX = P::max_bc_pts + 2, DMAX = QMAX * n;
    // per-node doubles: mesh, mesh2, y, y2, Kd, phi/delta, est, scr
    static constexpr int per_node = 2 + 2 * n + s * n + n + 1 + SCR;
    static constexpr int fixed = DMAX * (DMAX + 1) + 8;             // closing matrix + kept nodes / status (as ints)
    __host__ __device__ static constexpr size_t warp_doubles(int NC) { return (size_t)per_node * NC + fixed; }
};

// NaN-propagating max over the warp
__device__ __forceinline__ double warp_nmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = !(w <= v) ? w : v;
    }
    return v;
}

// out-of-line pieces (one copy per kernel): the whole-solve kernel is bounded by its code size
template <int ORDER> static __device__ __noinline__ void ens_weights(double tau, double* w, double* wp) { Tableau<ORDER>::weights(tau, w, wp); }
template <class P, int ORDER>
static __device__ __noinline__ void ens_interp_stages(const double* yi, const double* yi1, double h, double ti, const double* p,
                                                      const double* K, double* KI) {
    interp_stages_interval<P, ORDER>(yi, yi1, h, ti, p, K, KI);
}
static __device__ __noinline__ double ens_pow(double x, double e) { return pow(x, e); }

// Phi_i, the stages and the blocks [L_i R_i] of one interval by STAGE-WISE Jacobians and the chain rule:
//     J_r = df/dy(Y_r)                                  one evaluation of f on DualN<n> seeded with the identity
//     A_r = dK_r/dy_i     = J_r [(1 - v_r) I + h sum_{j<r} x_rj A_j]
//     B_r = dK_r/dy_{i+1} = J_r [     v_r  I + h sum_{j<r} x_rj B_j]
//     L_i = -I - h sum_r b_r A_r ,   R_i = I - h sum_r b_r B_r
// instead of one sweep on DualN<2n>: half the tangents through f (and no products with the structural zeros of the
// seeds: K_1 does not depend on y_{i+1}, K_2 not on y_i, ...).  Values (stages, Phi) are computed in plain double in the
// order of phi_interval<double>, so they are bit-identical to the residual kernels'.
template <class P, int ORDER>
__device__ __forceinline__ void phi_interval_blocks(const double* __restrict__ yi, const double* __restrict__ yi1, double h, double ti,
                                                    const double* __restrict__ p, double (*K)[P::n], double* __restrict__ phi,
                                                    double (*L2)[P::n], double (*R2)[P::n]) {
    using TB = Tableau<ORDER>;
    constexpr int n = P::n, s = TB::s;
    using DN = DualN<n>;
    double A[s][n][n], B[s][n][n];
#pragma unroll
    for (int r = 0; r < s; r++) {
        const double vr = TB::v(r);
        double val[n], MA[n][n], MB[n][n];
#pragma unroll
        for (int k = 0; k < n; k++) {
            if (vr == 0.0) val[k] = yi[k];
            else if (vr == 1.0) val[k] = yi1[k];
            else val[k] = (1.0 - vr) * yi[k] + vr * yi1[k];
#pragma unroll
            for (int d = 0; d < n; d++) { MA[k][d] = k == d ? 1.0 - vr : 0.0; MB[k][d] = k == d ? vr : 0.0; }
        }
#pragma unroll
        for (int j = 0; j < r; j++) {
            const double xrj = TB::x(r, j);
            if (xrj != 0.0) {
                const double hx = h * xrj;
#pragma unroll
                for (int k = 0; k < n; k++) {
                    val[k] = val[k] + hx * K[j][k];
#pragma unroll
                    for (int d = 0; d < n; d++) { MA[k][d] = fma(hx, A[j][k][d], MA[k][d]); MB[k][d] = fma(hx, B[j][k][d], MB[k][d]); }
                }
            }
        }
        DN Y[n], Kr[n];
#pragma unroll
        for (int k = 0; k < n; k++) Y[k] = DN::seed(val[k], k);
        P::template f<DN>(Kr, Y, p, ti + TB::c(r) * h);
        if constexpr (HasSingular<P>::value) {
            const double tt = ti + TB::c(r) * h;
            if (tt > 0.0) {
                DN sv[n];
                P::template singular<DN>(sv, Y, p);
                const double it = 1.0 / tt;
#pragma unroll
                for (int k = 0; k < n; k++) Kr[k] = Kr[k] + it * sv[k];
            }
        }
#pragma unroll
        for (int k = 0; k < n; k++) {
            K[r][k] = Kr[k].v;
#pragma unroll
            for (int d = 0; d < n; d++) {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int e = 0; e < n; e++) { a = fma(Kr[k].d[e], MA[e][d], a); b = fma(Kr[k].d[e], MB[e][d], b); }
                A[r][k][d] = a;
                B[r][k][d] = b;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < n; k++) {
        double acc = TB::b(0) * K[0][k];
#pragma unroll
        for (int r = 1; r < s; r++) acc = acc + TB::b(r) * K[r][k];
        phi[k] = yi1[k] - yi[k] - h * acc;
#pragma unroll
        for (int d = 0; d < n; d++) {
            double la = TB::b(0) * A[0][k][d], lb = TB::b(0) * B[0][k][d];
#pragma unroll
            for (int r = 1; r < s; r++) { la = fma(TB::b(r), A[r][k][d], la); lb = fma(TB::b(r), B[r][k][d], lb); }
            L2[k][d] = (k == d ? -1.0 : 0.0) - h * la;
            R2[k][d] = (k == d ? 1.0 : 0.0) - h * lb;
        }
    }
}

template <class P, int ORDER> struct EnsWarpSolver {
    using TB = Tableau<ORDER>;
    using LY = EnsWarpLayout<P, ORDER>;
    static constexpr int n = P::n, s = TB::s, si = TB::si, L = P::n_bc, FAC = LY::FAC, SCR = LY::SCR, QMAX = LY::QMAX,
                         DMAX = LY::DMAX, ND = 2 * n, rows = 2 * n, cols = 3 * n + 1, LV = 5;
    using MD = DualN<ND>;
    static_assert(QMAX + 2 + P::max_bc_pts <= 16, "the integer scratch of a warp holds 16 entries");
    static_assert(!BcUsesDerivative<P>::value, "boundary conditions on sol(t, Val{1}) are not offered by the ensemble kernels");

    // shared-memory views of this warp
    double *mesh, *mesh2, *y, *y2, *Kd, *phi, *est, *scr, *M;
    int* ikept;  // [QMAX] kept nodes, [QMAX] = Q, [QMAX + 1] = status, then bc nodes
    int NC, N, lane;
    double p[P::np > 0 ? P::np : 1];
    // partition of the intervals over the lanes (per mesh)
    int Q, pa, pb, pls, ple, psg, pchunk, plv;  // plv: tree levels the widest segment needs
    bool pact;
    // carried relation of this lane: A d_left + E d_right = r, and the tree history
    double E[n][n], A[n][n], r[n];

    __device__ __forceinline__ void bind(double* base, int NC_, int lane_) {
        NC = NC_; lane = lane_;
        mesh = base; mesh2 = mesh + NC; y = mesh2 + NC; y2 = y + (size_t)NC * n; Kd = y2 + (size_t)NC * n;
        phi = Kd + (size_t)NC * s * n; est = phi + (size_t)NC * n; scr = est + NC; M = scr + (size_t)NC * SCR;
        ikept = reinterpret_cast<int*>(M + DMAX * (DMAX + 1));
    }
    __device__ __forceinline__ int* bcnodes() { return ikept + QMAX + 2; }

    // ---- boundary evaluation points (lane 0): U[k] = sol(times[k]), end-point short cut; nodes[k] as k_bc --------
    __device__ __forceinline__ int bc_gather(double* U, int* nodes) const {
        double tm[P::max_bc_pts];
        int m;
        const double t0 = mesh[0], t1 = mesh[N - 1];
        if (P::problem_type == 1) { m = 2; tm[0] = t0; tm[1] = t1; }
        else m = P::bc_times(tm, p, t0, t1);
        for (int k = 0; k < m; k++) {
            const double t = tm[k];
            if (t == t0) {
                nodes[k] = 0;
                for (int c = 0; c < n; c++) U[k * n + c] = y[c];
            } else if (t == t1) {
                nodes[k] = N - 1;
                for (int c = 0; c < n; c++) U[k * n + c] = y[(size_t)(N - 1) * n + c];
            } else {
                const int i = interval_of(mesh, N, t);
                nodes[k] = i;
                const double ti = mesh[i], h = mesh[i + 1] - ti, tau = (t - ti) / h;
                double yi[n], yi1[n], KIl[si * n > 0 ? si * n : 1];
                for (int c = 0; c < n; c++) { yi[c] = y[(size_t)i * n + c]; yi1[c] = y[(size_t)(i + 1) * n + c]; }
                const double* K = Kd + (size_t)i * s * n;
                ens_interp_stages<P, ORDER>(yi, yi1, h, ti, p, K, KIl);
                double w[TB::s_star], wp[TB::s_star];
                ens_weights<ORDER>(tau, w, wp);
                for (int c = 0; c < n; c++) {
                    double z = 0.0;
#pragma unroll
                    for (int q = 0; q < s; q++) z += K[q * n + c] * w[q];
#pragma unroll
                    for (int q = 0; q < si; q++) z += KIl[q * n + c] * w[s + q];
                    U[k * n + c] = z * h + yi[c];
                }
            }
        }
        return m;
    }

    // kept nodes (never eliminated) of the current mesh and the lane partition; call after every mesh change
    __device__ __forceinline__ void plan() {
        if (lane == 0) {
            double tm[P::max_bc_pts];
            int m, bcn[P::max_bc_pts];
            const double t0 = mesh[0], t1 = mesh[N - 1];
            if (P::problem_type == 1) { m = 2; tm[0] = t0; tm[1] = t1; }
            else m = P::bc_times(tm, p, t0, t1);
            for (int k = 0; k < m; k++) bcn[k] = tm[k] == t0 ? 0 : tm[k] == t1 ? N - 1 : interval_of(mesh, N, tm[k]);
            int kept[QMAX], q = 0;
            kept[q++] = 0;
            for (int k = 0; k < m; k++) {
                const int v = bcn[k];
                bool found = false;
                for (int e = 0; e < q; e++) found = found || kept[e] == v;
                if (!found && v != N - 1) kept[q++] = v;
            }
            kept[q++] = N - 1;
            for (int e = 1; e < q - 1; e++)
                for (int f = e + 1; f < q - 1; f++)
                    if (kept[f] < kept[e]) { const int t_ = kept[e]; kept[e] = kept[f]; kept[f] = t_; }
            for (int e = 0; e < q; e++) ikept[e] = kept[e];
            ikept[QMAX] = q;
        }
        __syncwarp();
        Q = ikept[QMAX];
        const int ni = N - 1;
        int C = (ni + 31) / 32;
        for (;;) {
            int tot = 0;
            for (int sg = 0; sg < Q - 1; sg++) tot += (ikept[sg + 1] - ikept[sg] + C - 1) / C;
            if (tot <= 32) break;
            C++;
        }
        pact = false; psg = -1; pls = ple = 0; pa = pb = 0;
        pchunk = C;
        plv = 0;
        int off = 0;
        for (int sg = 0; sg < Q - 1; sg++) {
            const int len = ikept[sg + 1] - ikept[sg], nl = (len + C - 1) / C;
            while ((1 << plv) < nl) plv++;
            if (lane >= off && lane < off + nl) {
                const int j = lane - off;
                psg = sg; pls = off; ple = off + nl; pact = true;
                pa = ikept[sg] + j * C;
                pb = ikept[sg] + (j + 1) * C;
                if (pb > ikept[sg + 1]) pb = ikept[sg + 1];
            }
            off += nl;
        }
    }

    // ---- merge: eliminate the node shared by the carried relation (E, A, r) and the incoming one (L2, R2, r2) -------
    // factors of that node to fac = [TL | TR | rt]; returns false on a singular pivot
    __device__ __forceinline__ bool merge(const double (&L2)[n][n], const double (&R2)[n][n], const double (&r2)[n], double* fac) {
        double W[rows][cols];
#pragma unroll
        for (int q = 0; q < n; q++) {
#pragma unroll
            for (int k = 0; k < n; k++) {
                W[q][k] = E[q][k]; W[q][n + k] = A[q][k]; W[q][2 * n + k] = 0.0;
                W[n + q][k] = L2[q][k]; W[n + q][n + k] = 0.0; W[n + q][2 * n + k] = R2[q][k];
            }
            W[q][3 * n] = r[q];
            W[n + q][3 * n] = r2[q];
        }
#pragma unroll
        for (int q = 0; q < n; q++) {
            int pr = q;
            double best = fabs(W[q][q]);
#pragma unroll
            for (int rr = q + 1; rr < rows; rr++) {
                const double av = fabs(W[rr][q]);
                if (av > best || !(av == av)) { best = av; pr = rr; }
            }
            if (!(best > 0.0) || !(best < INFINITY)) return false;
#pragma unroll
            for (int rr = q + 1; rr < rows; rr++) {
                if (pr == rr) {
#pragma unroll
                    for (int c = q; c < cols; c++) { const double t_ = W[q][c]; W[q][c] = W[rr][c]; W[rr][c] = t_; }
                }
            }
            const double inv = fast_rcp(W[q][q]);  // (IEEE division would add ~40 instructions per pivot to a hot, small loop)
#pragma unroll
            for (int c = q + 1; c < cols; c++) W[q][c] *= inv;
#pragma unroll
            for (int rr = 0; rr < rows; rr++) {
                if (rr != q) {
                    const double mlt = W[rr][q];
#pragma unroll
                    for (int c = q + 1; c < cols; c++) W[rr][c] = fma(-mlt, W[q][c], W[rr][c]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < n; q++) {
#pragma unroll
            for (int k = 0; k < n; k++) { fac[q * n + k] = W[q][n + k]; fac[n * n + q * n + k] = W[q][2 * n + k]; }
            fac[2 * n * n + q] = W[q][3 * n];
        }
#pragma unroll
        for (int q = 0; q < n; q++) {
#pragma unroll
            for (int k = 0; k < n; k++) { E[q][k] = W[n + q][2 * n + k]; A[q][k] = W[n + q][n + k]; }
            r[q] = W[n + q][3 * n];
        }
        return true;
    }

    // ---- one pass over the mesh at the current iterate: F(y) (stages stored), |F|_inf, and the complete forward
    //      elimination (chunk merges, shuffle tree, closing matrix assembled).  *ok = false on a singular block.
    __device__ __forceinline__ double sweep(bool* ok) {
        bool good = true;
        double nrm = 0.0;
        for (int e = lane; e < DMAX * (DMAX + 1); e += 32) M[e] = 0.0;
        // ONE loop over "leaf" steps (the intervals of this lane's chunk) and "tree" steps (the relation of the lane
        // `stride` to the right, by shuffle), so that the residual evaluation and the merge are inlined once each:
        // the kernel's code size is what bounds it (instruction-cache misses were the top stall of the first version)
        const int cmax = __reduce_max_sync(0xffffffffu, pact ? pb - pa : 0);
        int cur_b = pb;
#pragma unroll 1
        for (int step = 0; step < cmax + plv; step++) {
            double L2[n][n], R2[n][n], r2[n];
            bool act, first = false;
            if (step < cmax) {  // warp-uniform
                const int i = pa + step;
                act = pact && i < pb;
                first = step == 0;
                if (act) {
                    const double ti = mesh[i], h = mesh[i + 1] - ti;
                    double yi[n], yi1[n], K[s][n];
#pragma unroll
                    for (int k = 0; k < n; k++) { yi[k] = y[(size_t)i * n + k]; yi1[k] = y[(size_t)(i + 1) * n + k]; }
                    phi_interval_blocks<P, ORDER>(yi, yi1, h, ti, p, K, r2, L2, R2);
#pragma unroll
                    for (int q = 0; q < s; q++)
#pragma unroll
                        for (int k = 0; k < n; k++) Kd[((size_t)i * s + q) * n + k] = K[q][k];
#pragma unroll
                    for (int k = 0; k < n; k++) nrm = nmax(nrm, r2[k]);
                }
            } else {
                const int stride = 1 << (step - cmax);
                // the partner's relation over [p_a, p_b]: its A multiplies d(p_a) (our shared node), its E d(p_b)
#pragma unroll
                for (int q = 0; q < n; q++) {
#pragma unroll
                    for (int k = 0; k < n; k++) {
                        R2[q][k] = __shfl_down_sync(0xffffffffu, E[q][k], stride);
                        L2[q][k] = __shfl_down_sync(0xffffffffu, A[q][k], stride);
                    }
                    r2[q] = __shfl_down_sync(0xffffffffu, r[q], stride);
                }
                const int p_b = __shfl_down_sync(0xffffffffu, cur_b, stride);
                act = pact && ((lane - pls) % (2 * stride) == 0) && (lane + stride < ple);
                if (act) cur_b = p_b;
            }
            if (act) {
                if (first) {
#pragma unroll
                    for (int q = 0; q < n; q++) {
#pragma unroll
                        for (int k = 0; k < n; k++) { E[q][k] = R2[q][k]; A[q][k] = L2[q][k]; }
                        r[q] = r2[q];
                    }
                } else {
                    // the eliminated node: leaf step -> node i = pa + step; tree step -> first node of the partner's range
                    const int c = step < cmax ? pa + step : tree_node(step - cmax);
                    good = merge(L2, R2, r2, scr + (size_t)c * SCR) && good;
                }
            }
        }
        __syncwarp();  // every interval's stages are in shared memory: the boundary rows may interpolate
        // boundary rows (lane 0): residual and reference-pattern Jacobian (quirk Q2) into rows [0, L) of M
        if (lane == 0) {
            double U[P::max_bc_pts * n], rbc[L];
            int* bcn = bcnodes();
            const int m = bc_gather(U, bcn);
            P::template bc<double>(rbc, U, p);
            const int D = Q * n;
            for (int q = 0; q < L; q++) { nrm = nmax(nrm, rbc[q]); M[q * (DMAX + 1) + D] = rbc[q]; }
#pragma unroll 1
            for (int d = 0; d < m * n; d++) {
                Dual Ud[P::max_bc_pts * n], rd[L];
                for (int e = 0; e < m * n; e++) Ud[e] = Dual(U[e], e == d ? 1.0 : 0.0);
                P::template bc<Dual>(rd, Ud, p);
                const int k = d / n, c = d % n;
                int slot = 0;
                for (int e = 0; e < Q; e++) if (ikept[e] == bcn[k]) slot = e;
                for (int q = 0; q < L; q++) M[q * (DMAX + 1) + slot * n + c] += rd[q].d;
            }
        }
        // collapsed relation of a segment -> rows L + sg n + q of the closing matrix
        if (pact && lane == pls) {
            const int D = Q * n;
#pragma unroll
            for (int q = 0; q < n; q++) {
                double* row = M + (size_t)(L + psg * n + q) * (DMAX + 1);
#pragma unroll
                for (int k = 0; k < n; k++) { row[psg * n + k] = A[q][k]; row[(psg + 1) * n + k] = E[q][k]; }
                row[D] = r[q];
            }
        }
        __syncwarp();
        *ok = __all_sync(0xffffffffu, good);
        return warp_nmax(nrm);
    }
    // tree level lv: this lane (a receiver) eliminates the first node of the range of lane + 2^lv; after the merge its
    // range ends where the range of lane + 2^(lv+1) starts (or at the end of the segment)
    __device__ __forceinline__ int tree_node(int lv) const { return lane_start(lane + (1 << lv)); }
    __device__ __forceinline__ int tree_right(int lv) const {
        const int nxt = lane + (2 << lv);
        return nxt < ple ? lane_start(nxt) : ikept[psg + 1];
    }
    // first node of lane l's chunk (l in this lane's segment): chunks are `pchunk` intervals long
    __device__ __forceinline__ int lane_start(int l) const { return ikept[psg] + (l - pls) * pchunk; }

    // closing solve, back substitution (tree top down, then the chunks) and y -= delta.  Needs sweep()'s state.
    __device__ __forceinline__ bool solve_and_update() {
        const int D = Q * n;
        double* delta = phi;  // Phi_i was consumed by sweep()
        // compact the closing matrix rows to pitch D + 1 is not needed: warp_dense_solve32 takes the pitch
        ikept[QMAX + 1] = 0;
        __syncwarp();
        warp_dense_solve32(M, D, DMAX + 1, ikept, n, delta, ikept + QMAX + 1);
        __syncwarp();
        if (ikept[QMAX + 1] != 0) return false;
#pragma unroll 1
        for (int lv = plv - 1; lv >= 0; lv--) {
            const int stride = 1 << lv;
            if (pact && ((lane - pls) % (2 * stride) == 0) && (lane + stride < ple)) {
                const int c = tree_node(lv), b = tree_right(lv);
                const double* fac = scr + (size_t)c * SCR;
#pragma unroll
                for (int q = 0; q < n; q++) {
                    double acc = fac[2 * n * n + q];
#pragma unroll
                    for (int k = 0; k < n; k++)
                        acc -= fac[q * n + k] * delta[(size_t)pa * n + k] + fac[n * n + q * n + k] * delta[(size_t)b * n + k];
                    delta[(size_t)c * n + q] = acc;
                }
            }
            __syncwarp();
        }
        if (pact) {
            double da[n], dr[n];
#pragma unroll
            for (int k = 0; k < n; k++) { da[k] = delta[(size_t)pa * n + k]; dr[k] = delta[(size_t)pb * n + k]; }
            for (int c = pb - 1; c > pa; c--) {
                const double* fac = scr + (size_t)c * SCR;
                double dc[n];
#pragma unroll
                for (int q = 0; q < n; q++) {
                    double acc = fac[2 * n * n + q];
#pragma unroll
                    for (int k = 0; k < n; k++) acc -= fac[q * n + k] * da[k] + fac[n * n + q * n + k] * dr[k];
                    dc[q] = acc;
                }
#pragma unroll
                for (int k = 0; k < n; k++) { delta[(size_t)c * n + k] = dc[k]; dr[k] = dc[k]; }
            }
        }
        __syncwarp();
        for (int e = lane; e < N * n; e += 32) y[e] -= delta[e];
        __syncwarp();
        return true;
    }

    // ---- the adaptive outer loop of one trajectory; returns the ReturnCode (or MIRK_ENS_OVERFLOW) ----------------
    __device__ int run(const EnsArgs& a, int* newton_out, int* outer_out, double* resid_out, double* defect_out) {
        const double abstol = a.abstol;
        int info = MIRK_RET_SUCCESS_, newton_total = 0, outer = 0;
        double error_norm = 2.0 * abstol, resid_norm = 0.0;
        do {
            plan();
            // -- Newton solve with best-iterate bookkeeping (same control flow as newton_solve in mirk_b200.cu)
            // (one call site of sweep() and of solve_and_update(): both are large once inlined)
            int ret = MIRK_RET_MAXITERS_, it = 0;
            double best = INFINITY, nrm = 0.0;
            bool have_best = false, ok = true, restoring = false;
            for (;;) {
                nrm = sweep(&ok);
                if (restoring) break;
                if (it > 0) {  // the checks that follow a Newton step
                    if (!(fabs(nrm) < INFINITY)) {
                        ret = MIRK_RET_UNSTABLE_;
                    } else {
                        if (nrm < best) {
                            best = nrm;
                            have_best = true;
                            for (int e = lane; e < N * n; e += 32) y2[e] = y[e];
                        }
                        if (nrm <= abstol) ret = MIRK_RET_SUCCESS_;
                    }
                }
                bool stop = (it > 0 && ret != MIRK_RET_MAXITERS_) || it >= a.maxiters;
                if (!stop) {
                    if (!ok || !solve_and_update()) { ret = MIRK_RET_FAILURE_; stop = true; }
                    else { it++; continue; }
                }
                if (ret != MIRK_RET_SUCCESS_ && it > 0 && have_best && ret != MIRK_RET_FAILURE_) {
                    __syncwarp();
                    for (int e = lane; e < N * n; e += 32) y[e] = y2[e];
                    __syncwarp();
                    restoring = true;
                    continue;
                }
                break;
            }
            if (ret != MIRK_RET_SUCCESS_ && a.nlsolve == 0) return MIRK_ENS_NEEDS_POLY;  // host re-runs it with the fallbacks
            resid_norm = nrm;
            newton_total += it;
            error_norm = 2.0 * abstol;
            info = ret;
            outer++;
            if (!a.adaptive) break;
            if (info == MIRK_RET_SUCCESS_) {
                // -- defect estimate (Appendix A.5), parallel over the intervals; interpolation stages kept in scr
                double defect = 0.0;
                for (int i = lane; i < N - 1; i += 32) {
                    const double ti = mesh[i], h = mesh[i + 1] - ti;
                    double yi[n], yi1[n], KIl[si * n > 0 ? si * n : 1];
#pragma unroll
                    for (int k = 0; k < n; k++) { yi[k] = y[(size_t)i * n + k]; yi1[k] = y[(size_t)(i + 1) * n + k]; }
                    const double* Kl = Kd + (size_t)i * s * n;
                    ens_interp_stages<P, ORDER>(yi, yi1, h, ti, p, Kl, KIl);
#pragma unroll
                    for (int e = 0; e < si * n; e++) scr[(size_t)i * SCR + e] = KIl[e];
                    double e12[2];
#pragma unroll 1
                    for (int smp = 0; smp < 2; smp++) {
                        const double tau = smp ? (1.0 - TB::tau_star()) : TB::tau_star();
                        double w[TB::s_star], wp[TB::s_star], z[n], zp[n], g[n];
                        ens_weights<ORDER>(tau, w, wp);
#pragma unroll
                        for (int k = 0; k < n; k++) {
                            double za = 0.0, zb = 0.0;
#pragma unroll
                            for (int q = 0; q < s; q++) { za += Kl[q * n + k] * w[q]; zb += Kl[q * n + k] * wp[q]; }
#pragma unroll
                            for (int q = 0; q < si; q++) { za += KIl[q * n + k] * w[s + q]; zb += KIl[q * n + k] * wp[s + q]; }
                            z[k] = za * h + yi[k];
                            zp[k] = zb;
                        }
                        P::template f<double>(g, z, p, ti + tau * h);
                        double e = 0.0;
                        bool isnan_ = false;
#pragma unroll
                        for (int k = 0; k < n; k++) {
                            const double dd = (zp[k] - g[k]) / (fabs(g[k]) + 1.0);
                            if (fabs(dd) > e) e = fabs(dd);
                            isnan_ = isnan_ || !(dd == dd);
                        }
                        if (smp) e12[1] = isnan_ ? NAN : e; else e12[0] = isnan_ ? NAN : e;
                    }
                    const double em = (e12[0] > e12[1]) ? e12[0] : e12[1];
                    est[i] = em;
                    defect = !(em <= defect) ? em : defect;
                }
                error_norm = warp_nmax(defect);
                __syncwarp();
                if (!(error_norm <= a.defect_threshold)) info = MIRK_RET_FAILURE_;
                if (info == MIRK_RET_SUCCESS_ && error_norm > abstol) {
                    // -- mesh selection (Appendix A.6): powers in parallel, the order-dependent sums by lane 0
                    const int ni = N - 1;
                    const double ex = 1.0 / (double)((ORDER == 7 ? 6 : ORDER) + 1);
                    for (int i = lane; i < ni; i += 32) est[i] = ens_pow(est[i] / abstol, ex);
                    __syncwarp();
                    int ns = 0, halve = 0;
                    if (lane == 0) {
                        double r1 = 0.0, r2 = 0.0;
                        for (int i = 0; i < ni; i++) {
                            const double sh = est[i];
                            if (sh > r1) r1 = sh;
                            r2 += sh;
                        }
                        const double r3 = r2 / ni;
                        long long n_predict = (long long)nearbyint(1.3 * r2 + 1.0);
                        const double n_ = 0.1 * ni;
                        if (fabs((double)(n_predict - ni)) < n_) n_predict = (long long)nearbyint(ni + n_);
                        if (r1 <= 1.0 * r3) { ns = 2 * ni; halve = 1; }
                        else {
                            const long long lb = N / 2, ub = 4LL * ni;
                            ns = (int)(n_predict < lb ? lb : (n_predict > ub ? ub : n_predict));
                        }
                    }
                    ns = __shfl_sync(0xffffffffu, ns, 0);
                    halve = __shfl_sync(0xffffffffu, halve, 0);
                    if (ns > a.max_sub) {
                        info = MIRK_RET_FAILURE_;
                    } else if (ns + 1 > NC) {
                        return MIRK_ENS_OVERFLOW;
                    } else {
                        if (halve) {
                            for (int i = lane; i < N; i += 32) {
                                mesh2[2 * i] = mesh[i];
                                if (i < ni) mesh2[2 * i + 1] = (mesh[i + 1] + mesh[i]) / 2.0;
                            }
                        } else {
                            const double tend = mesh[ni];
                            for (int i = lane; i <= ns; i += 32) mesh2[i] = (i < N) ? mesh[i] : tend;
                            __syncwarp();
                            if (lane == 0) {
                                double tot = 0.0;
                                for (int i = 0; i < ni; i++) {
                                    const double hh = mesh[i + 1] - mesh[i];
                                    const double sh = est[i] / hh;
                                    est[i] = sh;
                                    tot += sh * hh;
                                }
                                const double zeta = tot / (double)ns;
                                int k = 0;
                                long long i = 0;
                                double t = mesh[0], integral = 0.0;
                                mesh2[0] = t;
                                while (k < ni) {
                                    const double next_piece = est[k] * (mesh[k + 1] - t);
                                    const double int_next = integral + next_piece;
                                    if (int_next > zeta) {
                                        const double tn2 = (zeta - integral) / est[k] + t;
                                        if (i + 1 <= ns) mesh2[i + 1] = tn2;
                                        t = tn2;
                                        i++;
                                        integral = 0.0;
                                    } else {
                                        integral = int_next;
                                        t = mesh[k + 1];
                                        k++;
                                    }
                                }
                                mesh2[ns] = tend;
                            }
                        }
                        __syncwarp();
                        // -- new guess: old interpolant at the new nodes (Appendix A.7)
                        const int Nn = ns + 1;
                        // (quirk Q3, reinterp_inplace: the reference adds the base from the array it is rewriting —
                        //  that order dependence makes it a one-lane sequential loop; one call site either way)
                        const bool inpl = a.reinterp_inplace != 0;
#pragma unroll 1
                        for (int j = inpl ? 0 : lane; j < Nn; j += inpl ? 1 : 32)
                            if (!inpl || lane == 0) reinterp_node(j, inpl);
                        __syncwarp();
                        for (int j = lane; j < Nn; j += 32) mesh[j] = mesh2[j];
                        for (int e = lane; e < Nn * n; e += 32) y[e] = y2[e];
                        N = Nn;
                        __syncwarp();
                    }
                    continue;  // (a failed mesh selection ends the solve: mirk.jl:360-372 has no halving on that path)
                }
            }
            if (info != MIRK_RET_SUCCESS_) {
                // mirk.jl:374-385: halve the mesh, zero the guess, restart (quirk Q4)
                if (2 * (N - 1) > a.max_sub) {
                    info = MIRK_RET_FAILURE_;
                } else if (2 * (N - 1) + 1 > NC) {
                    return MIRK_ENS_OVERFLOW;
                } else {
                    const int ni = N - 1;
                    for (int i = lane; i < N; i += 32) {
                        mesh2[2 * i] = mesh[i];
                        if (i < ni) mesh2[2 * i + 1] = (mesh[i + 1] + mesh[i]) / 2.0;
                    }
                    __syncwarp();
                    N = 2 * ni + 1;
                    for (int j = lane; j < N; j += 32) mesh[j] = mesh2[j];
                    for (int e = lane; e < N * n; e += 32) y[e] = 0.0;
                    __syncwarp();
                    info = MIRK_RET_SUCCESS_;
                }
            }
        } while (info == MIRK_RET_SUCCESS_ && error_norm > abstol && outer < a.max_outer);
        if (info == MIRK_RET_SUCCESS_ && a.adaptive && error_norm > abstol) info = MIRK_RET_MAXITERS_;
        *newton_out = newton_total;
        *outer_out = outer;
        *resid_out = resid_norm;
        *defect_out = error_norm;
        return info;
    }

    // y2[j] = old interpolant at mesh2[j]
    __device__ __forceinline__ void reinterp_node(int j, bool inplace) {
        const double t = mesh2[j];
        const int i = interval_of(mesh, N, t);
        const double ti = mesh[i], h = mesh[i + 1] - ti, tau = (t - ti) / h;
        double w[TB::s_star], wp[TB::s_star];
        ens_weights<ORDER>(tau, w, wp);
#pragma unroll
        for (int k = 0; k < n; k++) {
            double z = 0.0;
#pragma unroll
            for (int q = 0; q < s; q++) z += Kd[((size_t)i * s + q) * n + k] * w[q];
#pragma unroll
            for (int q = 0; q < si; q++) z += scr[(size_t)i * SCR + q * n + k] * w[s + q];
            const double base = (inplace && i < j) ? y2[(size_t)i * n + k] : y[(size_t)i * n + k];
            y2[(size_t)j * n + k] = z * h + base;
        }
    }
};

constexpr int kEnsWarpsPerBlock = 4;

template <class P, int ORDER>
__global__ void __launch_bounds__(kEnsWarpsPerBlock * 32, 4)
k_ensemble_warp(EnsWarpArgs w) {  // blockDim.x / 32 warps per CTA (fewer at the large-capacity re-run stages)
    using ES = EnsWarpSolver<P, ORDER>;
    using LY = EnsWarpLayout<P, ORDER>;
    constexpr int n = P::n;
    extern __shared__ __align__(16) double esm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    ES S;
    S.bind(esm + (size_t)wib * LY::warp_doubles(w.NCs), w.NCs, lane);
    const EnsArgs& a = w.a;
    for (;;) {
        unsigned long long t = 0ull;
        if (lane == 0) t = atomicAdd(w.counter, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned long long)w.nwork) break;
        const long long tid = w.idx ? w.idx[t] : (long long)t;  // global trajectory: parameters in, outcomes out
#pragma unroll
        for (int k = 0; k < P::np; k++) S.p[k] = a.params[tid * P::np + k];
        S.N = a.N0;
        for (int i = lane; i < S.N; i += 32) {
            S.mesh[i] = a.mesh0[i];
#pragma unroll
            for (int k = 0; k < n; k++) S.y[(size_t)i * n + k] = a.u0[(a.u0_per_traj ? tid * n : 0) + k];
        }
        __syncwarp();
        int newton = 0, outer = 0;
        double rn = 0.0, dn = 0.0;
        const int info = S.run(a, &newton, &outer, &rn, &dn);
        __syncwarp();
        if (info == MIRK_ENS_NEEDS_POLY) {
            if (lane == 0) {
                a.retcode[tid] = MIRK_ENS_NEEDS_POLY;
                a.poly_list[atomicAdd(a.poly_count, 1ull)] = tid;
            }
        } else if (info == MIRK_ENS_OVERFLOW) {
            if (lane == 0) {
                const unsigned long long slot = atomicAdd(w.overflow_count, 1ull);
                w.overflow_list[slot] = tid;
                a.retcode[tid] = MIRK_ENS_OVERFLOW;
            }
        } else {
            if (lane == 0) {
                a.retcode[tid] = info;
                a.n_mesh[tid] = S.N;
                a.newton_iters[tid] = newton;
                a.outer_iters[tid] = outer;
                a.resid_norm[tid] = rn;
                a.defect_norm[tid] = dn;
            }
            double* om = w.out_mesh + (size_t)t * w.NCs;      // solution slot = work item
            double* oy = w.out_y + (size_t)t * w.NCs * n;
            for (int i = lane; i < S.N; i += 32) om[i] = S.mesh[i];
            for (int e = lane; e < S.N * n; e += 32) oy[e] = S.y[e];
            if (lane < n) w.y_first[tid * n + lane] = S.y[lane];
        }
        __syncwarp();
    }
}

}  // namespace mirk
