// ops.cuh — per-(problem, order) launch table.  The host driver (mirk_b200.cu) only sees this
// table, so ahead-of-time instantiations (the built-in registry) and, later, NVRTC-compiled user
// functors plug in the same way.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "problems.cuh"
#include "stagejac.cuh"

namespace mirk {

struct ProblemOps {
    const char* name;
    int order, n, np, n_bc, n_bca, problem_type, max_bc_pts, s, s_star;
    void (*residual)(cudaStream_t, int N, const double* mesh, const double* y, const double* p,
                     double* Kd, double* phi_out, unsigned long long* norm_bits);
    void (*bc)(cudaStream_t, int N, const double* mesh, const double* y, const double* p,
               const double* Kd, double* Ki, double* resid, int* bc_nodes, double* Bc, int* m_out,
               unsigned long long* norm_bits, int want_jac);
    void (*jac_blocks)(cudaStream_t, int N, const double* mesh, const double* y, const double* p,
                       double* Lb, double* Rb);
    void (*resjac)(cudaStream_t, int N, const double* mesh, const double* y, const double* p, double* Kd,
                   double* phi_out, unsigned long long* norm_bits, double* Lb, double* Rb, double* scratch);
    // doubles of scratch resjac wants for a mesh of N nodes (0: none) — the stage-wise dense path of large n
    size_t (*resjac_scratch_doubles)(int N);
    void (*defect)(cudaStream_t, int N, const double* mesh, const double* y, const double* p,
                   const double* Kd, double* Ki, double* errors, double* est,
                   unsigned long long* defect_bits);
    void (*interp_setup)(cudaStream_t, int N, const double* mesh, const double* y, const double* p,
                         const double* Kd, double* Ki);
    // host twin of k_bc's node selection: the nodes the boundary rows touch (never eliminated)
    int (*bc_nodes_host)(int N, const double* mesh, const double* p, int* nodes);
};

// shape of the taped Jacobian kernel (one warp per CTA by default): intervals per CTA chosen so the grid is a
// whole number of FULL waves of resident CTAs — at C2's size 2500 CTAs of 8 intervals are 2.11 waves (the third
// round runs 11 % full), 1177 CTAs of 17 intervals are one wave.  `slots` = resident CTAs on the device.
inline int tape_intervals_per_cta(int intervals, int slots) {
    if (slots < 1) slots = 1;
    const int rounds = (intervals + 32 * slots - 1) / (32 * slots);  // waves needed at the maximum of 32 per CTA
    int ipw = (intervals + slots * rounds - 1) / (slots * rounds);
    if (ipw < 4) ipw = 4;  // short meshes: keep a few intervals per recording pass
    return ipw > 32 ? 32 : ipw;
}

template <class P, int ORDER> struct OpsImpl {
    static void residual(cudaStream_t st, int N, const double* mesh, const double* y, const double* p,
                         double* Kd, double* phi_out, unsigned long long* nb) {
        const int nb_ = (N - 1 + 127) / 128;
        k_residual<P, ORDER><<<nb_, 128, 0, st>>>(N, mesh, y, p, Kd, phi_out, nb);
    }
    static void bc(cudaStream_t st, int N, const double* mesh, const double* y, const double* p,
                   const double* Kd, double* Ki, double* resid, int* bc_nodes, double* Bc, int* m_out,
                   unsigned long long* nb, int want_jac) {
        k_bc<P, ORDER><<<1, 256, 0, st>>>(N, mesh, y, p, Kd, Ki, resid, bc_nodes, Bc, m_out, nb, want_jac);
    }
    static void jac_blocks(cudaStream_t st, int N, const double* mesh, const double* y, const double* p,
                           double* Lb, double* Rb) {
        const long long tot = (long long)(N - 1) * 2 * P::n;
        k_jac_blocks<P, ORDER><<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(N, mesh, y, p, Lb, Rb);
    }
    // large state dimension: stage Jacobians + dense DMMA chain-rule products (stagejac.cuh); MIRK_RESJAC=tape keeps
    // the per-column taped sweep for A/B runs
    static constexpr bool kDense = (P::n == 64 || P::n == 128) && (ORDER == 4 || ORDER == 6) && !HasSingular<P>::value;
    static constexpr size_t kDenseBatchDoubles = (size_t)1 << 28;  // 2 GiB of scratch at most: longer meshes run in batches
    static bool dense_enabled() {
        static const bool off = getenv("MIRK_RESJAC") && (!strcmp(getenv("MIRK_RESJAC"), "tape") || !strcmp(getenv("MIRK_RESJAC"), "dual"));
        return kDense && !off;
    }
    static size_t resjac_scratch_doubles(int N) {
        if constexpr (kDense) {
            if (!dense_enabled()) return 0;
            const size_t per = stagejac_doubles_per_interval<P, ORDER>(), want = per * (size_t)(N - 1);
            return want < kDenseBatchDoubles ? want : (kDenseBatchDoubles / per) * per;
        } else {
            return 0;
        }
    }
    static void resjac(cudaStream_t st, int N, const double* mesh, const double* y, const double* p, double* Kd,
                       double* phi_out, unsigned long long* nb, double* Lb, double* Rb, double* scratch) {
        if constexpr (kDense) {
            if (dense_enabled() && scratch) {
                residual(st, N, mesh, y, p, Kd, phi_out, nb);  // stages, Phi, |Phi|_inf
                const size_t per = stagejac_doubles_per_interval<P, ORDER>();
                const int batch = (int)(resjac_scratch_doubles(N) / per);
                static const cudaError_t attr = cudaFuncSetAttribute(k_chain_gemm<P::n, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                     (int)(sizeof(double) * ChainGemm<P::n, ORDER>::smem_doubles));
                (void)attr;
                for (int i0 = 0; i0 < N - 1; i0 += batch) {
                    const int cnt = (N - 1 - i0) < batch ? (N - 1 - i0) : batch;
                    k_stage_jac<P, ORDER><<<cnt, P::n, 0, st>>>(i0, N, mesh, y, p, Kd, scratch);
                    k_chain_gemm<P::n, ORDER><<<cnt, 256, sizeof(double) * ChainGemm<P::n, ORDER>::smem_doubles, st>>>(i0, N, mesh, scratch, Lb, Rb);
                }
                return;
            }
        }
        // default: taped values + tangent replay (k_resjac_tape); MIRK_RESJAC=dual selects the plain dual
        // sweep per column, MIRK_TAPE_IPW the intervals per warp (tuning / A-B measurements)
        static const bool use_dual = getenv("MIRK_RESJAC") && !strcmp(getenv("MIRK_RESJAC"), "dual");
        if (use_dual) {
            const long long tot = (long long)(N - 1) * 2 * P::n;
            k_resjac<P, ORDER><<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(N, mesh, y, p, Kd, phi_out, nb, Lb, Rb);
            return;
        }
        static const int ipw_env = getenv("MIRK_TAPE_IPW") ? atoi(getenv("MIRK_TAPE_IPW")) : 0;
        static const int warps_env = getenv("MIRK_TAPE_WARPS") ? atoi(getenv("MIRK_TAPE_WARPS")) : 0;
        static const int slots = [] {
            int occ = 0, dev = 0, sms = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_resjac_tape<P, ORDER>, 32, 0);
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            return (occ > 0 ? occ : 1) * (sms > 0 ? sms : 1);
        }();
        int ipw = tape_intervals_per_cta(N - 1, slots), warps = 1;
        if (ipw_env >= 1 && ipw_env <= 32) ipw = ipw_env;
        if (warps_env >= 1 && warps_env <= kTapeThreads / 32) warps = warps_env;
        k_resjac_tape<P, ORDER><<<(unsigned)((N - 1 + ipw - 1) / ipw), 32 * warps, 0, st>>>(N, ipw, mesh, y, p, Kd,
                                                                                          phi_out, nb, Lb, Rb);
    }
    static void defect(cudaStream_t st, int N, const double* mesh, const double* y, const double* p,
                       const double* Kd, double* Ki, double* errors, double* est,
                       unsigned long long* db) {
        k_defect<P, ORDER><<<(N - 1 + 127) / 128, 128, 0, st>>>(N, mesh, y, p, Kd, Ki, errors, est, db);
    }
    static void interp_setup(cudaStream_t st, int N, const double* mesh, const double* y, const double* p,
                             const double* Kd, double* Ki) {
        k_interp_setup<P, ORDER><<<(N - 1 + 127) / 128, 128, 0, st>>>(N, mesh, y, p, Kd, Ki);
    }
    static int bc_nodes_host(int N, const double* mesh, const double* p, int* nodes) {
        if (P::problem_type == 1) { nodes[0] = 0; nodes[1] = N - 1; return 2; }
        double tm[P::max_bc_pts];
        const int m = P::bc_times(tm, p, mesh[0], mesh[N - 1]);
        for (int k = 0; k < m; k++)
            nodes[k] = tm[k] == mesh[0] ? 0 : tm[k] == mesh[N - 1] ? N - 1 : interval_of(mesh, N, tm[k]);
        return m;
    }
    static ProblemOps make(const char* name) {
        using TB = Tableau<ORDER>;
        return ProblemOps{name, ORDER, P::n, P::np, P::n_bc, P::n_bca, P::problem_type, P::max_bc_pts,
                          TB::s, TB::s_star, &residual, &bc, &jac_blocks, &resjac, &resjac_scratch_doubles, &defect, &interp_setup,
                          &bc_nodes_host};
    }
};

// defined one per translation unit group (ops_*.cu) so the registry compiles in parallel
const ProblemOps* ops_small(int id, int order);      // MIRK4, MIRK6
const ProblemOps* ops_small_235(int id, int order);  // MIRK2, MIRK3, MIRK5
const ProblemOps* ops_small_6i(int id, int order);   // MIRK6I (order code kMIRK6I = 7)
const ProblemOps* ops_chain8(int order);
const ProblemOps* ops_chain16(int order);
const ProblemOps* ops_bratu64(int order);

// K6 (ensemble.cuh): whole adaptive solves, one thread per trajectory, for the small built-ins
struct EnsArgs;
struct EnsWarpArgs;
struct EnsembleOps {
    int n, np, slots_per_node, oMESH, oY;
    void (*run)(cudaStream_t, const EnsArgs& a);                     // thread per trajectory, HBM slab (ensemble.cuh)
    cudaError_t (*run_warp)(cudaStream_t, const EnsWarpArgs& w);      // warp per trajectory, on-chip state (n <= 2), or null
    size_t (*warp_smem_bytes)(int NC);                                // dynamic shared memory of ONE warp at capacity NC
};
const EnsembleOps* ensemble_ops_small(int id, int order);

}  // namespace mirk
