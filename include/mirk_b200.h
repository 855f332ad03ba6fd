/*
 * mirk_b200.h — C ABI of libmirkb200.so: the B200-native MIRK4/MIRK6 collocation Newton path.
 *
 * Drop-in boundary for the hot path of SciML/BoundaryValueDiffEq.jl (reference paths relative to
 * /root/reference; MIRK/ = lib/BoundaryValueDiffEqMIRK/src, CORE/ = lib/BoundaryValueDiffEqCore/src).
 * A Julia `BoundaryValueDiffEqMIRK` backend binds these with `ccall` (INTEGRATION.md shows the
 * stubs); the tests and bench bind them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every HOST buffer is owned by the caller, every device
 *     buffer by the library; one CUDA stream per handle; handles are independent and re-entrant
 *     (the reference allows concurrent solves on Julia threads, MIRK/BoundaryValueDiffEqMIRK.jl:106-109)
 *   - every entry point returns an int status: 0 ok, < 0 argument / environment error, > 0 a
 *     numerical outcome mirroring SciMLBase.ReturnCode; nothing throws or aborts.
 *     mirk_last_error() gives the message of the calling thread's last failure
 *   - there is NO CPU fallback: without a CUDA device mirk_create fails with MIRK_ERR_NO_DEVICE
 *   - unknowns are node-major `y[N][n]` exactly like the reference's flat vector
 *     (CORE/utils.jl:59-66); residual order is the reference's: Standard `[bc; Phi_1..Phi_{N-1}]`
 *     (MIRK/mirk.jl:471-482), TwoPoint `[bc_a; Phi...; bc_b]` (CORE/utils.jl:42-52)
 */
#ifndef MIRK_B200_H
#define MIRK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* argument / environment errors */
#define MIRK_OK 0
#define MIRK_ERR_ARG (-1)         /* bad argument, e.g. dt <= 0 ("dt must be positive", CORE/utils.jl:354) */
#define MIRK_ERR_NO_DEVICE (-2)   /* no CUDA device: the product path has no CPU fallback */
#define MIRK_ERR_CUDA (-3)        /* a CUDA runtime call failed */
#define MIRK_ERR_UNSUPPORTED (-4) /* unknown problem id / order */
#define MIRK_ERR_STATE (-5)       /* call sequence error (e.g. no mesh set) */
/* numerical outcomes (SciMLBase.ReturnCode as far as this path produces them) */
#define MIRK_RET_SUCCESS 0
#define MIRK_RET_FAILURE 1
#define MIRK_RET_MAXITERS 2
#define MIRK_RET_UNSTABLE 3
#define MIRK_RET_STALLED 4

typedef struct mirk_solver_s* mirk_handle;

/* Replaces the BVProblem + algorithm + solve kwargs a `__init(prob, alg; ...)` call receives
 * (MIRK/mirk.jl:49-53, MIRK/algorithms.jl:55-61). */
typedef struct {
    int32_t problem_id;            /* device functor: id from mirk_problem_lookup()             */
    int32_t order;                 /* 2..6 = MIRK2() .. MIRK6(), 7 = MIRK6I() (2, 3, 5, 7: built-ins with n <= 6) */
    double abstol;                 /* 1e-6                                                      */
    int32_t adaptive;              /* 1                                                         */
    double defect_threshold;       /* DefectControl().defect_threshold = 0.1                    */
    int32_t max_num_subintervals;  /* 3000                                                      */
    int32_t maxiters;              /* nlsolve_kwargs maxiters, 1000                             */
    int32_t reinterp_inplace;      /* 0; 1 reproduces the reference's in-place hazard (Q3)      */
    int32_t chunk;                 /* relations collapsed per group and level of the ABD solve  */
    int32_t device;                /* CUDA device ordinal                                       */
    int32_t n_params;              /* length of params                                          */
    const double* params;          /* prob.p (host)                                             */
    int32_t nlsolve;               /* alg.nlsolve: 0 = nothing given -> the reference's default polyalgorithm
                                      NewtonRaphson -> NewtonRaphson + BackTracking -> TrustRegion, each restarting
                                      from the original iterate (CORE/default_internal_solve.jl:31-45); 1, 2, 3 = that
                                      single solver */
    int32_t controller;            /* 0 DefectControl, 1 GlobalErrorControl, 2 SequentialErrorControl,
                                      3 HybridErrorControl (CORE/calc_errors.jl:54-106, MIRK/adaptivity.jl:77-243,464-567) */
    int32_t ge_method;             /* global-error estimate: 0 HOErrorControl (method of order + 2 on the same mesh),
                                      1 REErrorControl (Richardson: same method on the halved mesh) */
    double DE, GE;                 /* HybridErrorControl weights: error = DE * defect + GE * global error */
} mirk_desc;

typedef struct {
    int32_t n, n_params, problem_type /* 0 Standard, 1 TwoPoint */, n_bc, n_bca, max_bc_pts;
} mirk_problem_info;

/* What `solve!` returns (MIRK/mirk.jl:286-332): sizes and scalars; arrays via mirk_get_*. */
typedef struct {
    int32_t retcode;       /* MIRK_RET_*                                                        */
    int32_t n_mesh;        /* nodes of the final mesh (length of sol.t)                         */
    int32_t outer_iters;   /* calls of __perform_mirk_iteration                                  */
    int32_t newton_iters;  /* Newton steps over all outer iterations                             */
    double resid_norm;     /* |sol.resid|_inf of the last Newton solve                           */
    double defect_norm;    /* last defect estimate (2 abstol when adaptive = false)              */
    int32_t n_hist;        /* entries used below                                                 */
    int32_t hist_n_mesh[64];
    int32_t hist_newton[64];
    double hist_defect[64];
} mirk_result;

/* -- library ----------------------------------------------------------------------------------- */
int mirk_version(void);
const char* mirk_last_error(void);
int mirk_device_count(int32_t* count);

/* -- device-function registry (replaces prob.f / prob.f.bc closures, MIRK/mirk.jl:71-116) ------- */
int mirk_problem_lookup(const char* name, int32_t* problem_id);
int mirk_problem_info_get(int32_t problem_id, mirk_problem_info* info);
/* register a user functor compiled to a shared object from csrc/plugin template (the CUDA-side twin of
 * passing Julia closures f!/bc! to BVProblem); see INTEGRATION.md */
int mirk_problem_register_plugin(const char* name, const char* so_path, int32_t* problem_id);

/* collect(range(t0; stop = t1, length = nint + 1)) — CORE/utils.jl:694 (host helper) */
int mirk_mesh_uniform(double t0, double t1, int32_t nint, double* mesh);

/* -- cache life cycle: SciMLBase.__init / finalizer (MIRK/mirk.jl:49-265) ------------------------ */
int mirk_create(const mirk_desc* desc, mirk_handle* out);
int mirk_destroy(mirk_handle h);
int mirk_set_params(mirk_handle h, const double* params, int32_t n_params);
/* mesh + initial guess: __extract_mesh / __initial_guess_on_mesh (CORE/utils.jl:694,750-773) */
int mirk_set_mesh_guess(mirk_handle h, int32_t n_mesh, const double* mesh, const double* y);
/* the same with the guess already on the handle's device (a CuArray on the Julia side; ext/ CUDA extension): no
 * host round trip for y; the mesh stays a host array */
int mirk_set_mesh_guess_device(mirk_handle h, int32_t n_mesh, const double* mesh, const double* d_y);
/* uniform mesh of cld(t1 - t0, dt) intervals with u0 copied to every node (CORE/utils.jl:349-363,766-769) */
int mirk_set_uniform_guess(mirk_handle h, double t0, double t1, double dt, const double* u0);

/* -- pieces of one Newton step (parity tests and the Newton-steps/s metric) ---------------------- */
/* loss(du,u,p): __mirk_loss! (MIRK/mirk.jl:471-534); resid may be NULL */
int mirk_residual(mirk_handle h, double* resid, double* resid_norm);
/* jac(J,u,p): __mirk_mpoint_jacobian! / __mirk_2point_jacobian! (MIRK/mirk.jl:810-862,994-1002) in
 * block form: Lb,Rb are (N-1) row-major n×n blocks; bc_nodes[m] 0-based pinned nodes and Bc[m][n_bc][n]
 * the boundary rows (reference pattern).  Any pointer may be NULL. */
int mirk_jacobian_blocks(mirk_handle h, double* Lb, double* Rb, int32_t* bc_nodes, double* Bc, int32_t* m);
/* J \ F for the current iterate (stands in for LinearSolve inside NonlinearSolve); delta is N×n, may be NULL.
 * Requires mirk_residual + mirk_jacobian_blocks state; leaves y untouched. */
int mirk_linear_solve(mirk_handle h, double* delta);
/* The almost-block-diagonal solver on its own (what FIRK's expanded form — blocks of n (s + 1),
 * lib/BoundaryValueDiffEqFIRK/src/sparse_jacobians.jl:35-70 — and MIRKN — lib/BoundaryValueDiffEqMIRKN/src/
 * collocation.jl:8-41 — share with MIRK): solves J delta = rhs for J = [boundary rows; blockbidiag(Lb_i, Rb_i)] with any
 * block size n.  Lb, Rb: (N-1) row-major n x n blocks; Bc[m][n][n]: boundary rows' blocks on nodes bc_nodes[m] (0-based);
 * rhs in residual order (two_point = 0: [bc; Phi], 1: [bc_a(La); Phi; bc_b]); delta[N][n].  Square systems (n boundary
 * rows) only; host arrays; returns MIRK_RET_SUCCESS / MIRK_RET_FAILURE (singular block). */
int mirk_abd_solve(int32_t n, int32_t N, int32_t two_point, int32_t La, const double* Lb, const double* Rb, int32_t m,
                   const int32_t* bc_nodes, const double* Bc, const double* rhs, double* delta, int32_t device);
/* one NewtonRaphson iteration: J d = F, y -= d, F(y) ; returns the new |F|_inf */
int mirk_newton_step(mirk_handle h, double* resid_norm);
/* __internal_solve(nlprob, alg; abstol, maxiters) with alg per desc.nlsolve (CORE/default_internal_solve.jl:31-45,
 * 107-110); returns Success / Failure / MaxIters / Unstable / Stalled, iters = steps of every sub-solver that ran */
int mirk_newton_solve(mirk_handle h, int32_t* iters, double* resid_norm);
/* steps and return code of each sub-solver (NewtonRaphson, + BackTracking, TrustRegion) in the last nonlinear solve;
 * -1 where a sub-solver did not run.  (A sub-solver that DIVERGES amplifies rounding differences, so its step count
 * is not reproducible across linear solvers; the converging one's is.) */
int mirk_nlsolve_stats(mirk_handle h, int32_t* steps3, int32_t* retcodes3);
/* error_estimate!(cache, DefectControl, ...) (MIRK/adaptivity.jl:370-415); errors is (N-1)×n, may be NULL */
int mirk_defect(mirk_handle h, double* errors, double* defect_norm);
/* mesh_selector! + interp_eval! + __expand_cache! (MIRK/adaptivity.jl:23-75, mirk.jl:364-372);
 * returns MIRK_RET_SUCCESS / MIRK_RET_FAILURE and the new node count */
int mirk_refine_mesh(mirk_handle h, int32_t* n_mesh_new);

/* mesh_selector! + redistribute! / half_mesh! of DefectControl (MIRK/adaptivity.jl:23-75,250-304) as a function of a mesh and
 * its per-interval error estimates est[n_mesh - 1] = |errors[i]|_inf (host arrays): what the mesh-partitioned mode runs on
 * the estimates gathered from every rank (partition.solve_partitioned).  mesh_new must hold
 * max(n_mesh, max_num_subintervals + 1) entries.  Returns MIRK_RET_SUCCESS and *n_mesh_new, or MIRK_RET_FAILURE when the
 * new mesh would exceed max_num_subintervals (mesh_new untouched). */
int mirk_mesh_select(int32_t order, double abstol, int32_t max_num_subintervals, int32_t n_mesh, const double* mesh,
                     const double* est, int32_t* n_mesh_new, double* mesh_new, int32_t device);

/* -- SciMLBase.solve!(cache) (MIRK/mirk.jl:286-388) ---------------------------------------------- */
int mirk_solve(mirk_handle h, mirk_result* result);

/* -- solution access: sol.t, sol.u, sol(t), sol(t, Val{1}) (MIRK/interpolation.jl:17-204) -------- */
int mirk_get_mesh_size(mirk_handle h, int32_t* n_mesh);
int mirk_get_solution(mirk_handle h, double* mesh, double* y);
int mirk_get_solution_device(mirk_handle h, double* d_y); /* sol.u into a device buffer [n_mesh][n] */
int mirk_get_stages(mirk_handle h, double* Kd, double* Ki);
int mirk_get_residual(mirk_handle h, double* resid);
int mirk_interp(mirk_handle h, const double* t, int32_t m, int32_t deriv, double* out);

/* -- measurement helpers (bench.py): device-timed Newton steps with inputs resident in HBM ------- */
/* runs `steps` full Newton steps (residual + Jacobian + ABD solve + update) from the stored guess,
 * resetting y to the guess before each step so every step does identical work, through the same code
 * path as mirk_newton_solve (CUDA-graph replay once warm); CUDA-event time of that region in ms and the
 * kernels it launched in *launches.  The same steps are then repeated with an event after every phase
 * (direct launches) for the per-phase sums (unused, residual + jacobian, reduce level 0, upper reduce
 * levels, tail + closing solve, back substitution, update, unused) in phase_ms[8]. */
int mirk_bench_newton_steps(mirk_handle h, int32_t steps, float* total_ms, float* phase_ms, int64_t* launches);
/* roofline denominators measured on the device: FP64 FMA TFLOP/s and HBM copy GB/s */
int mirk_measure_peaks(int32_t device, double* fp64_tflops, double* hbm_gbs);

/* -- mesh-partitioned single problem (no reference counterpart: the reference is single-threaded).
 *    Every rank creates a handle on its GPU holding one contiguous mesh segment (neighbours share their
 *    boundary node) of a TwoPointBVProblem — or of a Standard BVProblem whose boundary condition reads the solution at the
 *    two END POINTS only: the outer end states are then exchanged before every boundary evaluation and each rank evaluates
 *    the rows on them (interior evaluation times: MIRK_ERR_UNSUPPORTED) —, then attaches it to a communicator.  The handle's mesh is fixed between
 *    mirk_set_mesh_guess calls; the adaptive outer loop runs above the ABI (partition.solve_partitioned: local defect
 *    estimates, mirk_mesh_select on the gathered estimates, local re-interpolation, re-partitioning).  From then
 *    on mirk_residual / mirk_newton_step / mirk_newton_solve / mirk_solve / mirk_bench_newton_steps are
 *    COLLECTIVE calls: per Newton step one 8-byte all-reduce(max) of |F|_inf and one all-gather of the
 *    (2n^2 + n + 2Ln + L)-double reduced interface relation per rank, over NCCL on the solver's stream.
 *    NCCL is dlopen'ed on first use (libnccl_path or the default soname), never at link time. */
int mirk_nccl_unique_id(void* id128 /* 128 bytes */, const char* libnccl_path);
int mirk_partition_attach(mirk_handle h, int32_t rank, int32_t nranks, const void* id128, const char* libnccl_path);
/*    The same exchange over NVLink peer memory, without a library collective (the default of the host layer):
 *    every rank allocates an exchange buffer and exports its CUDA IPC handle (64 bytes); the caller gathers the
 *    handles of all ranks (rank order) by any means and attaches.  Per Newton step each rank then PUSHES its packed
 *    relation into every peer's buffer from inside the pack kernel and spins on local flags (k_part_push /
 *    k_part_wait_unpack / k_words_allmax in csrc/abd.cuh); these are plain kernels with device-resident epochs, so the
 *    whole partitioned step replays as a CUDA graph.  All ranks must make the same sequence of collective calls;
 *    a rank that waits longer than ~2 s for a peer gives up and reports MIRK_RET_FAILURE.  Ranks must be processes
 *    on one node whose GPUs have peer access (NVLink / NVSwitch). */
int mirk_partition_p2p_export(mirk_handle h, int32_t rank, int32_t nranks, void* ipc_handle64 /* 64 bytes out */);
int mirk_partition_attach_p2p(mirk_handle h, int32_t rank, int32_t nranks, const void* ipc_handles /* nranks x 64 bytes */);

/* -- ensembles: solve(EnsembleProblem(prob; prob_func), alg; trajectories, dt)
 *    (SciMLBase driver; usage lib/BoundaryValueDiffEqMIRK/test/Core/ensemble_tests.jl:20-38).
 *    The host harvests prob_func's parameters (and optionally per-trajectory u0) into packed arrays;
 *    every trajectory then runs the complete adaptive solve! loop on the device, one thread each. ---- */
typedef struct mirk_ensemble_s* mirk_ensemble_handle;
typedef struct {
    int32_t problem_id, order;
    double abstol;
    int32_t adaptive;
    double defect_threshold;
    int32_t max_num_subintervals, maxiters, reinterp_inplace, device;
    int32_t node_cap;       /* per-trajectory mesh capacity in nodes; 0: max_num_subintervals + 1 (state lives on chip
                               up to the kernel's shared-memory capacity, beyond it on an HBM slab).  A mesh that would
                               outgrow node_cap ends that trajectory with MIRK_RET_FAILURE, like max_num_subintervals */
    double t0, t1, dt;      /* tspan and dt: uniform initial mesh of cld(t1 - t0, dt) intervals */
    int32_t nlsolve;        /* 0: the default polyalgorithm (the batched kernels run NewtonRaphson; a trajectory whose
                               Newton solve fails is re-run by the single-problem driver with the BackTracking and
                               TrustRegion fallbacks); 1: NewtonRaphson only */
} mirk_ensemble_desc;
int mirk_ensemble_create(const mirk_ensemble_desc* desc, int64_t ntraj, mirk_ensemble_handle* out);
int mirk_ensemble_destroy(mirk_ensemble_handle h);
/* params[ntraj][n_params]; u0[n] shared by all trajectories or u0[ntraj][n] when u0_per_traj != 0 */
int mirk_ensemble_set_inputs(mirk_ensemble_handle h, const double* params, const double* u0, int32_t u0_per_traj);
/* the same from device-resident arrays (CuArray params / u0 on the handle's device) */
int mirk_ensemble_set_inputs_device(mirk_ensemble_handle h, const double* d_params, const double* d_u0, int32_t u0_per_traj);
int mirk_ensemble_run(mirk_ensemble_handle h, float* device_ms);
/* per-trajectory outcomes (any pointer may be NULL): sol.retcode, length(sol.t), Newton steps, outer
 * iterations, |sol.resid|_inf, last defect, sol.u[1] */
int mirk_ensemble_get_results(mirk_ensemble_handle h, int32_t* retcodes, int32_t* n_mesh, int32_t* newton_iters,
                              int32_t* outer_iters, double* resid_norm, double* defect_norm, double* y_first);
/* the per-trajectory mesh capacity in nodes the handle was created with (sizes the buffers below) */
int mirk_ensemble_node_cap(mirk_ensemble_handle h, int32_t* node_cap);
/* sol.t / sol.u of one trajectory; mesh[node_cap], y[node_cap][n] */
int mirk_ensemble_get_trajectory(mirk_ensemble_handle h, int64_t traj, int32_t* n_mesh, double* mesh, double* y);
/* one-shot convenience: create + set_inputs + run + get_results + destroy */
int mirk_ensemble_solve(const mirk_ensemble_desc* desc, int64_t ntraj, const double* params, const double* u0,
                        int32_t u0_per_traj, int32_t* retcodes, int32_t* n_mesh, int32_t* newton_iters, double* y_first);

#ifdef __cplusplus
}
#endif
#endif
