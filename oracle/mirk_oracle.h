/*
 * mirk_oracle.h — CPU (FP64, single-thread) restatement of the MIRK collocation hot path of
 * SciML/BoundaryValueDiffEq.jl v5.23.3.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (boundaryvaluediffeq.jl_b200/csrc) never links or calls anything in oracle/.
 *
 * PARITY UNPINNED against a live Julia run: julia is not installed in this image or on the GPU
 * box, and the reference holds no golden vectors.  The oracle is pinned
 * instead against (i) exact-rational tableau identities, (ii) scipy.integrate.solve_bvp's
 * collocation residual and global Jacobian (same 4th-order Lobatto formula as MIRK4),
 * (iii) the analytic solutions / convergence orders / known-answer constants the reference's own
 * tests use (lib/BoundaryValueDiffEqMIRK/test/Core/mirk_basic_tests.jl), see tests/test_oracle_*.py.
 *
 * Reference paths below are relative to /root/reference:
 *   MIRK/ = lib/BoundaryValueDiffEqMIRK/src/, CORE/ = lib/BoundaryValueDiffEqCore/src/.
 */
#ifndef MIRK_ORACLE_H
#define MIRK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_S 5      /* discrete stages (MIRK6)                       */
#define ORC_MAX_SI 4     /* interpolation stages s* - s (MIRK6)           */
#define ORC_MAX_SS 9     /* s*                                            */
#define ORC_MAX_BC_PTS 8 /* evaluation points a boundary condition may use */

/* return codes (mirror SciMLBase.ReturnCode as far as the path uses them) */
enum { ORC_SUCCESS = 0, ORC_FAILURE = 1, ORC_MAXITERS = 2, ORC_UNSTABLE = 3, ORC_STALLED = 4 };

/* f!(du,u,p,t) — CORE BVProblem in-place RHS */
typedef void (*orc_rhs_fn)(double *du, const double *u, const double *p, double t, void *ctx);
/* analytic df/du, row-major n×n: J[i*n+j] = d f_i / d u_j (the reference gets it from ForwardDiff) */
typedef void (*orc_rhs_jac_fn)(double *J, const double *u, const double *p, double t, void *ctx);
/* evaluation times the boundary condition reads `sol(t)` at; returns the count m */
typedef int (*orc_bc_times_fn)(double *times, const double *p, double t0, double t1, void *ctx);
/* bc!(res, sol, p, t) restated pointwise: U is m×n row-major, U[k] = sol(times[k]) */
typedef void (*orc_bc_fn)(double *res, const double *U, const double *p, void *ctx);
/* d res / d U, row-major L × (m*n) */
typedef void (*orc_bc_jac_fn)(double *dres, const double *U, const double *p, void *ctx);

typedef struct {
    int n;            /* states (M in the reference)                                   */
    int n_p;          /* parameter count                                               */
    int problem_type; /* 0 = StandardBVProblem, 1 = TwoPointBVProblem                  */
    int n_bc;         /* L boundary rows (== n on this path; NLLS is out of scope)     */
    int n_bca;        /* two-point: rows of bca (first block of resid), rest are bcb   */
    orc_rhs_fn f;
    orc_rhs_jac_fn dfdu;
    orc_bc_times_fn bc_times;
    orc_bc_fn bc;
    orc_bc_jac_fn dbc;
    void *ctx;
    const double *singular_term; /* prob.singular_term: n×n row-major S of y' = S y / t + f(t, y), or NULL
                                    (CORE/src/utils.jl:932-941; added to the discrete stages with t > 0 only) */
    int bc_uses_derivative;      /* bc! also reads sol(t, Val{1}) (MIRK/src/interpolation.jl:277-292): U[k] (m×n) is then
                                    followed by dU[k] = sol'(times[k]) (m×n more doubles).  The derivative is built from the
                                    Float64 stage buffers, so it carries no dual part: dbc stays L × (m*n), d/dU only */
} orc_problem;

typedef struct {
    int order, s, s_star;
    double c[ORC_MAX_S], v[ORC_MAX_S], b[ORC_MAX_S], x[ORC_MAX_S][ORC_MAX_S];
    double c_star[ORC_MAX_SI], v_star[ORC_MAX_SI], x_star[ORC_MAX_SI][ORC_MAX_SS];
    double tau_star;
} orc_tableau;

typedef struct {
    double abstol;             /* 1e-6   MIRK/mirk.jl:50                       */
    int adaptive;              /* 1                                             */
    double defect_threshold;   /* 0.1    CORE/calc_errors.jl DefectControl      */
    int max_num_subintervals;  /* 3000   MIRK/algorithms.jl:55-61               */
    int maxiters;              /* 1000   NonlinearSolve default                 */
    int reinterp_inplace;      /* 1 = reproduce quirk Q3 (MIRK/mirk.jl:368-370); default 0 */
    int max_outer;             /* safety cap on adaptive outer iterations       */
    int nlsolve;               /* 0 default polyalgorithm (CORE/default_internal_solve.jl:31-45); 1 NewtonRaphson,
                                  2 NewtonRaphson + BackTracking, 3 TrustRegion */
    int controller;            /* 0 DefectControl, 1 GlobalErrorControl, 2 SequentialErrorControl, 3 HybridErrorControl
                                  (CORE/src/calc_errors.jl:54-106) */
    int ge_method;             /* 0 HOErrorControl (order + 2), 1 REErrorControl (Richardson on the halved mesh) */
    double DE, GE;             /* HybridErrorControl weights */
} orc_options;

typedef struct {
    int N;                /* final mesh nodes                                     */
    double *mesh;         /* N                                                    */
    double *y;            /* N×n node-major (reference flat layout)               */
    double *Kd;           /* (N-1)×s×n discrete stages of the last evaluation     */
    double *Ki;           /* (N-1)×(s*-s)×n interpolation stages                  */
    int retcode;
    double resid_norm;    /* ||F||inf of the last Newton solve                    */
    double defect_norm;   /* last defect estimate (2*abstol when non adaptive)    */
    int outer_iters;      /* calls of __perform_mirk_iteration                    */
    int newton_iters;     /* total Newton steps over all outer iterations         */
    int n_hist;           /* entries used in the histories below (<= 64)          */
    int hist_N[64];       /* mesh nodes per outer iteration                       */
    int hist_newton[64];  /* Newton steps per outer iteration                     */
    double hist_defect[64];
} orc_result;

void orc_default_options(orc_options *o);
#define ORC_MIRK6I 7 /* `order` code of MIRK6I (irrational 6th-order tableau); 2..6 are MIRK2..MIRK6 */
int orc_convergence_order(int order);
int orc_tableau_get(int order, orc_tableau *T);
/* threads over mesh intervals in orc_phi / orc_jac_blocks (bench.py's CPU baseline only; default 1, results unchanged) */
void orc_set_interval_threads(int nthreads);
void orc_interp_weights(int order, double tau, double *w, double *wp);
void orc_mesh_uniform(double t0, double t1, int nint, double *mesh);
int orc_interval(const double *mesh, int N, double t); /* 0-based interval index */

void orc_phi(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
             const double *y, double *Kd, double *phi);
void orc_interp_setup(const orc_problem *P, const orc_tableau *T, const double *p, int N,
                      const double *mesh, const double *y, const double *Kd, double *Ki);
void orc_eval_sol(const orc_problem *P, const orc_tableau *T, int N, const double *mesh,
                  const double *y, const double *Kd, const double *Ki, double t, int deriv,
                  int bc_shortcut, double *out);
void orc_loss(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
              const double *y, double *Kd, double *Ki, double *resid);
void orc_jac_blocks(const orc_problem *P, const orc_tableau *T, const double *p, int N,
                    const double *mesh, const double *y, double *Lb, double *Rb);
int orc_bc_jac(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
               const double *y, const double *Kd, const double *Ki, int *nodes, double *B);
int orc_abd_solve(int n, int N, int L, const double *Lb, const double *Rb, int m, const int *nodes,
                  const double *B, const double *rhs_bc, const double *rhs_phi, double *delta);
int orc_newton(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
               double *y, double *Kd, double *Ki, double abstol, int maxiters, double *resid_norm,
               int *iters);
int orc_nlsolve(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
                double *y, double *Kd, double *Ki, double abstol, int maxiters, int nlsolve, double *resid_norm,
                int *iters);
double orc_defect(const orc_problem *P, const orc_tableau *T, const double *p, int N,
                  const double *mesh, const double *y, const double *Kd, double *Ki, double *errors);
int orc_mesh_select(int order, int n, int N, const double *mesh, const double *errors, double abstol,
                    int max_num_subintervals, int *N_new, double *mesh_new /* cap 4N */);
int orc_mesh_select_ex(int order, int n, int N, const double *mesh, const double *errors, const double *errors2,
                       double abstol, int max_num_subintervals, int expo_den, double rho, int *N_new, double *mesh_new);
double orc_global_error(const orc_problem *P, int order, const double *p, int N, const double *mesh, const double *y,
                        const orc_options *opt, int method, double *errors);
void orc_reinterp(const orc_problem *P, const orc_tableau *T, int N_old, const double *mesh_old,
                  const double *y_old, const double *Kd, const double *Ki, int N_new,
                  const double *mesh_new, double *y_new, int inplace_quirk);
int orc_solve(const orc_problem *P, int order, const double *p, int N0, const double *mesh0,
              const double *y0, const orc_options *opt, orc_result *out);
void orc_result_free(orc_result *r);

/* built-in problems, ids shared with the CUDA registry (boundaryvaluediffeq.jl_b200/csrc/problems.cuh) */
int orc_builtin_problem(int id, orc_problem *P);
const char *orc_builtin_name(int id);

/* ensemble: ntraj independent solves from a constant guess on a uniform mesh; p is ntraj×n_p.
 * nthreads>1 uses OpenMP over trajectories (mirrors SciMLBase EnsembleThreads). */
int orc_ensemble_solve(const orc_problem *P, int order, int ntraj, const double *p, const double *u0,
                       double t0, double t1, int nint, const orc_options *opt, int nthreads,
                       int *retcodes, int *N_final, double *y_first /* ntraj×n: y at node 0 */,
                       int *newton_iters);

#ifdef __cplusplus
}
#endif
#endif
