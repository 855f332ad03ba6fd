/*
 * mirk_problems.c — built-in BVPs for the oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Same ids / parameter layouts as the CUDA registry in
 * boundaryvaluediffeq.jl_b200/csrc/problems.cuh, but written independently: the oracle carries
 * hand-derived analytic Jacobians df/du and dbc/dU, the CUDA side differentiates the templated
 * RHS with dual numbers (what the reference does through ForwardDiff), so agreement of the two
 * is a real check of both.
 *
 * Problem sources (paths relative to /root/reference):
 *   0 pendulum      benchmark/simple_pendulum.jl:5-19,32                     (BASELINE C1/C3)
 *   1 linear2       lib/BoundaryValueDiffEqMIRK/test/Core/mirk_basic_tests.jl:16-33,205-245,
 *                   ensemble_tests.jl:10-18                                  (u''=-k u, 2 BC points)
 *   2 linear2_tp    mirk_basic_tests.jl:42-47                                (TwoPointBVProblem)
 *   3 swirling      mirk_basic_tests.jl:315-344                              (n=6, eps)
 *   4 lotka         mirk_basic_tests.jl:438-455                              (both BC at t0)
 *   5 torus         mirk_basic_tests.jl:726-756                              (n=4 geodesic)
 *   6 layer         test/misc/adaptivity_tests.jl:7-17                       (boundary layer, p=1e-3)
 *   7 chain8        SURVEY.md §8d C2: 8 torsionally coupled pendula, n=16, two-point
 *   8 chain16       SURVEY.md §8d C5: 16 pendula, n=32
 *   9 bratu64       SURVEY.md §8d C4: 2-D Bratu by method of lines, n=128
 *  10 lane_emden    lib/BoundaryValueDiffEqMIRK/test/Core/singular_bvp_tests.jl:15-61 (singular_term S = [0 0; 0 -2])
 *  11 robin_sine    u'' = -u with sol(t, Val{1}) inside bc! (MIRK/src/interpolation.jl:277-292; mirk_basic_tests.jl:453-479 uses
 *                   the derivative through maxsol / minsol)
 */
#include "mirk_oracle.h"

#include <math.h>
#include <string.h>

static const double PI = 3.14159265358979323846;

/* ---- 0: simple pendulum, p = [g/L] ---------------------------------------------------------- */
static void pend_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    du[0] = u[1];
    du[1] = -p[0] * sin(u[0]);
}
static void pend_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    J[0] = 0.0; J[1] = 1.0;
    J[2] = -p[0] * cos(u[0]); J[3] = 0.0;
}
static int pend_times(double *tm, const double *p, double t0, double t1, void *c) {
    (void)p; (void)c;
    tm[0] = (t0 + t1) / 2;
    tm[1] = t1;
    return 2;
}
static void pend_bc(double *r, const double *U, const double *p, void *c) {
    (void)p; (void)c;
    r[0] = U[0] + PI / 2;
    r[1] = U[2] - PI / 2;
}
static void pend_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    memset(d, 0, sizeof(double) * 2 * 4);
    d[0 * 4 + 0] = 1.0;
    d[1 * 4 + 2] = 1.0;
}

/* ---- 1: u'' = -k u with two scalar conditions, p = [k, ta, va, tb, vb, ca, cb] -------------- */
static void lin_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    du[0] = u[1];
    du[1] = -p[0] * u[0];
}
static void lin_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)u; (void)t; (void)c;
    J[0] = 0.0; J[1] = 1.0; J[2] = -p[0]; J[3] = 0.0;
}
static int lin_times(double *tm, const double *p, double t0, double t1, void *c) {
    (void)t0; (void)t1; (void)c;
    tm[0] = p[1];
    tm[1] = p[3];
    return 2;
}
static void lin_bc(double *r, const double *U, const double *p, void *c) {
    (void)c;
    r[0] = U[0 + (int)p[5]] - p[2];
    r[1] = U[2 + (int)p[6]] - p[4];
}
static void lin_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)c;
    memset(d, 0, sizeof(double) * 2 * 4);
    d[0 * 4 + 0 + (int)p[5]] = 1.0;
    d[1 * 4 + 2 + (int)p[6]] = 1.0;
}

/* ---- 2: same ODE as TwoPointBVProblem, p = [k, va, vb] -------------------------------------- */
static void lintp_bc(double *r, const double *U, const double *p, void *c) {
    (void)c;
    r[0] = U[0] - p[1];
    r[1] = U[2] - p[2];
}
static void lintp_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    memset(d, 0, sizeof(double) * 2 * 4);
    d[0] = 1.0;
    d[1 * 4 + 2] = 1.0;
}

/* ---- 3: swirling flow III, p = [eps] --------------------------------------------------------- */
static void swirl_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    const double e = p[0];
    du[0] = u[1];
    du[1] = (u[0] * u[3] - u[2] * u[1]) / e;
    du[2] = u[3];
    du[3] = u[4];
    du[4] = u[5];
    du[5] = (-u[2] * u[5] - u[0] * u[1]) / e;
}
static void swirl_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    const double e = p[0];
    memset(J, 0, sizeof(double) * 36);
    J[0 * 6 + 1] = 1.0;
    J[1 * 6 + 0] = u[3] / e; J[1 * 6 + 1] = -u[2] / e; J[1 * 6 + 2] = -u[1] / e; J[1 * 6 + 3] = u[0] / e;
    J[2 * 6 + 3] = 1.0;
    J[3 * 6 + 4] = 1.0;
    J[4 * 6 + 5] = 1.0;
    J[5 * 6 + 0] = -u[1] / e; J[5 * 6 + 1] = -u[0] / e; J[5 * 6 + 2] = -u[5] / e; J[5 * 6 + 5] = -u[2] / e;
}
static int ends_times(double *tm, const double *p, double t0, double t1, void *c) {
    (void)p; (void)c;
    tm[0] = t0;
    tm[1] = t1;
    return 2;
}
static void swirl_bc(double *r, const double *U, const double *p, void *c) {
    (void)p; (void)c;
    r[0] = U[0] + 1.0; r[1] = U[2]; r[2] = U[3];
    r[3] = U[6 + 0] - 1.0; r[4] = U[6 + 2]; r[5] = U[6 + 3];
}
static void swirl_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    memset(d, 0, sizeof(double) * 6 * 12);
    d[0 * 12 + 0] = 1.0; d[1 * 12 + 2] = 1.0; d[2 * 12 + 3] = 1.0;
    d[3 * 12 + 6] = 1.0; d[4 * 12 + 8] = 1.0; d[5 * 12 + 9] = 1.0;
}

/* ---- 4: Lotka-Volterra with both conditions at t0, p = [a,b,c,d] ---------------------------- */
static void lotka_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    du[0] = p[0] * u[0] - p[1] * u[0] * u[1];
    du[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
}
static void lotka_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    J[0] = p[0] - p[1] * u[1]; J[1] = -p[1] * u[0];
    J[2] = p[3] * u[1];        J[3] = -p[2] + p[3] * u[0];
}
static int lotka_times(double *tm, const double *p, double t0, double t1, void *c) {
    (void)p; (void)t1; (void)c;
    tm[0] = t0;
    return 1;
}
static void lotka_bc(double *r, const double *U, const double *p, void *c) {
    (void)p; (void)c;
    r[0] = U[0] - 1.0;
    r[1] = U[1] - 2.0;
}
static void lotka_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    d[0] = 1.0; d[1] = 0.0; d[2] = 0.0; d[3] = 1.0;
}

/* ---- 5: torus geodesic, p = [R, r, a1_1, a1_2, a2_1, a2_2] ---------------------------------- */
static void torus_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    const double R = p[0], r = p[1], s = sin(u[0]), co = cos(u[0]), Rt = R + r * co;
    du[0] = u[2];
    du[1] = u[3];
    du[2] = -u[3] * u[3] * Rt * s / r;
    du[3] = 2 * r * s / Rt * u[2] * u[3];
}
static void torus_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    const double R = p[0], r = p[1], s = sin(u[0]), co = cos(u[0]), Rt = R + r * co;
    memset(J, 0, sizeof(double) * 16);
    J[0 * 4 + 2] = 1.0;
    J[1 * 4 + 3] = 1.0;
    /* d/dth [-w^2 (R + r cos) sin / r] = -w^2 (-r sin^2 + (R + r cos) cos)/r */
    J[2 * 4 + 0] = -u[3] * u[3] * (-r * s * s + Rt * co) / r;
    J[2 * 4 + 3] = -2.0 * u[3] * Rt * s / r;
    /* d/dth [2 r sin/(R + r cos)] = 2 r (cos Rt + r sin^2)/Rt^2 */
    J[3 * 4 + 0] = 2 * r * (co * Rt + r * s * s) / (Rt * Rt) * u[2] * u[3];
    J[3 * 4 + 2] = 2 * r * s / Rt * u[3];
    J[3 * 4 + 3] = 2 * r * s / Rt * u[2];
}
static void torus_bc(double *r, const double *U, const double *p, void *c) {
    (void)c;
    r[0] = U[0] - p[2]; r[1] = U[1] - p[3];
    r[2] = U[4 + 0] - p[4]; r[3] = U[4 + 1] - p[5];
}
static void torus_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    memset(d, 0, sizeof(double) * 4 * 8);
    d[0 * 8 + 0] = 1.0; d[1 * 8 + 1] = 1.0; d[2 * 8 + 4] = 1.0; d[3 * 8 + 5] = 1.0;
}

/* ---- 6: boundary layer, p = [eps] ------------------------------------------------------------ */
static void layer_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)c;
    du[0] = u[1];
    du[1] = -t / p[0] * u[1] - PI * PI * cos(PI * t) - PI * t / p[0] * sin(PI * t);
}
static void layer_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)u; (void)c;
    J[0] = 0.0; J[1] = 1.0; J[2] = 0.0; J[3] = -t / p[0];
}
static void layer_bc(double *r, const double *U, const double *p, void *c) {
    (void)p; (void)c;
    r[0] = U[0] + 2.0;
    r[1] = U[2];
}
static void layer_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    memset(d, 0, sizeof(double) * 2 * 4);
    d[0] = 1.0;
    d[1 * 4 + 2] = 1.0;
}

/* ---- 7/8: chain of NP torsionally coupled pendula, u = [th_1..th_NP, om_1..om_NP],
 *           p = [g, kappa, a_1..a_NP, b_1..b_NP]; fixed chain ends th_0 = th_{NP+1} = 0 ---------- */
static void chain_f_np(int NP, double *du, const double *u, const double *p) {
    const double g = p[0], kap = p[1];
    for (int k = 0; k < NP; k++) {
        const double l = k > 0 ? u[k - 1] : 0.0, r = k < NP - 1 ? u[k + 1] : 0.0;
        du[k] = u[NP + k];
        du[NP + k] = -g * sin(u[k]) + kap * (r - 2.0 * u[k] + l);
    }
}
static void chain_df_np(int NP, double *J, const double *u, const double *p) {
    const int n = 2 * NP;
    const double g = p[0], kap = p[1];
    memset(J, 0, sizeof(double) * n * n);
    for (int k = 0; k < NP; k++) {
        J[k * n + NP + k] = 1.0;
        J[(NP + k) * n + k] = -g * cos(u[k]) - 2.0 * kap;
        if (k > 0) J[(NP + k) * n + k - 1] = kap;
        if (k < NP - 1) J[(NP + k) * n + k + 1] = kap;
    }
}
static void chain_bc_np(int NP, double *r, const double *U, const double *p) {
    const int n = 2 * NP;
    for (int k = 0; k < NP; k++) {
        r[k] = U[k] - p[2 + k];
        r[NP + k] = U[n + k] - p[2 + NP + k];
    }
}
static void chain_dbc_np(int NP, double *d) {
    const int n = 2 * NP;
    memset(d, 0, sizeof(double) * n * 2 * n);
    for (int k = 0; k < NP; k++) {
        d[k * 2 * n + k] = 1.0;
        d[(NP + k) * 2 * n + n + k] = 1.0;
    }
}
#define CHAIN_DEF(NP)                                                                              \
    static void chain##NP##_f(double *du, const double *u, const double *p, double t, void *c) {   \
        (void)t; (void)c; chain_f_np(NP, du, u, p); }                                              \
    static void chain##NP##_df(double *J, const double *u, const double *p, double t, void *c) {   \
        (void)t; (void)c; chain_df_np(NP, J, u, p); }                                              \
    static void chain##NP##_bc(double *r, const double *U, const double *p, void *c) {             \
        (void)c; chain_bc_np(NP, r, U, p); }                                                       \
    static void chain##NP##_dbc(double *d, const double *U, const double *p, void *c) {            \
        (void)U; (void)p; (void)c; chain_dbc_np(NP, d); }
CHAIN_DEF(8)
CHAIN_DEF(16)

/* ---- 9: 2-D Bratu, method of lines with M=64 interior lines, u = [u_1..u_M, v_1..v_M], p=[lambda] */
#define BRATU_M 64
static void bratu_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    const int M = BRATU_M;
    const double dz = 1.0 / (M + 1), idz2 = 1.0 / (dz * dz);
    for (int j = 0; j < M; j++) {
        const double l = j > 0 ? u[j - 1] : 0.0, r = j < M - 1 ? u[j + 1] : 0.0;
        du[j] = u[M + j];
        du[M + j] = -(r - 2.0 * u[j] + l) * idz2 - p[0] * exp(u[j]);
    }
}
static void bratu_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)t; (void)c;
    const int M = BRATU_M, n = 2 * M;
    const double dz = 1.0 / (M + 1), idz2 = 1.0 / (dz * dz);
    memset(J, 0, sizeof(double) * n * n);
    for (int j = 0; j < M; j++) {
        J[j * n + M + j] = 1.0;
        J[(M + j) * n + j] = 2.0 * idz2 - p[0] * exp(u[j]);
        if (j > 0) J[(M + j) * n + j - 1] = -idz2;
        if (j < M - 1) J[(M + j) * n + j + 1] = -idz2;
    }
}
static void bratu_bc(double *r, const double *U, const double *p, void *c) {
    (void)p; (void)c;
    const int M = BRATU_M, n = 2 * M;
    for (int j = 0; j < M; j++) { r[j] = U[j]; r[M + j] = U[n + j]; }
}
static void bratu_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    const int M = BRATU_M, n = 2 * M;
    memset(d, 0, sizeof(double) * n * 2 * n);
    for (int j = 0; j < M; j++) { d[j * 2 * n + j] = 1.0; d[(M + j) * 2 * n + n + j] = 1.0; }
}

/* ---- 10: Lane-Emden index 1, y' = S y / t + f, f = [y2, -y1], S = [0 0; 0 -2]; y(0) = 1, y(1) = sin 1 ------------- */
static void lane_f(double *du, const double *u, const double *p, double t, void *c) {
    (void)p; (void)t; (void)c;
    du[0] = u[1];
    du[1] = -u[0];
}
static void lane_df(double *J, const double *u, const double *p, double t, void *c) {
    (void)u; (void)p; (void)t; (void)c;
    J[0] = 0.0; J[1] = 1.0; J[2] = -1.0; J[3] = 0.0;
}
static void lane_bc(double *r, const double *U, const double *p, void *c) {
    (void)p; (void)c;
    r[0] = U[0] - 1.0;
    r[1] = U[2] - sin(1.0);
}
static void lane_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    memset(d, 0, sizeof(double) * 2 * 4);
    d[0 * 4 + 0] = 1.0;
    d[1 * 4 + 2] = 1.0;
}
static const double LANE_S[4] = {0.0, 0.0, 0.0, -2.0};

/* ---- 11: u'' = -u with a boundary condition that reads the DERIVATIVE of the interpolant (sol(t, Val{1}) inside bc!,
 *      MIRK/src/interpolation.jl:277-292):  u1(t0) = 0,  u1(t1) - 1 + alpha (u1'(tm) - c) = 0,  tm = (t0 + t1) / 2;
 *      p = [alpha, c].  On [0, pi/2] with c = cos(pi/4) the solution is (sin t, cos t).  The derivative comes from the
 *      Float64 stage buffers: it contributes nothing to the boundary Jacobian (Newton converges linearly in that row). */
static int robin_times(double *tm, const double *p, double t0, double t1, void *c) {
    (void)p; (void)c;
    tm[0] = t0; tm[1] = (t0 + t1) / 2; tm[2] = t1;
    return 3;
}
static void robin_bc(double *r, const double *U, const double *p, void *c) {
    (void)c;
    const double *dU = U + 3 * 2;
    r[0] = U[0];
    r[1] = U[4] - 1.0 + p[0] * (dU[2] - p[1]);
}
static void robin_dbc(double *d, const double *U, const double *p, void *c) {
    (void)U; (void)p; (void)c;
    memset(d, 0, sizeof(double) * 2 * 6);
    d[0 * 6 + 0] = 1.0;
    d[1 * 6 + 4] = 1.0;
}

static const char *NAMES[] = {"pendulum", "linear2", "linear2_tp", "swirling", "lotka",
                              "torus", "layer", "chain8", "chain16", "bratu64", "lane_emden", "robin_sine"};

const char *orc_builtin_name(int id) { return (id >= 0 && id < 12) ? NAMES[id] : 0; }

int orc_builtin_problem(int id, orc_problem *P) {
    memset(P, 0, sizeof(*P));
    switch (id) {
    case 0: *P = (orc_problem){2, 1, 0, 2, 0, pend_f, pend_df, pend_times, pend_bc, pend_dbc, 0}; break;
    case 1: *P = (orc_problem){2, 7, 0, 2, 0, lin_f, lin_df, lin_times, lin_bc, lin_dbc, 0}; break;
    case 2: *P = (orc_problem){2, 3, 1, 2, 1, lin_f, lin_df, ends_times, lintp_bc, lintp_dbc, 0}; break;
    case 3: *P = (orc_problem){6, 1, 0, 6, 0, swirl_f, swirl_df, ends_times, swirl_bc, swirl_dbc, 0}; break;
    case 4: *P = (orc_problem){2, 4, 0, 2, 0, lotka_f, lotka_df, lotka_times, lotka_bc, lotka_dbc, 0}; break;
    case 5: *P = (orc_problem){4, 6, 0, 4, 0, torus_f, torus_df, ends_times, torus_bc, torus_dbc, 0}; break;
    case 6: *P = (orc_problem){2, 1, 0, 2, 0, layer_f, layer_df, ends_times, layer_bc, layer_dbc, 0}; break;
    case 7: *P = (orc_problem){16, 18, 1, 16, 8, chain8_f, chain8_df, ends_times, chain8_bc, chain8_dbc, 0}; break;
    case 8: *P = (orc_problem){32, 34, 1, 32, 16, chain16_f, chain16_df, ends_times, chain16_bc, chain16_dbc, 0}; break;
    case 9: *P = (orc_problem){128, 1, 1, 128, 64, bratu_f, bratu_df, ends_times, bratu_bc, bratu_dbc, 0}; break;
    case 10:
        *P = (orc_problem){2, 0, 1, 2, 1, lane_f, lane_df, ends_times, lane_bc, lane_dbc, 0};
        P->singular_term = LANE_S;
        break;
    case 11:
        *P = (orc_problem){2, 2, 0, 2, 0, lane_f, lane_df, robin_times, robin_bc, robin_dbc, 0};
        P->bc_uses_derivative = 1;
        break;
    default: return -1;
    }
    return 0;
}
