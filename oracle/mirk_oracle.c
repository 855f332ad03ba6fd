/*
 * mirk_oracle.c — CPU restatement of the MIRK4/MIRK6 collocation Newton path.
 * TEST INFRASTRUCTURE ONLY (see mirk_oracle.h).  PARITY UNPINNED against a live Julia run.
 *
 * Every function cites the reference code it restates (paths relative to /root/reference,
 * MIRK/ = lib/BoundaryValueDiffEqMIRK/src/, CORE/ = lib/BoundaryValueDiffEqCore/src/).
 * The Newton loop and the linear solve live in third-party packages that are not vendored
 * (NonlinearSolveFirstOrder "1.2, 2", LinearSolve, FastAlmostBandedMatrices "0.1.4",
 * BandedMatrices "1.7.5"; compat ranges in lib/BoundaryValueDiffEqMIRK/Project.toml); their
 * published algorithms are restated: Newton-Raphson with an |F|_inf <= abstol stop, and a
 * row-pivoted LU elimination of the almost-block-diagonal matrix.
 */
#include "mirk_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* options / tableaus                                                                          */
/* ------------------------------------------------------------------------------------------ */

void orc_default_options(orc_options *o) {
    o->abstol = 1e-6;              /* MIRK/mirk.jl:50 */
    o->adaptive = 1;
    o->defect_threshold = 0.1;     /* CORE/calc_errors.jl:54-60 DefectControl() */
    o->max_num_subintervals = 3000; /* MIRK/algorithms.jl:55-61 */
    o->maxiters = 1000;
    o->reinterp_inplace = 0; /* see DESIGN.md "Q3": 1 reproduces the reference's in-place hazard */
    o->controller = 0; o->ge_method = 0; o->DE = 1.0; o->GE = 1.0;
    o->nlsolve = 0; /* the reference default: NewtonRaphson -> NewtonRaphson + BackTracking -> TrustRegion */
    o->max_outer = 1000; /* safety net only: the reference loop has no cap (mirk.jl:296-322) */
}

/* alg_order (MIRK/alg_utils.jl:1-11): MIRK6I is a 6th-order method selected by its own code */
int orc_convergence_order(int order) { return order == ORC_MIRK6I ? 6 : order; }

/* MIRK/mirk_tableaus.jl:13-34 (MIRK2), :36-60 (MIRK3), :62-87 (MIRK4), :89-118 (MIRK5), :120-152 (MIRK6).
 * Only 1:s of c,v,b is used; x_star[r][j] is the reference's x_star[(j-1)(s*-s) + r] (interpolation.jl:308-309). */
int orc_tableau_get(int order, orc_tableau *T) {
    memset(T, 0, sizeof(*T));
    T->order = order;
    if (order == 2) { /* implicit midpoint rule; interpolant on f(y_i), f(y_{i+1}) */
        T->s = 1;
        T->s_star = 3;
        T->c[0] = 0.5; T->v[0] = 0.5; T->b[0] = 1.0;
        T->c_star[0] = 0.0; T->v_star[0] = 0.0;
        T->c_star[1] = 1.0; T->v_star[1] = 1.0;
        T->tau_star = 0.25;
        return 0;
    }
    if (order == 3) {
        T->s = 2;
        T->s_star = 3;
        T->c[0] = 0.0; T->c[1] = 2.0 / 3.0;
        T->v[0] = 0.0; T->v[1] = 4.0 / 9.0;
        T->b[0] = 1.0 / 4.0; T->b[1] = 3.0 / 4.0;
        T->x[1][0] = 2.0 / 9.0;
        T->c_star[0] = 1.0; T->v_star[0] = 1.0;
        T->tau_star = 0.25;
        return 0;
    }
    if (order == 5) {
        T->s = 4;
        T->s_star = 6;
        const double c[4] = {0.0, 1.0, 3.0 / 4.0, 3.0 / 10.0};
        const double v[4] = {0.0, 1.0, 27.0 / 32.0, 837.0 / 1250.0};
        const double b[4] = {5.0 / 54.0, 1.0 / 14.0, 32.0 / 81.0, 250.0 / 567.0};
        for (int r = 0; r < 4; r++) { T->c[r] = c[r]; T->v[r] = v[r]; T->b[r] = b[r]; }
        T->x[2][0] = 3.0 / 64.0;    T->x[2][1] = -9.0 / 64.0;
        T->x[3][0] = 21.0 / 1000.0; T->x[3][1] = 63.0 / 5000.0; T->x[3][2] = -252.0 / 625.0;
        T->c_star[0] = 4.0 / 5.0;   T->v_star[0] = 4.0 / 5.0;
        T->c_star[1] = 13.0 / 23.0; T->v_star[1] = 13.0 / 23.0;
        T->x_star[0][0] = 14.0 / 1125.0; T->x_star[0][1] = -74.0 / 875.0;
        T->x_star[0][2] = -128.0 / 3375.0; T->x_star[0][3] = 104.0 / 945.0;
        T->x_star[1][0] = 1.0 / 2.0; T->x_star[1][1] = 4508233.0 / 1958887.0;
        T->x_star[1][2] = 48720832.0 / 2518569.0; T->x_star[1][3] = -27646420.0 / 17629983.0;
        T->x_star[1][4] = -11517095.0 / 559682.0;
        T->tau_star = 0.3;
        return 0;
    }
    if (order == 4) {
        T->s = 3;
        T->s_star = 4;
        const double c[3] = {0.0, 1.0, 1.0 / 2.0};
        const double v[3] = {0.0, 1.0, 1.0 / 2.0};
        const double b[3] = {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0};
        for (int r = 0; r < 3; r++) { T->c[r] = c[r]; T->v[r] = v[r]; T->b[r] = b[r]; }
        T->x[2][0] = 1.0 / 8.0;
        T->x[2][1] = -1.0 / 8.0;
        T->c_star[0] = 3.0 / 4.0;
        T->v_star[0] = 27.0 / 32.0;
        T->x_star[0][0] = 3.0 / 64.0;
        T->x_star[0][1] = -9.0 / 64.0;
        T->tau_star = 0.226;
        return 0;
    }
    if (order == 6) {
        T->s = 5;
        T->s_star = 9;
        const double c[5] = {0.0, 1.0, 1.0 / 4.0, 3.0 / 4.0, 1.0 / 2.0};
        const double v[5] = {0.0, 1.0, 5.0 / 32.0, 27.0 / 32.0, 1.0 / 2.0};
        const double b[5] = {7.0 / 90.0, 7.0 / 90.0, 16.0 / 45.0, 16.0 / 45.0, 2.0 / 15.0};
        for (int r = 0; r < 5; r++) { T->c[r] = c[r]; T->v[r] = v[r]; T->b[r] = b[r]; }
        T->x[2][0] = 9.0 / 64.0;  T->x[2][1] = -3.0 / 64.0;
        T->x[3][0] = 3.0 / 64.0;  T->x[3][1] = -9.0 / 64.0;
        T->x[4][0] = -5.0 / 24.0; T->x[4][1] = 5.0 / 24.0;
        T->x[4][2] = 2.0 / 3.0;   T->x[4][3] = -2.0 / 3.0;
        const double cs[4] = {7.0 / 16.0, 3.0 / 8.0, 9.0 / 16.0, 1.0 / 8.0};
        for (int r = 0; r < 4; r++) { T->c_star[r] = cs[r]; T->v_star[r] = cs[r]; }
        const double xs[4][9] = {
            {1547.0 / 32768.0, -1225.0 / 32768.0, 749.0 / 4096.0, -287.0 / 2048.0,
             -861.0 / 16384.0, 0, 0, 0, 0},
            {83.0 / 1536.0, -13.0 / 384.0, 283.0 / 1536.0, -167.0 / 1536.0, -49.0 / 512.0, 0, 0, 0, 0},
            {1225.0 / 32768.0, -1547.0 / 32768.0, 287.0 / 2048.0, -749.0 / 4096.0,
             861.0 / 16384.0, 0, 0, 0, 0},
            {233.0 / 3456.0, -19.0 / 1152.0, 0, 0, 0, -5.0 / 72.0, 7.0 / 72.0, -17.0 / 216.0, 0}};
        memcpy(T->x_star, xs, sizeof(xs));
        T->tau_star = 0.7156;
        return 0;
    }
    if (order == ORC_MIRK6I) { /* MIRK/mirk_tableaus.jl:154-194: Lobatto-type points with irrational coefficients.
                                * `b` and `x_star` are assigned twice there; the second assignment is the one in force. */
        const double s21 = sqrt(21.0), s7 = sqrt(7.0), s3 = sqrt(3.0);
        T->s = 5;
        T->s_star = 8;
        const double c[5] = {0.0, 1.0, 0.5 - s21 / 14.0, 0.5 + s21 / 14.0, 0.5};
        const double v[5] = {0.0, 1.0, 0.5 - 9.0 * s21 / 98.0, 0.5 + 9.0 * s21 / 98.0, 0.5};
        const double b[5] = {1.0 / 20.0, 1.0 / 20.0, 49.0 / 180.0, 49.0 / 180.0, 16.0 / 45.0};
        for (int r = 0; r < 5; r++) { T->c[r] = c[r]; T->v[r] = v[r]; T->b[r] = b[r]; }
        T->x[2][0] = 1.0 / 14.0 + s21 / 98.0;  T->x[2][1] = -1.0 / 14.0 + s21 / 98.0;
        T->x[3][0] = 1.0 / 14.0 - s21 / 98.0;  T->x[3][1] = -1.0 / 14.0 - s21 / 98.0;
        T->x[4][0] = -5.0 / 128.0;             T->x[4][1] = 5.0 / 128.0;
        T->x[4][2] = 7.0 * s21 / 128.0;        T->x[4][3] = -7.0 * s21 / 128.0;
        const double cs[3] = {0.5, 0.5 - s7 / 14.0, 87.0 / 100.0};
        for (int r = 0; r < 3; r++) { T->c_star[r] = cs[r]; T->v_star[r] = cs[r]; }
        T->x_star[0][0] = 1.0 / 64.0;  T->x_star[0][1] = -1.0 / 64.0;
        T->x_star[0][2] = 7.0 / 192.0 * s21;  T->x_star[0][3] = -7.0 / 192.0 * s21;
        T->x_star[1][0] = 3.0 / 112.0 + 9.0 / 1960.0 * s7;
        T->x_star[1][1] = -3.0 / 112.0 + 9.0 / 1960.0 * s7;
        T->x_star[1][2] = 11.0 / 840.0 * s7 + 3.0 / 112.0 * s7 * s3;
        T->x_star[1][3] = 11.0 / 840.0 * s7 - 3.0 / 112.0 * s7 * s3;
        T->x_star[1][4] = 88.0 / 5145.0 * s7;
        T->x_star[1][5] = -18.0 / 343.0 * s7;
        T->x_star[2][0] = 2707592511.0 / 1000000000000.0 - 1006699707.0 / 1000000000000.0 * s7;
        T->x_star[2][1] = -51527976591.0 / 1000000000000.0 - 1006699707.0 / 1000000000000.0 * s7;
        T->x_star[2][2] = -610366393.0 / 75000000000.0 + 7046897949.0 / 1000000000000.0 * s7 +
                          14508670449.0 / 1000000000000.0 * s7 * s3;
        T->x_star[2][3] = -610366393.0 / 75000000000.0 + 7046897949.0 / 1000000000000.0 * s7 -
                          14508670449.0 / 1000000000000.0 * s7 * s3;
        T->x_star[2][4] = -12456457.0 / 1171875000.0 + 1006699707.0 / 109375000000.0 * s7;
        T->x_star[2][5] = 3020099121.0 / 437500000000.0 * s7 + 47328957.0 / 625000000.0;
        T->x_star[2][6] = -7046897949.0 / 250000000000.0 * s7;
        T->tau_star = 0.4;
        return 0;
    }
    return -1;
}

/* MIRK/interpolation.jl:463-527 (orders 2, 3, 4, 5), :528-575 (order 6) and :582-710 (MIRK6I): weights w(tau), w'(tau). */
void orc_interp_weights(int order, double tau, double *w, double *wp) {
    const double t = tau;
    if (order == ORC_MIRK6I) {
        const double s7 = sqrt(7.0), t2 = t * t, t3 = t2 * t, t4 = t2 * t2, t5 = t4 * t, tm1 = t - 1.0;
        /* the quartic shared by the weights of stages 3, 4 and 5, and the common factor of their derivatives */
        const double q = 14000.0 * t4 - 48216.0 * t3 + 1200.0 * s7 * t3 - 3555.0 * s7 * t2 + 62790.0 * t2 +
                         3610.0 * s7 * t - 37450.0 * t + 9135.0 - 1305.0 * s7;
        const double g = (259.0 + 50.0 * s7) * (14.0 * t - 7.0 + s7) * tm1 * (100.0 * t - 87.0) * (2.0 * t - 1.0) * t;
        w[0] = -(12233.0 + 1450.0 * s7) *
               (800086000.0 * t5 + 63579600.0 * s7 * t4 - 2936650584.0 * t4 + 4235152620.0 * t3 -
                201404565.0 * s7 * t3 + 232506630.0 * s7 * t2 - 3033109390.0 * t2 + 1116511695.0 * t -
                116253315.0 * s7 * t + 22707000.0 * s7 - 191568780.0) * t / 2112984835740.0;
        w[1] = -(-10799.0 + 650.0 * s7) *
               (24962000.0 * t4 + 473200.0 * s7 * t3 - 67024328.0 * t3 - 751855.0 * s7 * t2 + 66629600.0 * t2 -
                29507250.0 * t + 236210.0 * s7 * t + 5080365.0 + 50895.0 * s7) * t2 / 29551834260.0;
        w[2] = 7.0 / 1274940.0 * (259.0 + 50.0 * s7) * q * t2;
        w[3] = w[2];
        w[4] = 16.0 / 2231145.0 * (259.0 + 50.0 * s7) * q * t2;
        w[5] = 4.0 / 1227278493.0 * (740.0 * s7 - 6083.0) *
               (1561000.0 * t2 - 2461284.0 * t - 109520.0 * s7 * t + 979272.0 + 86913.0 * s7) * tm1 * tm1 * t2;
        w[6] = -49.0 / 63747.0 * s7 * (20000.0 * t2 - 20000.0 * t + 3393.0) * tm1 * tm1 * t2;
        w[7] = -1250000000.0 / 889206903.0 * (28.0 * t2 - 28.0 * t + 9.0) * tm1 * tm1 * t2;
        wp[0] = (1450.0 * s7 + 12233.0) * (14.0 * t - 7.0 + s7) * tm1 * (-400043.0 * t + 75481.0 + 2083.0 * s7) *
                (100.0 * t - 87.0) * (2.0 * t - 1.0) / 493029795006.0;
        wp[1] = -(650.0 * s7 - 10799.0) * (14.0 * t - 7.0 + s7) * (37443.0 * t - 13762.0 - 2083.0 * s7) *
                (100.0 * t - 87.0) * (2.0 * t - 1.0) * t / 20686283982.0;
        wp[2] = 7.0 / 42498.0 * g;
        wp[3] = wp[2];
        wp[4] = 32.0 / 148743.0 * g;
        wp[5] = 4.0 / 1227278493.0 * (740.0 * s7 - 6083.0) * (14.0 * t - 7.0 + s7) * tm1 * (100.0 * t - 87.0) *
                (6690.0 * t - 4085.0 - 869.0 * s7) * t;
        wp[6] = -98.0 / 21249.0 * s7 * tm1 * (100.0 * t - 13.0) * (100.0 * t - 87.0) * (2.0 * t - 1.0) * t;
        wp[7] = -1250000000.0 / 2074816107.0 * (14.0 * t - 7.0 + s7) * tm1 * (14.0 * t - 7.0 - s7) * (2.0 * t - 1.0) * t;
        return;
    }
    if (order == 2) {
        w[0] = 0.0; w[1] = t * (1.0 - t / 2.0); w[2] = t * t / 2.0;
        wp[0] = 0.0; wp[1] = 1.0 - t; wp[2] = t;
        return;
    }
    if (order == 3) {
        w[0] = t / 4.0 * (2.0 * t * t - 5.0 * t + 4.0);
        w[1] = -3.0 / 4.0 * t * t * (2.0 * t - 3.0);
        w[2] = t * t * (t - 1.0);
        wp[0] = 3.0 / 2.0 * (t - 2.0 / 3.0) * (t - 1.0);
        wp[1] = -9.0 / 2.0 * t * (t - 1.0);
        wp[2] = 3.0 * t * (t - 2.0 / 3.0);
        return;
    }
    if (order == 5) {
        const double t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
        w[0] = t * (22464.0 - 83910.0 * t + 143041.0 * t2 - 113808.0 * t3 + 33256.0 * t4) / 22464.0;
        w[1] = t2 * (-2418.0 + 12303.0 * t - 19512.0 * t2 + 10904.0 * t3) / 3360.0;
        w[2] = -8.0 / 81.0 * t2 * (-78.0 + 209.0 * t - 204.0 * t2 + 8.0 * t3);
        w[3] = -25.0 / 1134.0 * t2 * (-390.0 + 1045.0 * t - 1020.0 * t2 + 328.0 * t3);
        w[4] = -25.0 / 5184.0 * t2 * (390.0 + 255.0 * t - 1680.0 * t2 + 2072.0 * t3);
        w[5] = 279841.0 / 168480.0 * t2 * (-6.0 + 21.0 * t - 24.0 * t2 + 8.0 * t3);
        wp[0] = 1.0 - 13985.0 / 1872.0 * t + 143041.0 / 7488.0 * t2 - 2371.0 / 117.0 * t3 + 20785.0 / 2808.0 * t4;
        wp[1] = -403.0 / 280.0 * t + 12303.0 / 1120.0 * t2 - 813.0 / 35.0 * t3 + 1363.0 / 84.0 * t4;
        wp[2] = 416.0 / 27.0 * t - 1672.0 / 27.0 * t2 + 2176.0 / 27.0 * t3 - 320.0 / 81.0 * t4;
        wp[3] = 3250.0 / 189.0 * t - 26125.0 / 378.0 * t2 + 17000.0 / 189.0 * t3 - 20500.0 / 567.0 * t4;
        wp[4] = -1625.0 / 432.0 * t - 2125.0 / 576.0 * t2 + 875.0 / 27.0 * t3 - 32375.0 / 648.0 * t4;
        wp[5] = -279841.0 / 14040.0 * t + 1958887.0 / 18720.0 * t2 - 279841.0 / 1755.0 * t3 + 279841.0 / 4212.0 * t4;
        return;
    }
    if (order == 4) {
        const double t2 = t * t, tm1 = t - 1.0, t4m3 = t * 4.0 - 3.0, t2m1 = t * 2.0 - 1.0;
        w[0] = -t * (2.0 * t - 3.0) * (2.0 * t2 - 3.0 * t + 2.0) / 6.0;
        w[1] = t2 * (12.0 * t2 - 20.0 * t + 9.0) / 6.0;
        w[2] = 2.0 * t2 * (6.0 * t2 - 14.0 * t + 9.0) / 3.0;
        w[3] = -16.0 * t2 * tm1 * tm1 / 3.0;
        wp[0] = -tm1 * t4m3 * t2m1 / 3.0;
        wp[1] = t * t2m1 * t4m3;
        wp[2] = 4.0 * t * t4m3 * tm1;
        wp[3] = -32.0 * t * t2m1 * tm1 / 3.0;
        return;
    }
    const double t2 = t * t, t3 = t2 * t, t4 = t2 * t2, t5 = t4 * t, t6 = t3 * t3;
    w[0] = t - 28607.0 / 7434.0 * t2 - 166210.0 / 33453.0 * t3 + 334780.0 / 11151.0 * t4 -
           1911296.0 / 55755.0 * t5 + 406528.0 / 33453.0 * t6;
    w[1] = 777.0 / 590.0 * t2 - 2534158.0 / 234171.0 * t3 + 2088580.0 / 78057.0 * t4 -
           10479104.0 / 390285.0 * t5 + 11328512.0 / 1170855.0 * t6;
    w[2] = -1008.0 / 59.0 * t2 + 222176.0 / 1593.0 * t3 - 180032.0 / 531.0 * t4 +
           876544.0 / 2655.0 * t5 - 180224.0 / 1593.0 * t6;
    w[3] = w[2];
    w[4] = -378.0 / 59.0 * t2 + 27772.0 / 531.0 * t3 - 22504.0 / 177.0 * t4 +
           109568.0 / 885.0 * t5 - 22528.0 / 531.0 * t6;
    w[5] = -95232.0 / 413.0 * t2 + 62384128.0 / 33453.0 * t3 - 49429504.0 / 11151.0 * t4 +
           46759936.0 / 11151.0 * t5 - 46661632.0 / 33453.0 * t6;
    w[6] = 896.0 / 5.0 * t2 - 4352.0 / 3.0 * t3 + 3456.0 * t4 - 16384.0 / 5.0 * t5 +
           16384.0 / 15.0 * t6;
    w[7] = 50176.0 / 531.0 * t2 - 179554304.0 / 234171.0 * t3 + 143363072.0 / 78057.0 * t4 -
           136675328.0 / 78057.0 * t5 + 137363456.0 / 234171.0 * t6;
    w[8] = 16384.0 / 441.0 * t3 - 16384.0 / 147.0 * t4 + 16384.0 / 147.0 * t5 - 16384.0 / 441.0 * t6;
    wp[0] = 1.0 - 28607.0 / 3717.0 * t - 166210.0 / 11151.0 * t2 + 1339120.0 / 11151.0 * t3 -
            1911296.0 / 11151.0 * t4 + 813056.0 / 11151.0 * t5;
    wp[1] = 777.0 / 295.0 * t - 2534158.0 / 78057.0 * t2 + 8354320.0 / 78057.0 * t3 -
            10479104.0 / 78057.0 * t4 + 22657024.0 / 390285.0 * t5;
    wp[2] = -2016.0 / 59.0 * t + 222176.0 / 531.0 * t2 - 720128.0 / 531.0 * t3 +
            876544.0 / 531.0 * t4 - 360448.0 / 531.0 * t5;
    wp[3] = wp[2];
    wp[4] = -756.0 / 59.0 * t + 27772.0 / 177.0 * t2 - 90016.0 / 177.0 * t3 + 109568.0 / 177.0 * t4 -
            45056.0 / 177.0 * t5;
    wp[5] = -190464.0 / 413.0 * t + 62384128.0 / 11151.0 * t2 - 197718016.0 / 11151.0 * t3 +
            233799680.0 / 11151.0 * t4 - 93323264.0 / 11151.0 * t5;
    wp[6] = 1792.0 / 5.0 * t - 4352.0 * t2 + 13824.0 * t3 - 16384.0 * t4 + 32768.0 / 5.0 * t5;
    wp[7] = 100352.0 / 531.0 * t - 179554304.0 / 78057.0 * t2 + 573452288.0 / 78057.0 * t3 -
            683376640.0 / 78057.0 * t4 + 274726912.0 / 78057.0 * t5;
    wp[8] = 16384.0 / 147.0 * t2 - 65536.0 / 147.0 * t3 + 81920.0 / 147.0 * t4 - 32768.0 / 147.0 * t5;
}

/* CORE/utils.jl:694 — collect(range(t0; stop=t1, length=nint+1)).  Julia's range is a
 * twice-precision StepRangeLen, i.e. effectively correctly rounded; binary128 here. */
void orc_mesh_uniform(double t0, double t1, int nint, double *mesh) {
    const __float128 a = (__float128)t0, b = (__float128)t1;
    for (int i = 0; i <= nint; i++)
        mesh[i] = (double)(a + ((b - a) * (__float128)i) / (__float128)nint);
    mesh[0] = t0;
    mesh[nint] = t1;
}

/* CORE/utils.jl:119-121 — clamp(searchsortedfirst(mesh,t)-1, 1, N-1), returned 0-based. */
int orc_interval(const double *mesh, int N, double t) {
    int lo = 0, hi = N; /* first j with mesh[j] >= t */
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (mesh[mid] < t) lo = mid + 1; else hi = mid;
    }
    int j = lo; /* == 1-based searchsortedfirst - 1 */
    if (j < 1) j = 1;
    if (j > N - 1) j = N - 1;
    return j - 1;
}

/* ------------------------------------------------------------------------------------------ */
/* collocation residual  (MIRK/collocation.jl:44-72)                      */
/* ------------------------------------------------------------------------------------------ */

/* Threads over mesh intervals for the two interval-parallel loops (Phi and the Jacobian blocks) — used ONLY by the CPU
 * baseline of bench.py (`--impl reference`, "all the host threads it can use").  Default 1: the reference's hot path is
 * single-threaded, and the Python callbacks of the tests' custom problems must not be entered from several threads.
 * Intervals are independent, so the results do not depend on the thread count. */
static int g_interval_threads = 1;
void orc_set_interval_threads(int nthreads) { g_interval_threads = nthreads > 1 ? nthreads : 1; }

static void phi_range(const orc_problem *P, const orc_tableau *T, const double *p, const double *mesh,
                      const double *y, double *Kd, double *phi, int i0, int i1) {
    const int n = P->n, s = T->s;
    double *tmp = (double *)malloc(sizeof(double) * n);
    for (int i = i0; i < i1; i++) {
        const double h = mesh[i + 1] - mesh[i];
        const double *yi = y + (size_t)i * n, *yi1 = yi + n;
        double *K = Kd + (size_t)i * s * n;
        for (int r = 0; r < s; r++) {
            for (int k = 0; k < n; k++) tmp[k] = (1.0 - T->v[r]) * yi[k] + T->v[r] * yi1[k];
            /* __maybe_matmul!(tmp, K[:,1:r-1], x[r,1:r-1], h, 1): tmp += h * K_j * x_rj, j ascending */
            for (int j = 0; j < r; j++)
                for (int k = 0; k < n; k++) tmp[k] = h * K[j * n + k] * T->x[r][j] + tmp[k];
            const double tt = mesh[i] + T->c[r] * h;
            P->f(K + r * n, tmp, p, tt, P->ctx);
            if (P->singular_term && tt > 0.0) /* __add_singular_term!: K_r += S tmp / t (CORE/utils.jl:932-941) */
                for (int k = 0; k < n; k++) {
                    double acc = 0.0;
                    for (int j = 0; j < n; j++) acc += P->singular_term[k * n + j] * tmp[j];
                    K[r * n + k] += acc / tt;
                }
        }
        double *res = phi + (size_t)i * n;
        for (int k = 0; k < n; k++) res[k] = yi1[k] - yi[k];
        for (int r = 0; r < s; r++)
            for (int k = 0; k < n; k++) res[k] = -h * K[r * n + k] * T->b[r] + res[k];
    }
    free(tmp);
}

void orc_phi(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
             const double *y, double *Kd, double *phi) {
    const int ni = N - 1, nt = g_interval_threads;
    if (nt <= 1 || ni < 4 * nt) { phi_range(P, T, p, mesh, y, Kd, phi, 0, ni); return; }
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int c = 0; c < nt; c++)
        phi_range(P, T, p, mesh, y, Kd, phi, (int)((long long)ni * c / nt), (int)((long long)ni * (c + 1) / nt));
}

/* interpolation stages (MIRK/interpolation.jl:300-372; Appendix A.3) */
void orc_interp_setup(const orc_problem *P, const orc_tableau *T, const double *p, int N,
                      const double *mesh, const double *y, const double *Kd, double *Ki) {
    const int n = P->n, s = T->s, si = T->s_star - T->s;
    double *tmp = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < N - 1; i++) {
        const double h = mesh[i + 1] - mesh[i];
        const double *yi = y + (size_t)i * n, *yi1 = yi + n;
        const double *K = Kd + (size_t)i * s * n;
        double *KI = Ki + (size_t)i * si * n;
        for (int r = 0; r < si; r++) {
            for (int k = 0; k < n; k++) tmp[k] = 0.0;
            for (int j = 0; j < s; j++)
                for (int k = 0; k < n; k++) tmp[k] += K[j * n + k] * T->x_star[r][j];
            for (int j = 0; j < r; j++)
                for (int k = 0; k < n; k++) tmp[k] += KI[j * n + k] * T->x_star[r][s + j];
            for (int k = 0; k < n; k++)
                tmp[k] = tmp[k] * h + (1.0 - T->v_star[r]) * yi[k] + T->v_star[r] * yi1[k];
            P->f(KI + r * n, tmp, p, mesh[i] + T->c_star[r] * h, P->ctx);
        }
    }
    free(tmp);
}

/* continuous extension (MIRK/interpolation.jl:98-126 for sol(t), :214-242 for the BC-time EvalSol;
 * Appendix A.4).  bc_shortcut=1 reproduces EvalSol's endpoint short-circuit (:218-219). */
void orc_eval_sol(const orc_problem *P, const orc_tableau *T, int N, const double *mesh,
                  const double *y, const double *Kd, const double *Ki, double t, int deriv,
                  int bc_shortcut, double *out) {
    const int n = P->n, s = T->s, si = T->s_star - T->s;
    if (bc_shortcut && deriv == 0) {
        if (t == mesh[0]) { memcpy(out, y, sizeof(double) * n); return; }
        if (t == mesh[N - 1]) { memcpy(out, y + (size_t)(N - 1) * n, sizeof(double) * n); return; }
    }
    const int i = orc_interval(mesh, N, t);
    const double dt = mesh[i + 1] - mesh[i];
    const double tau = (t - mesh[i]) / dt;
    double w[ORC_MAX_SS], wp[ORC_MAX_SS];
    orc_interp_weights(T->order, tau, w, wp);
    const double *ww = deriv ? wp : w;
    const double *K = Kd + (size_t)i * s * n, *KI = Ki + (size_t)i * si * n;
    for (int k = 0; k < n; k++) out[k] = 0.0;
    for (int r = 0; r < s; r++)
        for (int k = 0; k < n; k++) out[k] += K[r * n + k] * ww[r];
    for (int r = 0; r < si; r++)
        for (int k = 0; k < n; k++) out[k] += KI[r * n + k] * ww[s + r];
    if (!deriv)
        for (int k = 0; k < n; k++) out[k] = out[k] * dt + y[(size_t)i * n + k];
}

/* gather U[k] = sol(times[k]) for the boundary condition */
static int bc_gather(const orc_problem *P, const orc_tableau *T, const double *p, int N,
                     const double *mesh, const double *y, const double *Kd, const double *Ki,
                     double *times, double *U) {
    const int n = P->n;
    if (P->problem_type == 1) { /* TwoPoint: bca(res_a, y_1, p), bcb(res_b, y_N, p) CORE/utils.jl:157-199 */
        memcpy(U, y, sizeof(double) * n);
        memcpy(U + n, y + (size_t)(N - 1) * n, sizeof(double) * n);
        times[0] = mesh[0];
        times[1] = mesh[N - 1];
        return 2;
    }
    const int m = P->bc_times(times, p, mesh[0], mesh[N - 1], P->ctx);
    for (int k = 0; k < m; k++) orc_eval_sol(P, T, N, mesh, y, Kd, Ki, times[k], 0, 1, U + k * n);
    /* sol(t, Val{1}) inside bc! (interpolation.jl:277-292): no end-point short cut, interval(t) and the derivative
     * weights on the stage buffers of that interval */
    if (P->bc_uses_derivative)
        for (int k = 0; k < m; k++) orc_eval_sol(P, T, N, mesh, y, Kd, Ki, times[k], 1, 0, U + (size_t)(m + k) * n);
    return m;
}

/* full residual (MIRK/mirk.jl:471-534).  Standard: [bc; Phi_1..Phi_{N-1}] and interp_setup! on
 * every call (quirk Q7, interpolation.jl:382-403); TwoPoint: [bc_a; Phi...; bc_b]
 * (CORE/utils.jl:42-52). */
void orc_loss(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
              const double *y, double *Kd, double *Ki, double *resid) {
    const int n = P->n, L = P->n_bc;
    double times[ORC_MAX_BC_PTS];
    double *U = (double *)malloc(sizeof(double) * 2 * ORC_MAX_BC_PTS * n);
    double *bc = (double *)malloc(sizeof(double) * L);
    if (P->problem_type == 0) {
        orc_phi(P, T, p, N, mesh, y, Kd, resid + L);
        orc_interp_setup(P, T, p, N, mesh, y, Kd, Ki);
        bc_gather(P, T, p, N, mesh, y, Kd, Ki, times, U);
        P->bc(bc, U, p, P->ctx);
        memcpy(resid, bc, sizeof(double) * L);
    } else {
        const int La = P->n_bca;
        orc_phi(P, T, p, N, mesh, y, Kd, resid + La);
        bc_gather(P, T, p, N, mesh, y, Kd, Ki, times, U);
        P->bc(bc, U, p, P->ctx);
        memcpy(resid, bc, sizeof(double) * La);
        memcpy(resid + La + (size_t)(N - 1) * n, bc + La, sizeof(double) * (L - La));
    }
    free(U);
    free(bc);
}

/* ------------------------------------------------------------------------------------------ */
/* exact block Jacobian  (what sparse ForwardDiff of loss_collocation yields when the pattern  */
/* is wide enough: MIRK/mirk.jl:810-838; Appendix A.2)                                         */
/* ------------------------------------------------------------------------------------------ */

static void matmul_nn(int n, const double *A, const double *B, double *C) { /* C = A*B row-major */
    /* column blocks of 16 / 4 / 1 with the accumulators of a block held in registers across the k loop (unit-stride,
     * vectorisable).  Every C[i][j] still accumulates its products in ascending k from +0.0, so the result is
     * bit-identical to the textbook i-j-k loop (no FMA contraction: see the Makefile) */
    for (int i = 0; i < n; i++) {
        const double *restrict Ai = A + (size_t)i * n;
        double *restrict Ci = C + (size_t)i * n;
        int j0 = 0;
        for (; j0 + 16 <= n; j0 += 16) {
            double acc[16];
            for (int jj = 0; jj < 16; jj++) acc[jj] = 0.0;
            for (int k = 0; k < n; k++) {
                const double a = Ai[k];
                const double *restrict Bk = B + (size_t)k * n + j0;
                for (int jj = 0; jj < 16; jj++) acc[jj] += a * Bk[jj];
            }
            for (int jj = 0; jj < 16; jj++) Ci[j0 + jj] = acc[jj];
        }
        for (; j0 + 4 <= n; j0 += 4) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = 0; k < n; k++) {
                const double a = Ai[k];
                const double *restrict Bk = B + (size_t)k * n + j0;
                for (int jj = 0; jj < 4; jj++) acc[jj] += a * Bk[jj];
            }
            for (int jj = 0; jj < 4; jj++) Ci[j0 + jj] = acc[jj];
        }
        for (; j0 < n; j0++) {
            double acc = 0.0;
            for (int k = 0; k < n; k++) acc += Ai[k] * B[(size_t)k * n + j0];
            Ci[j0] = acc;
        }
    }
}

static void jac_blocks_range(const orc_problem *P, const orc_tableau *T, const double *p,
                             const double *mesh, const double *y, double *Lb, double *Rb, int i0, int i1) {
    const int n = P->n, s = T->s, nn = n * n;
    double *tmp = (double *)malloc(sizeof(double) * n);
    double *K = (double *)malloc(sizeof(double) * s * n);
    double *J = (double *)malloc(sizeof(double) * nn);
    double *M = (double *)malloc(sizeof(double) * nn);
    double *A = (double *)malloc(sizeof(double) * s * nn); /* dK_r/dy_i     */
    double *B = (double *)malloc(sizeof(double) * s * nn); /* dK_r/dy_{i+1} */
    int zA[16], zB[16];                                    /* stage derivative structurally zero */
    for (int i = i0; i < i1; i++) {
        const double h = mesh[i + 1] - mesh[i];
        const double *yi = y + (size_t)i * n, *yi1 = yi + n;
        for (int r = 0; r < s; r++) {
            for (int k = 0; k < n; k++) tmp[k] = (1.0 - T->v[r]) * yi[k] + T->v[r] * yi1[k];
            for (int j = 0; j < r; j++)
                for (int k = 0; k < n; k++) tmp[k] = h * K[j * n + k] * T->x[r][j] + tmp[k];
            const double tt = mesh[i] + T->c[r] * h;
            P->f(K + r * n, tmp, p, tt, P->ctx);
            P->dfdu(J, tmp, p, tt, P->ctx);
            if (P->singular_term && tt > 0.0) {
                for (int k = 0; k < n; k++) {
                    double acc = 0.0;
                    for (int j = 0; j < n; j++) acc += P->singular_term[k * n + j] * tmp[j];
                    K[r * n + k] += acc / tt;
                }
                for (int e = 0; e < nn; e++) J[e] += P->singular_term[e] / tt;
            }
            /* A_r = J_r [(1-v_r) I + h sum_j x_rj A_j];  a stage whose bracket is structurally zero (K_1 does not depend
             * on y_{i+1}, a stage with v_r = 1 and no coupling not on y_i, ...) is the zero matrix: the product is skipped,
             * which leaves the same bits as multiplying J_r by zeros */
            int zero = (1.0 - T->v[r] == 0.0);
            for (int j = 0; j < r; j++) if (T->x[r][j] != 0.0 && !zA[j]) zero = 0;
            zA[r] = zero;
            int coupled = 0;
            for (int j = 0; j < r; j++) if (T->x[r][j] != 0.0 && !zA[j]) coupled = 1;
            if (zero) {
                for (int e = 0; e < nn; e++) A[r * nn + e] = 0.0;
            } else if (!coupled) {
                /* bracket = d I: J_r (d I) = d J_r entry by entry; "0.0 +" keeps the +0.0 the product loop leaves */
                const double d = 1.0 - T->v[r];
                for (int e = 0; e < nn; e++) A[r * nn + e] = 0.0 + J[e] * d;
            } else {
                for (int e = 0; e < nn; e++) M[e] = 0.0;
                for (int k = 0; k < n; k++) M[k * n + k] = 1.0 - T->v[r];
                for (int j = 0; j < r; j++)
                    if (T->x[r][j] != 0.0) {
                        const double hx = h * T->x[r][j];  /* (h * x) * A: the product order of the plain expression */
                        const double *restrict Aj = A + (size_t)j * nn;
                        double *restrict Mw = M;
                        for (int e = 0; e < nn; e++) Mw[e] += hx * Aj[e];
                    }
                matmul_nn(n, J, M, A + r * nn);
            }
            /* B_r = J_r [v_r I + h sum_j x_rj B_j] */
            zero = (T->v[r] == 0.0);
            for (int j = 0; j < r; j++) if (T->x[r][j] != 0.0 && !zB[j]) zero = 0;
            zB[r] = zero;
            coupled = 0;
            for (int j = 0; j < r; j++) if (T->x[r][j] != 0.0 && !zB[j]) coupled = 1;
            if (zero) {
                for (int e = 0; e < nn; e++) B[r * nn + e] = 0.0;
            } else if (!coupled) {
                const double d = T->v[r];
                for (int e = 0; e < nn; e++) B[r * nn + e] = 0.0 + J[e] * d;
            } else {
                for (int e = 0; e < nn; e++) M[e] = 0.0;
                for (int k = 0; k < n; k++) M[k * n + k] = T->v[r];
                for (int j = 0; j < r; j++)
                    if (T->x[r][j] != 0.0) {
                        const double hx = h * T->x[r][j];
                        const double *restrict Bj = B + (size_t)j * nn;
                        double *restrict Mw = M;
                        for (int e = 0; e < nn; e++) Mw[e] += hx * Bj[e];
                    }
                matmul_nn(n, J, M, B + r * nn);
            }
        }
        double *Li = Lb + (size_t)i * nn, *Ri = Rb + (size_t)i * nn;
        for (int e = 0; e < nn; e++) { Li[e] = 0.0; Ri[e] = 0.0; }
        for (int k = 0; k < n; k++) { Li[k * n + k] = -1.0; Ri[k * n + k] = 1.0; }
        for (int r = 0; r < s; r++) {
            const double hb = h * T->b[r];
            const double *restrict Ar = A + (size_t)r * nn, *restrict Br = B + (size_t)r * nn;
            double *restrict Lw = Li, *restrict Rw = Ri;
            for (int e = 0; e < nn; e++) {
                Lw[e] -= hb * Ar[e];
                Rw[e] -= hb * Br[e];
            }
        }
    }
    free(tmp); free(K); free(J); free(M); free(A); free(B);
}

void orc_jac_blocks(const orc_problem *P, const orc_tableau *T, const double *p, int N,
                    const double *mesh, const double *y, double *Lb, double *Rb) {
    const int ni = N - 1, nt = g_interval_threads;
    if (nt <= 1 || ni < 4 * nt) { jac_blocks_range(P, T, p, mesh, y, Lb, Rb, 0, ni); return; }
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int c = 0; c < nt; c++)
        jac_blocks_range(P, T, p, mesh, y, Lb, Rb, (int)((long long)ni * c / nt), (int)((long long)ni * (c + 1) / nt));
}

/* BC Jacobian in the reference's pattern (quirk Q2): the interpolant inside loss_bc reads the
 * Float64 stage buffers (MIRK/interpolation.jl:227), so d bc / d y lands only on the LEFT node of
 * the interval that contains the evaluation time (identity through `z .* dt .+ u[ii]`, :239);
 * endpoints short-circuit to y_1 / y_N (:218-219).  Returns m; nodes[k] is 0-based; B is
 * m blocks of L×n row-major. */
int orc_bc_jac(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
               const double *y, const double *Kd, const double *Ki, int *nodes, double *B) {
    const int n = P->n, L = P->n_bc;
    double times[ORC_MAX_BC_PTS];
    double *U = (double *)malloc(sizeof(double) * 2 * ORC_MAX_BC_PTS * n);
    const int m = bc_gather(P, T, p, N, mesh, y, Kd, Ki, times, U);
    double *d = (double *)malloc(sizeof(double) * L * m * n);
    P->dbc(d, U, p, P->ctx);
    for (int k = 0; k < m; k++) {
        if (times[k] == mesh[0]) nodes[k] = 0;
        else if (times[k] == mesh[N - 1]) nodes[k] = N - 1;
        else nodes[k] = orc_interval(mesh, N, times[k]);
        for (int r = 0; r < L; r++)
            for (int c = 0; c < n; c++) B[((size_t)k * L + r) * n + c] = d[(size_t)r * m * n + k * n + c];
    }
    free(U);
    free(d);
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* almost-block-diagonal solve: sequential row-pivoted LU with carried boundary rows           */
/* (stands in for LinearSolve on AlmostBandedMatrix / BandedMatrix gbtrf; call site            */
/* CORE/default_internal_solve.jl:107-110 -> NonlinearSolve)                                   */
/* ------------------------------------------------------------------------------------------ */

int orc_abd_solve(int n, int N, int L, const double *Lb, const double *Rb, int m, const int *nodes,
                  const double *B, const double *rhs_bc, const double *rhs_phi, double *delta) {
    if (L != n) return -1;
    /* unique sorted pinned nodes with accumulated blocks */
    int Q = 0;
    int *pn = (int *)malloc(sizeof(int) * (m > 0 ? m : 1));
    for (int k = 0; k < m; k++) {
        int found = 0;
        for (int q = 0; q < Q; q++) if (pn[q] == nodes[k]) found = 1;
        if (!found) pn[Q++] = nodes[k];
    }
    for (int a = 0; a < Q; a++)
        for (int b = a + 1; b < Q; b++)
            if (pn[b] < pn[a]) { int t = pn[a]; pn[a] = pn[b]; pn[b] = t; }
    const int W = Q * n;            /* tail width */
    const int RW = 2 * n + W + 1;   /* row width: cur | nxt | tail | rhs */
    const int RA = L + n;           /* active rows */
    double *act = (double *)calloc((size_t)RA * RW, sizeof(double));
    double *piv = (double *)malloc(sizeof(double) * (size_t)(N - 1) * n * RW);
    int *slot_of = (int *)malloc(sizeof(int) * N);
    for (int i = 0; i < N; i++) slot_of[i] = -1;
    for (int q = 0; q < Q; q++) slot_of[pn[q]] = q;
    int status = 0;

    for (int r = 0; r < L; r++) {
        double *row = act + (size_t)r * RW;
        for (int k = 0; k < m; k++) {
            const int q = slot_of[nodes[k]];
            for (int c = 0; c < n; c++) row[2 * n + q * n + c] += B[((size_t)k * L + r) * n + c];
        }
        row[RW - 1] = rhs_bc[r];
    }
    for (int k = 0; k < N - 1 && !status; k++) {
        if (slot_of[k] >= 0) {
            const int q = slot_of[k];
            for (int r = 0; r < L; r++) {
                double *row = act + (size_t)r * RW;
                for (int c = 0; c < n; c++) { row[c] += row[2 * n + q * n + c]; row[2 * n + q * n + c] = 0.0; }
            }
        }
        for (int r = 0; r < n; r++) {
            double *row = act + (size_t)(L + r) * RW;
            memset(row, 0, sizeof(double) * RW);
            for (int c = 0; c < n; c++) {
                row[c] = Lb[((size_t)k * n + r) * n + c];
                row[n + c] = Rb[((size_t)k * n + r) * n + c];
            }
            row[RW - 1] = rhs_phi[(size_t)k * n + r];
        }
        for (int j = 0; j < n; j++) {
            int pr = j;
            double best = fabs(act[(size_t)j * RW + j]);
            for (int r = j + 1; r < RA; r++) {
                const double a = fabs(act[(size_t)r * RW + j]);
                if (a > best) { best = a; pr = r; }
            }
            if (!(best > 0.0) || !isfinite(best)) { status = 1; break; }
            if (pr != j)
                for (int c = 0; c < RW; c++) {
                    const double t = act[(size_t)j * RW + c];
                    act[(size_t)j * RW + c] = act[(size_t)pr * RW + c];
                    act[(size_t)pr * RW + c] = t;
                }
            const double *restrict prow = act + (size_t)j * RW;
            for (int r = j + 1; r < RA; r++) {
                double *restrict row = act + (size_t)r * RW;  /* r != j: the rows do not overlap */
                const double f = row[j] / prow[j];
                if (f != 0.0) {
                    row[j] = 0.0;
                    for (int c = j + 1; c < RW; c++) row[c] -= f * prow[c];
                }
            }
        }
        if (status) break;
        memcpy(piv + (size_t)k * n * RW, act, sizeof(double) * (size_t)n * RW);
        for (int r = 0; r < L; r++) {
            double *dst = act + (size_t)r * RW;
            const double *src = act + (size_t)(n + r) * RW;
            for (int c = 0; c < n; c++) { dst[c] = src[n + c]; dst[n + c] = 0.0; }
            for (int c = 2 * n; c < RW; c++) dst[c] = src[c];
        }
    }
    if (!status) {
        const int k = N - 1;
        if (slot_of[k] >= 0) {
            const int q = slot_of[k];
            for (int r = 0; r < L; r++) {
                double *row = act + (size_t)r * RW;
                for (int c = 0; c < n; c++) { row[c] += row[2 * n + q * n + c]; row[2 * n + q * n + c] = 0.0; }
            }
        }
        for (int j = 0; j < n && !status; j++) {
            int pr = j;
            double best = fabs(act[(size_t)j * RW + j]);
            for (int r = j + 1; r < L; r++) {
                const double a = fabs(act[(size_t)r * RW + j]);
                if (a > best) { best = a; pr = r; }
            }
            if (!(best > 0.0) || !isfinite(best)) { status = 1; break; }
            if (pr != j)
                for (int c = 0; c < RW; c++) {
                    const double t = act[(size_t)j * RW + c];
                    act[(size_t)j * RW + c] = act[(size_t)pr * RW + c];
                    act[(size_t)pr * RW + c] = t;
                }
            const double *prow = act + (size_t)j * RW;
            for (int r = j + 1; r < L; r++) {
                double *row = act + (size_t)r * RW;
                const double f = row[j] / prow[j];
                row[j] = 0.0;
                for (int c = j + 1; c < RW; c++) row[c] -= f * prow[c];
            }
        }
        if (!status) {
            double *dk = delta + (size_t)k * n;
            for (int j = n - 1; j >= 0; j--) {
                const double *row = act + (size_t)j * RW;
                double acc = row[RW - 1];
                for (int c = j + 1; c < n; c++) acc -= row[c] * dk[c];
                dk[j] = acc / row[j];
            }
            for (int kk = N - 2; kk >= 0; kk--) {
                double *d = delta + (size_t)kk * n;
                const double *dn = d + n;
                for (int j = n - 1; j >= 0; j--) {
                    const double *row = piv + ((size_t)kk * n + j) * RW;
                    double acc = row[RW - 1];
                    for (int c = 0; c < n; c++) acc -= row[n + c] * dn[c];
                    for (int q = 0; q < Q; q++)
                        if (pn[q] > kk)
                            for (int c = 0; c < n; c++)
                                acc -= row[2 * n + q * n + c] * delta[(size_t)pn[q] * n + c];
                    for (int c = j + 1; c < n; c++) acc -= row[c] * d[c];
                    d[j] = acc / row[j];
                }
            }
        }
    }
    free(pn); free(act); free(piv); free(slot_of);
    return status;
}

/* ------------------------------------------------------------------------------------------ */
/* Newton-Raphson (NonlinearSolveFirstOrder NewtonRaphson, no line search; restated from its   */
/* published algorithm: J(u) d = F(u); u <- u - d; stop when |F(u)|_inf <= abstol; non-finite   */
/* norm => Unstable; the best iterate is what is returned (AbsNormSafeBest)).                   */
/* ------------------------------------------------------------------------------------------ */

static double norm_inf(const double *x, size_t len) {
    double m = 0.0;
    for (size_t i = 0; i < len; i++) {
        const double a = fabs(x[i]);
        if (!(a <= m)) m = a; /* propagates NaN */
    }
    return m;
}

/* ---- the reference's DEFAULT nonlinear solver: NonlinearSolvePolyAlgorithm(NewtonRaphson, NewtonRaphson +
 * BackTracking, TrustRegion), CORE/src/default_internal_solve.jl:31-45, each sub-solver restarting from the original
 * u0.  The sub-solvers live in third-party packages that are not under /root/reference (NonlinearSolveFirstOrder,
 * NonlinearSolveBase, LineSearch / LineSearches): what follows restates their published algorithms FROM MEMORY —
 * parity unpinned until julia/parity/dump_reference.jl has been run on a Julia box.
 *   termination  AbsNormSafeBest on |F|_inf (NonlinearSolveBase termination_conditions): Success when the objective
 *                <= abstol; Unstable when it is not finite; the best iterate is tracked and returned; Stalled when,
 *                after patience_steps = 100 steps with the objective within 3 abstol, the window's min and max are
 *                within a factor 1.3, or after max_stalled_steps = 32 consecutive steps with |u - u_prev|_2 <= abstol
 *                and <= reltol |u|_2 (reltol = eps^(4/5));
 *   NewtonRaphson            J d = F, u <- u - d
 *   + BackTracking           LineSearches.BackTracking(order 3, c_1 = 1e-4, rho_hi = 0.5, rho_lo = 0.1, 1000 iterations)
 *                            on phi(alpha) = |F(u + alpha du)|_2^2 / 2 with phi'(0) = F . (J du); a failed search ends
 *                            the sub-solver (InternalLineSearchFailed)
 *   TrustRegion              Simple radius update: Delta_max = max(|F|_2, max(u) - min(u)), Delta_0 = Delta_max / 11,
 *                            dogleg step (Newton / Cauchy / their blend), rho = actual / predicted reduction, accept when
 *                            rho > 1e-4, shrink by 1/4 when rho < 1/4, expand by 2 when rho > 3/4, give up after 32
 *                            consecutive shrinks (ShrinkThresholdExceeded)
 *   polyalgorithm            first success wins; if all fail, the sub-solver with the smallest final |F|_inf provides
 *                            the iterate and the return code. */

typedef struct {
    const orc_problem *P; const orc_tableau *T; const double *p; int N; const double *mesh;
    int n, L, La; size_t nu, nr;
    double *Kd, *Ki;
    double *Lb, *Rb, *B; int nodes[ORC_MAX_BC_PTS]; int m;   /* Jacobian at the current iterate */
} nl_ctx;

/* rows of the residual vector: Standard [bc(L); Phi], TwoPoint [bc_a(La); Phi; bc_b] */
static size_t nl_bc_row(const nl_ctx *c, int q) { return q < c->La ? (size_t)q : (size_t)c->La + (size_t)(c->N - 1) * c->n + (q - c->La); }

static void nl_jacobian(nl_ctx *c, const double *y) {
    orc_jac_blocks(c->P, c->T, c->p, c->N, c->mesh, y, c->Lb, c->Rb);
    c->m = orc_bc_jac(c->P, c->T, c->p, c->N, c->mesh, y, c->Kd, c->Ki, c->nodes, c->B);
}
/* out = J v  (out has nr entries, v has nu) */
static void nl_jvec(const nl_ctx *c, const double *v, double *out) {
    const int n = c->n, L = c->L;
    for (int q = 0; q < L; q++) {
        double acc = 0.0;
        for (int k = 0; k < c->m; k++)
            for (int j = 0; j < n; j++) acc += c->B[((size_t)k * L + q) * n + j] * v[(size_t)c->nodes[k] * n + j];
        out[nl_bc_row(c, q)] = acc;
    }
    for (int i = 0; i < c->N - 1; i++)
        for (int r = 0; r < n; r++) {
            double acc = 0.0;
            for (int j = 0; j < n; j++)
                acc += c->Lb[((size_t)i * n + r) * n + j] * v[(size_t)i * n + j] + c->Rb[((size_t)i * n + r) * n + j] * v[(size_t)(i + 1) * n + j];
            out[(size_t)c->La + (size_t)i * n + r] = acc;
        }
}
/* out = J^T w  (w has nr entries, out has nu) */
static void nl_jtvec(const nl_ctx *c, const double *w, double *out) {
    const int n = c->n, L = c->L;
    for (size_t e = 0; e < c->nu; e++) out[e] = 0.0;
    for (int q = 0; q < L; q++)
        for (int k = 0; k < c->m; k++)
            for (int j = 0; j < n; j++) out[(size_t)c->nodes[k] * n + j] += c->B[((size_t)k * L + q) * n + j] * w[nl_bc_row(c, q)];
    for (int i = 0; i < c->N - 1; i++)
        for (int r = 0; r < n; r++) {
            const double wr = w[(size_t)c->La + (size_t)i * n + r];
            for (int j = 0; j < n; j++) {
                out[(size_t)i * n + j] += c->Lb[((size_t)i * n + r) * n + j] * wr;
                out[(size_t)(i + 1) * n + j] += c->Rb[((size_t)i * n + r) * n + j] * wr;
            }
        }
}
/* delta = J \ fu with the Jacobian stored in c */
static int nl_linsolve(const nl_ctx *c, const double *fu, double *delta, double *rhs_bc) {
    const int L = c->L, La = c->La;
    for (int q = 0; q < L; q++) rhs_bc[q] = fu[nl_bc_row(c, q)];
    return orc_abd_solve(c->n, c->N, L, c->Lb, c->Rb, c->m, c->nodes, c->B, rhs_bc, fu + La, delta);
}
static double nl_dot(const double *a, const double *b, size_t len) {
    double s = 0.0;
    for (size_t i = 0; i < len; i++) s += a[i] * b[i];
    return s;
}

/* AbsNormSafeBest termination state */
typedef struct {
    double abstol, reltol, best, initial, trace[100];
    int nsteps, stall_counter;
    double *ubest; int have_best;
} nl_term;
static void nl_term_init(nl_term *t, double abstol, double *ubest) {
    memset(t, 0, sizeof(*t));
    t->abstol = abstol; t->reltol = pow(2.220446049250313e-16, 0.8); t->best = INFINITY; t->ubest = ubest;
}
/* returns -1 to continue, else the return code that ends the sub-solver */
static int nl_term_check(nl_term *t, double objective, const double *u, const double *uprev, size_t nu) {
    if (!isfinite(objective)) return ORC_UNSTABLE;
    if (objective < t->best) { t->best = objective; memcpy(t->ubest, u, sizeof(double) * nu); t->have_best = 1; }
    if (objective <= t->abstol) return ORC_SUCCESS;
    t->nsteps++;
    if (t->nsteps == 1) t->initial = objective;
    t->trace[(t->nsteps - 1) % 100] = objective;
    if (objective <= 3.0 * t->abstol && t->nsteps >= 100) {
        const int cnt = t->nsteps < 100 ? t->nsteps : 100;
        double mn = INFINITY, mx = -INFINITY;
        for (int i = 0; i < cnt; i++) { if (t->trace[i] < mn) mn = t->trace[i]; if (t->trace[i] > mx) mx = t->trace[i]; }
        if (mn < 1.3 * mx) return ORC_STALLED;
    }
    double du2 = 0.0, u2 = 0.0;
    for (size_t i = 0; i < nu; i++) { const double d = u[i] - uprev[i]; du2 += d * d; u2 += u[i] * u[i]; }
    du2 = sqrt(du2); u2 = sqrt(u2);
    if (du2 <= t->abstol && du2 <= t->reltol * u2) t->stall_counter++; else t->stall_counter = 0;
    if (t->stall_counter >= 32) return ORC_STALLED;
    return -1;
}

/* LineSearches.BackTracking (order 3).  phi(alpha) evaluates F(y + alpha du) into fu_trial.  Returns alpha, or NAN. */
typedef struct { nl_ctx *c; const double *y; const double *du; double *ytrial; double *fu_trial; } nl_phi_ctx;
static double nl_phi(nl_phi_ctx *k, double alpha) {
    for (size_t i = 0; i < k->c->nu; i++) k->ytrial[i] = k->y[i] + alpha * k->du[i];
    orc_loss(k->c->P, k->c->T, k->c->p, k->c->N, k->c->mesh, k->ytrial, k->c->Kd, k->c->Ki, k->fu_trial);
    return 0.5 * nl_dot(k->fu_trial, k->fu_trial, k->c->nr);
}
static double nan_min(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : (a < b ? a : b); }
static double nan_max(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : (a > b ? a : b); }
static double nl_backtracking(nl_phi_ctx *k, double phi_0, double dphi_0) {
    const double c_1 = 1e-4, rho_hi = 0.5, rho_lo = 0.1;
    double a1 = 1.0, a2 = 1.0, phx0 = phi_0, phx1 = nl_phi(k, a1);
    int iterfinite = 0;
    while (!isfinite(phx1) && iterfinite < 52) { iterfinite++; a1 = a2; a2 = a1 / 2.0; phx1 = nl_phi(k, a2); }
    int iteration = 0;
    while (phx1 > phi_0 + c_1 * a2 * dphi_0) {
        iteration++;
        if (iteration > 1000) return NAN;
        double atmp;
        if (iteration == 1) {
            atmp = -(dphi_0 * a2 * a2) / (2.0 * (phx1 - phi_0 - dphi_0 * a2));
        } else {
            const double div = 1.0 / (a1 * a1 * a2 * a2 * (a2 - a1));
            const double a = (a1 * a1 * (phx1 - phi_0 - dphi_0 * a2) - a2 * a2 * (phx0 - phi_0 - dphi_0 * a1)) * div;
            const double b = (-a1 * a1 * a1 * (phx1 - phi_0 - dphi_0 * a2) + a2 * a2 * a2 * (phx0 - phi_0 - dphi_0 * a1)) * div;
            if (fabs(a) <= 2.220446049250313e-16) atmp = dphi_0 / (2.0 * b);
            else { double d = b * b - 3.0 * a * dphi_0; if (d < 0.0) d = 0.0; atmp = (-b + sqrt(d)) / (3.0 * a); }
        }
        a1 = a2;
        atmp = nan_min(atmp, a2 * rho_hi);
        a2 = nan_max(atmp, a2 * rho_lo);
        phx0 = phx1;
        phx1 = nl_phi(k, a2);
        if (isnan(a2)) return NAN;
    }
    return a2;
}

/* one sub-solver from the iterate in y; alg 1 NewtonRaphson, 2 + BackTracking, 3 TrustRegion */
static int nl_run(nl_ctx *c, int alg, double *y, double abstol, int maxiters, double *resid_norm, int *iters) {
    const size_t nu = c->nu, nr = c->nr;
    double *fu = (double *)malloc(sizeof(double) * nr), *fu2 = (double *)malloc(sizeof(double) * nr);
    double *du = (double *)malloc(sizeof(double) * nu), *uprev = (double *)malloc(sizeof(double) * nu);
    double *ytr = (double *)malloc(sizeof(double) * nu), *ubest = (double *)malloc(sizeof(double) * nu);
    double *g = (double *)malloc(sizeof(double) * nu), *Jg = (double *)malloc(sizeof(double) * nr);
    double *rhs_bc = (double *)malloc(sizeof(double) * c->L);
    nl_term term;
    nl_term_init(&term, abstol, ubest);
    int ret = ORC_MAXITERS, it = 0;
    orc_loss(c->P, c->T, c->p, c->N, c->mesh, y, c->Kd, c->Ki, fu);
    double nrm = norm_inf(fu, nr);
    /* trust region state */
    double Delta = 0.0, Delta_max = 0.0;
    int shrink_counter = 0, have_jac = 0;
    if (alg == 3) {
        double umax = -INFINITY, umin = INFINITY;
        for (size_t i = 0; i < nu; i++) { if (y[i] > umax) umax = y[i]; if (y[i] < umin) umin = y[i]; }
        const double f2 = sqrt(nl_dot(fu, fu, nr));
        Delta_max = f2 > umax - umin ? f2 : umax - umin;
        Delta = Delta_max / 11.0;
    }
    for (; it < maxiters;) {
        memcpy(uprev, y, sizeof(double) * nu);
        if (alg != 3 || !have_jac) { nl_jacobian(c, y); have_jac = 1; }
        if (nl_linsolve(c, fu, du, rhs_bc)) { ret = ORC_FAILURE; break; }
        for (size_t i = 0; i < nu; i++) du[i] = -du[i];   /* Newton direction */
        if (alg == 1) {
            for (size_t i = 0; i < nu; i++) y[i] += du[i];
            orc_loss(c->P, c->T, c->p, c->N, c->mesh, y, c->Kd, c->Ki, fu);
        } else if (alg == 2) {
            nl_jvec(c, du, Jg);
            const double phi_0 = 0.5 * nl_dot(fu, fu, nr), dphi_0 = nl_dot(fu, Jg, nr);
            nl_phi_ctx k = {c, uprev, du, ytr, fu2};
            const double alpha = nl_backtracking(&k, phi_0, dphi_0);
            if (isnan(alpha)) { ret = ORC_FAILURE; break; }   /* InternalLineSearchFailed */
            for (size_t i = 0; i < nu; i++) y[i] = uprev[i] + alpha * du[i];
            orc_loss(c->P, c->T, c->p, c->N, c->mesh, y, c->Kd, c->Ki, fu);
        } else {
            /* dogleg step within Delta */
            const double nN = sqrt(nl_dot(du, du, nu));
            if (!(nN <= Delta)) {
                nl_jtvec(c, fu, g);
                for (size_t i = 0; i < nu; i++) g[i] = -g[i];   /* steepest descent direction */
                const double lg = sqrt(nl_dot(g, g, nu));
                nl_jvec(c, g, Jg);
                const double dc = lg * lg * lg / nl_dot(Jg, Jg, nr);
                if (dc >= Delta) {
                    for (size_t i = 0; i < nu; i++) du[i] = (Delta / lg) * g[i];
                } else {
                    for (size_t i = 0; i < nu; i++) g[i] *= dc / lg;           /* Cauchy point */
                    double aa = 0.0, bb = 0.0;
                    for (size_t i = 0; i < nu; i++) { const double df = du[i] - g[i]; aa += df * df; bb += df * g[i]; }
                    const double cc = dc * dc - Delta * Delta;
                    double disc = bb * bb - aa * cc;
                    if (disc < 0.0) disc = 0.0;
                    const double tau = (-bb + sqrt(disc)) / aa;
                    for (size_t i = 0; i < nu; i++) du[i] = g[i] + tau * (du[i] - g[i]);
                }
            }
            for (size_t i = 0; i < nu; i++) ytr[i] = uprev[i] + du[i];
            orc_loss(c->P, c->T, c->p, c->N, c->mesh, ytr, c->Kd, c->Ki, fu2);
            nl_jvec(c, du, Jg);
            nl_jtvec(c, fu, g);
            const double num = 0.5 * (nl_dot(fu, fu, nr) - nl_dot(fu2, fu2, nr));
            const double den = nl_dot(du, g, nu) + 0.5 * nl_dot(Jg, Jg, nr);
            const double rho = num / (-den);
            const int accept = rho > 1e-4;
            if (rho < 0.25) { Delta *= 0.25; shrink_counter++; }
            else { shrink_counter = 0; if (rho > 0.75) { Delta *= 2.0; if (Delta > Delta_max) Delta = Delta_max; } }
            if (accept) {
                memcpy(y, ytr, sizeof(double) * nu);
                memcpy(fu, fu2, sizeof(double) * nr);
                have_jac = 0;
            } else {
                /* rejected: stay, keep the Jacobian; the stages must again belong to y */
                orc_loss(c->P, c->T, c->p, c->N, c->mesh, y, c->Kd, c->Ki, fu);
            }
            if (shrink_counter > 32) { it++; ret = ORC_FAILURE; break; }   /* ShrinkThresholdExceeded */
        }
        it++;
        nrm = norm_inf(fu, nr);
        const int tc = nl_term_check(&term, nrm, y, uprev, nu);
        if (tc >= 0) { ret = tc; break; }
    }
    if (ret != ORC_SUCCESS && it > 0 && term.have_best) {
        /* SafeBest: hand back the best iterate seen and stages consistent with it */
        memcpy(y, ubest, sizeof(double) * nu);
        orc_loss(c->P, c->T, c->p, c->N, c->mesh, y, c->Kd, c->Ki, fu);
        nrm = norm_inf(fu, nr);
    }
    *resid_norm = nrm;
    *iters = it;
    free(fu); free(fu2); free(du); free(uprev); free(ytr); free(ubest); free(g); free(Jg); free(rhs_bc);
    return ret;
}

/* nlsolve: 0 the default polyalgorithm, 1 NewtonRaphson, 2 NewtonRaphson + BackTracking, 3 TrustRegion.
 * *iters counts the steps of every sub-solver that ran. */
int orc_nlsolve(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
                double *y, double *Kd, double *Ki, double abstol, int maxiters, int nlsolve, double *resid_norm,
                int *iters) {
    nl_ctx c;
    memset(&c, 0, sizeof(c));
    c.P = P; c.T = T; c.p = p; c.N = N; c.mesh = mesh; c.n = P->n; c.L = P->n_bc;
    c.La = P->problem_type == 1 ? P->n_bca : P->n_bc;
    c.nu = (size_t)N * c.n; c.nr = (size_t)c.L + (size_t)(N - 1) * c.n;
    c.Kd = Kd; c.Ki = Ki;
    c.Lb = (double *)malloc(sizeof(double) * (size_t)(N - 1) * c.n * c.n);
    c.Rb = (double *)malloc(sizeof(double) * (size_t)(N - 1) * c.n * c.n);
    c.B = (double *)malloc(sizeof(double) * ORC_MAX_BC_PTS * c.L * c.n);
    int ret, total = 0;
    if (nlsolve != 0) {
        ret = nl_run(&c, nlsolve, y, abstol, maxiters, resid_norm, &total);
    } else {
        double *y0 = (double *)malloc(sizeof(double) * c.nu), *ybest = (double *)malloc(sizeof(double) * c.nu);
        memcpy(y0, y, sizeof(double) * c.nu);
        double best = INFINITY;
        int best_ret = ORC_FAILURE;
        ret = ORC_FAILURE;
        for (int alg = 1; alg <= 3; alg++) {
            int it = 0;
            double nrm = 0.0;
            memcpy(y, y0, sizeof(double) * c.nu);
            const int r = nl_run(&c, alg, y, abstol, maxiters, &nrm, &it);
            total += it;
            if (r == ORC_SUCCESS) { ret = r; *resid_norm = nrm; best = -1.0; break; }
            if (alg == 1 || !(nrm >= best)) { best = nrm; best_ret = r; memcpy(ybest, y, sizeof(double) * c.nu); }
        }
        if (ret != ORC_SUCCESS) {
            /* all failed: the sub-solver that got closest provides the iterate and the return code */
            memcpy(y, ybest, sizeof(double) * c.nu);
            double *fu = (double *)malloc(sizeof(double) * c.nr);
            orc_loss(P, T, p, N, mesh, y, Kd, Ki, fu);
            *resid_norm = norm_inf(fu, c.nr);
            free(fu);
            ret = best_ret;
        }
        free(y0); free(ybest);
    }
    *iters = total;
    free(c.Lb); free(c.Rb); free(c.B);
    return ret;
}

/* the plain NewtonRaphson sub-solver alone (kept under its round-1 name) */
int orc_newton(const orc_problem *P, const orc_tableau *T, const double *p, int N, const double *mesh,
               double *y, double *Kd, double *Ki, double abstol, int maxiters, double *resid_norm,
               int *iters) {
    return orc_nlsolve(P, T, p, N, mesh, y, Kd, Ki, abstol, maxiters, 1, resid_norm, iters);
}

/* ------------------------------------------------------------------------------------------ */
/* defect estimate (MIRK/adaptivity.jl:370-415; Appendix A.5)                                  */
/* ------------------------------------------------------------------------------------------ */

double orc_defect(const orc_problem *P, const orc_tableau *T, const double *p, int N,
                  const double *mesh, const double *y, const double *Kd, double *Ki, double *errors) {
    const int n = P->n, s = T->s, si = T->s_star - T->s;
    double w1[ORC_MAX_SS], w1p[ORC_MAX_SS], w2[ORC_MAX_SS], w2p[ORC_MAX_SS];
    orc_interp_weights(T->order, T->tau_star, w1, w1p);
    orc_interp_weights(T->order, 1.0 - T->tau_star, w2, w2p);
    orc_interp_setup(P, T, p, N, mesh, y, Kd, Ki);
    double *z = (double *)malloc(sizeof(double) * n), *zp = (double *)malloc(sizeof(double) * n);
    double *d1 = (double *)malloc(sizeof(double) * n), *d2 = (double *)malloc(sizeof(double) * n);
    double defect = 0.0;
    for (int i = 0; i < N - 1; i++) {
        const double dt = mesh[i + 1] - mesh[i];
        const double *K = Kd + (size_t)i * s * n, *KI = Ki + (size_t)i * si * n;
        double est[2];
        for (int smp = 0; smp < 2; smp++) {
            const double *w = smp ? w2 : w1, *wp = smp ? w2p : w1p;
            double *d = smp ? d2 : d1;
            for (int k = 0; k < n; k++) { z[k] = 0.0; zp[k] = 0.0; }
            for (int r = 0; r < s; r++) for (int k = 0; k < n; k++) z[k] += K[r * n + k] * w[r];
            for (int r = 0; r < si; r++) for (int k = 0; k < n; k++) z[k] += KI[r * n + k] * w[s + r];
            for (int r = 0; r < s; r++) for (int k = 0; k < n; k++) zp[k] += K[r * n + k] * wp[r];
            for (int r = 0; r < si; r++) for (int k = 0; k < n; k++) zp[k] += KI[r * n + k] * wp[s + r];
            for (int k = 0; k < n; k++) z[k] = z[k] * dt + y[(size_t)i * n + k];
            const double tau = smp ? (1.0 - T->tau_star) : T->tau_star;
            P->f(d, z, p, mesh[i] + tau * dt, P->ctx);
            double e = 0.0;
            for (int k = 0; k < n; k++) {
                d[k] = (zp[k] - d[k]) / (fabs(d[k]) + 1.0);
                if (fabs(d[k]) > e) e = fabs(d[k]);
            }
            est[smp] = e;
        }
        const double *pick = est[0] > est[1] ? d1 : d2;
        for (int k = 0; k < n; k++) {
            errors[(size_t)i * n + k] = pick[k];
            if (fabs(pick[k]) > defect) defect = fabs(pick[k]);
            if (isnan(pick[k])) defect = NAN;
        }
    }
    free(z); free(zp); free(d1); free(d2);
    return defect;
}

/* ------------------------------------------------------------------------------------------ */
/* mesh selection (MIRK/adaptivity.jl:23-75, redistribute! :250-278, half_mesh! :287-304;       */
/* Appendix A.6).  Returns ORC_SUCCESS / ORC_FAILURE; mesh_new must hold 4*(N-1)+1 doubles.     */
/* ------------------------------------------------------------------------------------------ */

int orc_mesh_select(int order, int n, int N, const double *mesh, const double *errors, double abstol,
                    int max_num_subintervals, int *N_new, double *mesh_new) {
    return orc_mesh_select_ex(order, n, N, mesh, errors, NULL, abstol, max_num_subintervals,
                              orc_convergence_order(order) + 1, 1.0, N_new, mesh_new);
}

/* the four mesh selectors of MIRK/adaptivity.jl:23-243 differ in three things only: the exponent 1 / expo_den of
 * s_hat (order + 1 for DefectControl / Sequential / Hybrid, order for GlobalErrorControl), the halving threshold rho
 * (1 for DefectControl, 2 for the others) and, for HybridErrorControl, s_hat being the SUM of the powers of two error
 * arrays (defect and global error, errors2 != NULL). */
int orc_mesh_select_ex(int order, int n, int N, const double *mesh, const double *errors, const double *errors2,
                       double abstol, int max_num_subintervals, int expo_den, double rho, int *N_new, double *mesh_new) {
    (void)order;
    const int ni = N - 1;
    double *sh = (double *)malloc(sizeof(double) * ni);
    double r1 = 0.0, r2 = 0.0;
    for (int i = 0; i < ni; i++) {
        double e = 0.0;
        for (int k = 0; k < n; k++) if (fabs(errors[(size_t)i * n + k]) > e) e = fabs(errors[(size_t)i * n + k]);
        sh[i] = pow(e / abstol, 1.0 / expo_den);
        if (errors2) {
            double e2 = 0.0;
            for (int k = 0; k < n; k++) if (fabs(errors2[(size_t)i * n + k]) > e2) e2 = fabs(errors2[(size_t)i * n + k]);
            sh[i] = sh[i] + pow(e2 / abstol, 1.0 / expo_den);
        }
        if (sh[i] > r1) r1 = sh[i];
        r2 += sh[i];
    }
    const double r3 = r2 / ni;
    long n_predict = (long)nearbyint(1.3 * r2 + 1.0); /* round(Int, .) is half-to-even */
    const double n_ = 0.1 * ni;
    if (fabs((double)(n_predict - ni)) < n_) n_predict = (long)nearbyint(ni + n_);
    int info = ORC_SUCCESS;
    if (r1 <= rho * r3) {
        const int ns = 2 * ni;
        if (ns > max_num_subintervals) {
            info = ORC_FAILURE;
            *N_new = N;
            memcpy(mesh_new, mesh, sizeof(double) * N);
        } else {
            *N_new = ns + 1;
            for (int i = 0; i < ni; i++) {
                mesh_new[2 * i] = mesh[i];
                mesh_new[2 * i + 1] = (mesh[i + 1] + mesh[i]) / 2.0;
            }
            mesh_new[2 * ni] = mesh[ni];
        }
    } else {
        long lb = N / 2, ub = 4L * ni;
        long ns = n_predict < lb ? lb : (n_predict > ub ? ub : n_predict);
        if (ns > max_num_subintervals) {
            info = ORC_FAILURE;
            *N_new = N;
            memcpy(mesh_new, mesh, sizeof(double) * N);
        } else {
            for (int i = 0; i < ni; i++) sh[i] /= (mesh[i + 1] - mesh[i]);
            double tot = 0.0;
            for (int i = 0; i < ni; i++) tot += sh[i] * (mesh[i + 1] - mesh[i]);
            const double zeta = tot / (double)ns;
            /* resize!(cache.mesh, ns+1) keeps the leading old entries; new tail entries are
             * uninitialised in Julia, here they default to t_end */
            for (long i = 0; i <= ns; i++) mesh_new[i] = (i < N) ? mesh[i] : mesh[ni];
            int k = 0;
            long i = 0;
            mesh_new[0] = mesh[0];
            double t = mesh[0], integral = 0.0;
            while (k < ni) {
                const double next_piece = sh[k] * (mesh[k + 1] - t);
                const double int_next = integral + next_piece;
                if (int_next > zeta) {
                    const double tn = (zeta - integral) / sh[k] + t;
                    if (i + 1 <= ns) mesh_new[i + 1] = tn;
                    t = tn;
                    i++;
                    integral = 0.0;
                } else {
                    integral = int_next;
                    t = mesh[k + 1];
                    k++;
                }
            }
            mesh_new[ns] = mesh[ni];
            *N_new = (int)ns + 1;
        }
    }
    free(sh);
    return info;
}

/* new guess on the new mesh (MIRK/mirk.jl:364-372, adaptivity.jl:6-13,590-621; Appendix A.7).
 * inplace_quirk=1 reproduces Q3: sum_stages! adds `cache.y0.u[i_old]`, the array being rewritten. */
void orc_reinterp(const orc_problem *P, const orc_tableau *T, int N_old, const double *mesh_old,
                  const double *y_old, const double *Kd, const double *Ki, int N_new,
                  const double *mesh_new, double *y_new, int inplace_quirk) {
    const int n = P->n, s = T->s, si = T->s_star - T->s;
    const int NB = N_old > N_new ? N_old : N_new;
    double *buf = (double *)calloc((size_t)NB * n, sizeof(double));
    memcpy(buf, y_old, sizeof(double) * (size_t)N_old * n);
    double w[ORC_MAX_SS], wp[ORC_MAX_SS];
    for (int j = 0; j < N_new; j++) {
        const double t = mesh_new[j];
        const int i = orc_interval(mesh_old, N_old, t);
        const double dt = mesh_old[i + 1] - mesh_old[i];
        const double tau = (t - mesh_old[i]) / dt;
        orc_interp_weights(T->order, tau, w, wp);
        const double *K = Kd + (size_t)i * s * n, *KI = Ki + (size_t)i * si * n;
        const double *base = inplace_quirk ? (buf + (size_t)i * n) : (y_old + (size_t)i * n);
        for (int k = 0; k < n; k++) {
            double acc = 0.0;
            for (int r = 0; r < s; r++) acc += K[r * n + k] * w[r];
            for (int r = 0; r < si; r++) acc += KI[r * n + k] * w[s + r];
            y_new[(size_t)j * n + k] = acc * dt + base[k];
        }
        if (inplace_quirk) memcpy(buf + (size_t)j * n, y_new + (size_t)j * n, sizeof(double) * n);
    }
    free(buf);
}

/* ------------------------------------------------------------------------------------------ */
/* adaptive outer loop (MIRK/mirk.jl:286-388; Appendix A.8)                                     */
/* ------------------------------------------------------------------------------------------ */

/* ---- global-error estimates (MIRK/adaptivity.jl:464-567) -------------------------------------------------------
 * HO  (HOErrorControl):  the same mesh solved again with the method of order + 2 (MIRK2 -> 4, 3 -> 5, 4 -> 6), current
 *     solution as the guess, non-adaptive;  err = (y_high - y) / (1 + |y|) per node.
 * RE  (REErrorControl):  the mesh halved (halve_sol: nodes copied, midpoints averaged), same method, non-adaptive;
 *     err = (y_half[1:2:end] - y) / (1 + |y|), norm scaled by 2^p / (2^p - 1).
 * GE_subinterval!: interval i keeps the error vector of node i if its max-norm is >= that of node i + 1, else that of
 * node i + 1.  Returns the error norm, < 0 when the higher-order method does not exist. */
double orc_global_error(const orc_problem *P, int order, const double *p, int N, const double *mesh, const double *y,
                        const orc_options *opt, int method, double *errors) {
    const int n = P->n, pconv = orc_convergence_order(order);
    orc_tableau Th;
    int Nh = N;
    double *mh = NULL, *yh = NULL;
    if (method == 0) {
        if (order == ORC_MIRK6I || pconv + 2 > 6 || orc_tableau_get(pconv + 2, &Th)) return -1.0;
        mh = (double *)malloc(sizeof(double) * N);
        yh = (double *)malloc(sizeof(double) * (size_t)N * n);
        memcpy(mh, mesh, sizeof(double) * N);
        memcpy(yh, y, sizeof(double) * (size_t)N * n);
    } else {
        orc_tableau_get(order, &Th);
        Nh = 2 * (N - 1) + 1;
        mh = (double *)malloc(sizeof(double) * Nh);
        yh = (double *)malloc(sizeof(double) * (size_t)Nh * n);
        for (int i = 0; i < N; i++) {
            mh[2 * i] = mesh[i];
            memcpy(yh + (size_t)2 * i * n, y + (size_t)i * n, sizeof(double) * n);
        }
        for (int i = 0; i < N - 1; i++) {
            mh[2 * i + 1] = (mh[2 * i + 2] + mh[2 * i]) / 2.0;
            for (int k = 0; k < n; k++) yh[(size_t)(2 * i + 1) * n + k] = (yh[(size_t)(2 * i + 2) * n + k] + yh[(size_t)2 * i * n + k]) / 2.0;
        }
    }
    const int sh = Th.s, sih = Th.s_star - Th.s;
    double *Kh = (double *)calloc((size_t)(Nh - 1) * sh * n, sizeof(double));
    double *Kih = (double *)calloc((size_t)(Nh - 1) * (sih > 0 ? sih : 1) * n, sizeof(double));
    double rn = 0.0;
    int its = 0;
    orc_nlsolve(P, &Th, p, Nh, mh, yh, Kh, Kih, opt->abstol, opt->maxiters, opt->nlsolve, &rn, &its);
    const int stride = method == 0 ? 1 : 2;
    double *err = (double *)malloc(sizeof(double) * (size_t)N * n), *nm = (double *)malloc(sizeof(double) * N);
    for (int i = 0; i < N; i++) {
        double m = 0.0;
        for (int k = 0; k < n; k++) {
            const double lo = y[(size_t)i * n + k];
            const double e = (yh[(size_t)i * stride * n + k] - lo) / (1.0 + fabs(lo));
            err[(size_t)i * n + k] = e;
            if (!(fabs(e) <= m)) m = fabs(e);
        }
        nm[i] = m;
    }
    double norm = 0.0;
    for (int i = 0; i < N - 1; i++) {
        const int pick = nm[i] >= nm[i + 1] ? i : i + 1;
        memcpy(errors + (size_t)i * n, err + (size_t)pick * n, sizeof(double) * n);
        if (!(nm[pick] <= norm)) norm = nm[pick];
    }
    free(mh); free(yh); free(Kh); free(Kih); free(err); free(nm);
    if (method == 1) norm = norm * pow(2.0, pconv) / (pow(2.0, pconv) - 1.0);
    return norm;
}

int orc_solve(const orc_problem *P, int order, const double *p, int N0, const double *mesh0,
              const double *y0, const orc_options *opt, orc_result *out) {
    orc_tableau T;
    if (orc_tableau_get(order, &T)) return -1;
    const int n = P->n, s = T.s, si = T.s_star - T.s;
    int N = N0;
    double *mesh = (double *)malloc(sizeof(double) * N);
    double *y = (double *)malloc(sizeof(double) * (size_t)N * n);
    memcpy(mesh, mesh0, sizeof(double) * N);
    memcpy(y, y0, sizeof(double) * (size_t)N * n);
    double *Kd = (double *)calloc((size_t)(N - 1) * s * n, sizeof(double));
    double *Ki = (double *)calloc((size_t)(N - 1) * si * n, sizeof(double));
    memset(out, 0, sizeof(*out));
    int info = ORC_SUCCESS;
    double error_norm = 2.0 * opt->abstol, resid_norm = 0.0;
    do {
        int iters = 0;
        const int nret = orc_nlsolve(P, &T, p, N, mesh, y, Kd, Ki, opt->abstol, opt->maxiters, opt->nlsolve,
                                    &resid_norm, &iters);
        out->newton_iters += iters;
        error_norm = 2.0 * opt->abstol;
        info = nret;
        const int h = out->n_hist < 64 ? out->n_hist : 63;
        out->hist_N[h] = N;
        out->hist_newton[h] = iters;
        out->hist_defect[h] = NAN;
        out->outer_iters++;
        if (out->n_hist < 64) out->n_hist++;
        if (!opt->adaptive) break;
        if (info == ORC_SUCCESS) {
            /* error_estimate!(cache, controller, ...) — MIRK/adaptivity.jl:355-369 (dispatch), :370-461 (defect), :464-567 */
            double *errors = (double *)malloc(sizeof(double) * 2 * (size_t)(N - 1) * n);
            double *errors2 = NULL;
            const int pconv = orc_convergence_order(order);
            int expo_den = pconv + 1;
            double rho = 1.0;
            if (opt->controller == 0) {          /* DefectControl */
                error_norm = orc_defect(P, &T, p, N, mesh, y, Kd, Ki, errors);
                if (!(error_norm <= opt->defect_threshold)) info = ORC_FAILURE;
            } else if (opt->controller == 1) {   /* GlobalErrorControl: no threshold test */
                if (P->problem_type == 0) orc_interp_setup(P, &T, p, N, mesh, y, Kd, Ki); /* filled by every loss call (Q7) */
                error_norm = orc_global_error(P, order, p, N, mesh, y, opt, opt->ge_method, errors);
                expo_den = pconv;
                rho = 2.0;
                if (error_norm < 0.0) { info = ORC_FAILURE; error_norm = 2.0 * opt->abstol; }
            } else if (opt->controller == 2) {   /* SequentialErrorControl: defect first, global error once it passes */
                error_norm = orc_defect(P, &T, p, N, mesh, y, Kd, Ki, errors);
                if (!(error_norm <= opt->defect_threshold)) info = ORC_FAILURE;
                if (error_norm <= opt->abstol) {
                    const double ge = orc_global_error(P, order, p, N, mesh, y, opt, opt->ge_method, errors);
                    if (ge < 0.0) info = ORC_FAILURE; else { error_norm = ge; info = ORC_SUCCESS; }
                }
                rho = 2.0;
            } else {                             /* HybridErrorControl: DE * defect + GE * global error, always Success */
                errors2 = errors + (size_t)(N - 1) * n;
                const double dn = orc_defect(P, &T, p, N, mesh, y, Kd, Ki, errors);
                const double ge = orc_global_error(P, order, p, N, mesh, y, opt, opt->ge_method, errors2);
                if (ge < 0.0) { info = ORC_FAILURE; error_norm = 2.0 * opt->abstol; }
                else error_norm = opt->DE * dn + opt->GE * ge;
                rho = 2.0;
            }
            out->hist_defect[h] = error_norm;
            if (info == ORC_SUCCESS && error_norm > opt->abstol) {
                int Nn = 0;
                double *mesh_new = (double *)malloc(sizeof(double) * (4 * (size_t)(N - 1) + 1));
                info = orc_mesh_select_ex(order, n, N, mesh, errors, errors2, opt->abstol,
                                          opt->max_num_subintervals, expo_den, rho, &Nn, mesh_new);
                if (info == ORC_SUCCESS) {
                    double *y_new = (double *)malloc(sizeof(double) * (size_t)Nn * n);
                    orc_reinterp(P, &T, N, mesh, y, Kd, Ki, Nn, mesh_new, y_new, opt->reinterp_inplace);
                    free(y); free(mesh); free(Kd); free(Ki);
                    y = y_new; mesh = mesh_new; N = Nn;
                    Kd = (double *)calloc((size_t)(N - 1) * s * n, sizeof(double));
                    Ki = (double *)calloc((size_t)(N - 1) * si * n, sizeof(double));
                } else {
                    free(mesh_new);
                }
                free(errors);
                continue;
            }
            free(errors);
        }
        if (info != ORC_SUCCESS) {
            /* MIRK/mirk.jl:374-385: halve the mesh, ZERO the guess, force a restart (quirk Q4) */
            if (2 * (N - 1) > opt->max_num_subintervals) {
                info = ORC_FAILURE;
            } else {
                const int Nn = 2 * (N - 1) + 1;
                double *mesh_new = (double *)malloc(sizeof(double) * Nn);
                for (int i = 0; i < N - 1; i++) {
                    mesh_new[2 * i] = mesh[i];
                    mesh_new[2 * i + 1] = (mesh[i + 1] + mesh[i]) / 2.0;
                }
                mesh_new[Nn - 1] = mesh[N - 1];
                free(mesh); free(y); free(Kd); free(Ki);
                mesh = mesh_new; N = Nn;
                y = (double *)calloc((size_t)N * n, sizeof(double));
                Kd = (double *)calloc((size_t)(N - 1) * s * n, sizeof(double));
                Ki = (double *)calloc((size_t)(N - 1) * si * n, sizeof(double));
                info = ORC_SUCCESS;
            }
        }
    } while (info == ORC_SUCCESS && error_norm > opt->abstol && out->outer_iters < opt->max_outer);
    if (info == ORC_SUCCESS && opt->adaptive && error_norm > opt->abstol) info = ORC_MAXITERS;
    out->N = N; out->mesh = mesh; out->y = y; out->Kd = Kd; out->Ki = Ki;
    out->retcode = info; out->resid_norm = resid_norm; out->defect_norm = error_norm;
    return 0;
}

void orc_result_free(orc_result *r) {
    free(r->mesh); free(r->y); free(r->Kd); free(r->Ki);
    r->mesh = r->y = r->Kd = r->Ki = NULL;
}

/* SciMLBase EnsembleProblem driver restated for the packed-parameter case
 * (MIRK/test/Core/ensemble_tests.jl:20-38): prob_func only swaps p. */
int orc_ensemble_solve(const orc_problem *P, int order, int ntraj, const double *p, const double *u0,
                       double t0, double t1, int nint, const orc_options *opt, int nthreads,
                       int *retcodes, int *N_final, double *y_first, int *newton_iters) {
    const int n = P->n;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int tr = 0; tr < ntraj; tr++) {
        double *mesh = (double *)malloc(sizeof(double) * (nint + 1));
        double *y = (double *)malloc(sizeof(double) * (size_t)(nint + 1) * n);
        orc_mesh_uniform(t0, t1, nint, mesh);
        for (int i = 0; i <= nint; i++) memcpy(y + (size_t)i * n, u0, sizeof(double) * n);
        orc_result R;
        orc_solve(P, order, p + (size_t)tr * P->n_p, nint + 1, mesh, y, opt, &R);
        retcodes[tr] = R.retcode;
        if (N_final) N_final[tr] = R.N;
        if (y_first) memcpy(y_first + (size_t)tr * n, R.y, sizeof(double) * n);
        if (newton_iters) newton_iters[tr] = R.newton_iters;
        orc_result_free(&R);
        free(mesh); free(y);
    }
    return 0;
}
