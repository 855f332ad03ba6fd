// tableau.cuh — MIRK2 / MIRK3 / MIRK4 / MIRK5 / MIRK6 coefficient tables and continuous-extension weights as
// compile-time constants, so every stage loop in the kernels unrolls to straight-line FP64.
//
// Restates (not copied; the reference stores them as runtime Julia arrays built from rationals):
//   lib/BoundaryValueDiffEqMIRK/src/mirk_tableaus.jl:13-34   (MIRK2: s=1, s*=3, tau*=0.25)
//   lib/BoundaryValueDiffEqMIRK/src/mirk_tableaus.jl:36-60   (MIRK3: s=2, s*=3, tau*=0.25)
//   lib/BoundaryValueDiffEqMIRK/src/mirk_tableaus.jl:62-87   (MIRK4: s=3, s*=4, tau*=0.226)
//   lib/BoundaryValueDiffEqMIRK/src/mirk_tableaus.jl:89-118  (MIRK5: s=4, s*=6, tau*=0.3)
//   lib/BoundaryValueDiffEqMIRK/src/mirk_tableaus.jl:120-152 (MIRK6: s=5, s*=9, tau*=0.7156)
//   lib/BoundaryValueDiffEqMIRK/src/interpolation.jl:463-575 (weights w, w' per order)
//   lib/BoundaryValueDiffEqMIRK/src/mirk_tableaus.jl:154-194 + interpolation.jl:582-710 (MIRK6I, code 7)
#pragma once
#include <cuda_runtime.h>

namespace mirk {

#define MIRK_HD __host__ __device__ __forceinline__

template <int ORDER> struct Tableau;

// MIRK2: the implicit midpoint rule; the interpolant is built on f(y_i), f(y_{i+1})
template <> struct Tableau<2> {
    static constexpr int order = 2, s = 1, s_star = 3, si = 2;
    MIRK_HD static constexpr double c(int) { return 0.5; }
    MIRK_HD static constexpr double v(int) { return 0.5; }
    MIRK_HD static constexpr double b(int) { return 1.0; }
    MIRK_HD static constexpr double x(int, int) { return 0.0; }
    MIRK_HD static constexpr double c_star(int r) { return r == 0 ? 0.0 : 1.0; }
    MIRK_HD static constexpr double v_star(int r) { return r == 0 ? 0.0 : 1.0; }
    MIRK_HD static constexpr double x_star(int, int) { return 0.0; }
    MIRK_HD static constexpr double tau_star() { return 0.25; }
    MIRK_HD static void weights(double t, double* w, double* wp) {
        w[0] = 0.0; w[1] = t * (1.0 - t / 2.0); w[2] = t * t / 2.0;
        wp[0] = 0.0; wp[1] = 1.0 - t; wp[2] = t;
    }
};

template <> struct Tableau<3> {
    static constexpr int order = 3, s = 2, s_star = 3, si = 1;
    MIRK_HD static constexpr double c(int r) { return r == 0 ? 0.0 : 2.0 / 3.0; }
    MIRK_HD static constexpr double v(int r) { return r == 0 ? 0.0 : 4.0 / 9.0; }
    MIRK_HD static constexpr double b(int r) { return r == 0 ? 1.0 / 4.0 : 3.0 / 4.0; }
    MIRK_HD static constexpr double x(int r, int j) { return (r == 1 && j == 0) ? 2.0 / 9.0 : 0.0; }
    MIRK_HD static constexpr double c_star(int) { return 1.0; }
    MIRK_HD static constexpr double v_star(int) { return 1.0; }
    MIRK_HD static constexpr double x_star(int, int) { return 0.0; }
    MIRK_HD static constexpr double tau_star() { return 0.25; }
    MIRK_HD static void weights(double t, double* w, double* wp) {
        w[0] = t / 4.0 * (2.0 * t * t - 5.0 * t + 4.0);
        w[1] = -3.0 / 4.0 * t * t * (2.0 * t - 3.0);
        w[2] = t * t * (t - 1.0);
        wp[0] = 3.0 / 2.0 * (t - 2.0 / 3.0) * (t - 1.0);
        wp[1] = -9.0 / 2.0 * t * (t - 1.0);
        wp[2] = 3.0 * t * (t - 2.0 / 3.0);
    }
};

template <> struct Tableau<5> {
    static constexpr int order = 5, s = 4, s_star = 6, si = 2;
    MIRK_HD static constexpr double c(int r) { return r == 0 ? 0.0 : r == 1 ? 1.0 : r == 2 ? 3.0 / 4.0 : 3.0 / 10.0; }
    MIRK_HD static constexpr double v(int r) {
        return r == 0 ? 0.0 : r == 1 ? 1.0 : r == 2 ? 27.0 / 32.0 : 837.0 / 1250.0;
    }
    MIRK_HD static constexpr double b(int r) {
        return r == 0 ? 5.0 / 54.0 : r == 1 ? 1.0 / 14.0 : r == 2 ? 32.0 / 81.0 : 250.0 / 567.0;
    }
    MIRK_HD static constexpr double x(int r, int j) {
        return r == 2 ? (j == 0 ? 3.0 / 64.0 : j == 1 ? -9.0 / 64.0 : 0.0)
             : r == 3 ? (j == 0 ? 21.0 / 1000.0 : j == 1 ? 63.0 / 5000.0 : j == 2 ? -252.0 / 625.0 : 0.0)
                      : 0.0;
    }
    MIRK_HD static constexpr double c_star(int r) { return r == 0 ? 4.0 / 5.0 : 13.0 / 23.0; }
    MIRK_HD static constexpr double v_star(int r) { return c_star(r); }
    MIRK_HD static constexpr double x_star(int r, int j) {
        return r == 0 ? (j == 0 ? 14.0 / 1125.0 : j == 1 ? -74.0 / 875.0 : j == 2 ? -128.0 / 3375.0
                         : j == 3 ? 104.0 / 945.0 : 0.0)
                      : (j == 0 ? 1.0 / 2.0 : j == 1 ? 4508233.0 / 1958887.0 : j == 2 ? 48720832.0 / 2518569.0
                         : j == 3 ? -27646420.0 / 17629983.0 : j == 4 ? -11517095.0 / 559682.0 : 0.0);
    }
    MIRK_HD static constexpr double tau_star() { return 0.3; }
    MIRK_HD static void weights(double t, double* w, double* wp) {
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
    }
};

template <> struct Tableau<4> {
    static constexpr int order = 4, s = 3, s_star = 4, si = 1;
    MIRK_HD static constexpr double c(int r) { return r == 0 ? 0.0 : r == 1 ? 1.0 : 0.5; }
    MIRK_HD static constexpr double v(int r) { return r == 0 ? 0.0 : r == 1 ? 1.0 : 0.5; }
    MIRK_HD static constexpr double b(int r) { return r == 2 ? 2.0 / 3.0 : 1.0 / 6.0; }
    MIRK_HD static constexpr double x(int r, int j) {
        return (r == 2 && j == 0) ? 1.0 / 8.0 : (r == 2 && j == 1) ? -1.0 / 8.0 : 0.0;
    }
    MIRK_HD static constexpr double c_star(int) { return 3.0 / 4.0; }
    MIRK_HD static constexpr double v_star(int) { return 27.0 / 32.0; }
    MIRK_HD static constexpr double x_star(int, int j) {
        return j == 0 ? 3.0 / 64.0 : j == 1 ? -9.0 / 64.0 : 0.0;
    }
    MIRK_HD static constexpr double tau_star() { return 0.226; }
    // w[0..s*), wp[0..s*)
    MIRK_HD static void weights(double t, double* w, double* wp) {
        const double t2 = t * t, tm1 = t - 1.0, t4m3 = t * 4.0 - 3.0, t2m1 = t * 2.0 - 1.0;
        w[0] = -t * (2.0 * t - 3.0) * (2.0 * t2 - 3.0 * t + 2.0) / 6.0;
        w[1] = t2 * (12.0 * t2 - 20.0 * t + 9.0) / 6.0;
        w[2] = 2.0 * t2 * (6.0 * t2 - 14.0 * t + 9.0) / 3.0;
        w[3] = -16.0 * t2 * tm1 * tm1 / 3.0;
        wp[0] = -tm1 * t4m3 * t2m1 / 3.0;
        wp[1] = t * t2m1 * t4m3;
        wp[2] = 4.0 * t * t4m3 * tm1;
        wp[3] = -32.0 * t * t2m1 * tm1 / 3.0;
    }
};

template <> struct Tableau<6> {
    static constexpr int order = 6, s = 5, s_star = 9, si = 4;
    MIRK_HD static constexpr double c(int r) {
        return r == 0 ? 0.0 : r == 1 ? 1.0 : r == 2 ? 0.25 : r == 3 ? 0.75 : 0.5;
    }
    MIRK_HD static constexpr double v(int r) {
        return r == 0 ? 0.0 : r == 1 ? 1.0 : r == 2 ? 5.0 / 32.0 : r == 3 ? 27.0 / 32.0 : 0.5;
    }
    MIRK_HD static constexpr double b(int r) {
        return r < 2 ? 7.0 / 90.0 : r < 4 ? 16.0 / 45.0 : 2.0 / 15.0;
    }
    MIRK_HD static constexpr double x(int r, int j) {
        return r == 2 ? (j == 0 ? 9.0 / 64.0 : j == 1 ? -3.0 / 64.0 : 0.0)
             : r == 3 ? (j == 0 ? 3.0 / 64.0 : j == 1 ? -9.0 / 64.0 : 0.0)
             : r == 4 ? (j == 0 ? -5.0 / 24.0 : j == 1 ? 5.0 / 24.0 : j == 2 ? 2.0 / 3.0
                                                       : j == 3 ? -2.0 / 3.0 : 0.0)
                      : 0.0;
    }
    MIRK_HD static constexpr double c_star(int r) {
        return r == 0 ? 7.0 / 16.0 : r == 1 ? 3.0 / 8.0 : r == 2 ? 9.0 / 16.0 : 1.0 / 8.0;
    }
    MIRK_HD static constexpr double v_star(int r) { return c_star(r); }
    MIRK_HD static constexpr double x_star(int r, int j) {
        return r == 0 ? (j == 0 ? 1547.0 / 32768.0 : j == 1 ? -1225.0 / 32768.0 : j == 2 ? 749.0 / 4096.0
                         : j == 3 ? -287.0 / 2048.0 : j == 4 ? -861.0 / 16384.0 : 0.0)
             : r == 1 ? (j == 0 ? 83.0 / 1536.0 : j == 1 ? -13.0 / 384.0 : j == 2 ? 283.0 / 1536.0
                         : j == 3 ? -167.0 / 1536.0 : j == 4 ? -49.0 / 512.0 : 0.0)
             : r == 2 ? (j == 0 ? 1225.0 / 32768.0 : j == 1 ? -1547.0 / 32768.0 : j == 2 ? 287.0 / 2048.0
                         : j == 3 ? -749.0 / 4096.0 : j == 4 ? 861.0 / 16384.0 : 0.0)
                      : (j == 0 ? 233.0 / 3456.0 : j == 1 ? -19.0 / 1152.0 : j == 5 ? -5.0 / 72.0
                         : j == 6 ? 7.0 / 72.0 : j == 7 ? -17.0 / 216.0 : 0.0);
    }
    MIRK_HD static constexpr double tau_star() { return 0.7156; }
    MIRK_HD static void weights(double t, double* w, double* wp) {
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
        w[8] = 16384.0 / 441.0 * t3 - 16384.0 / 147.0 * t4 + 16384.0 / 147.0 * t5 -
               16384.0 / 441.0 * t6;
        wp[0] = 1.0 - 28607.0 / 3717.0 * t - 166210.0 / 11151.0 * t2 + 1339120.0 / 11151.0 * t3 -
                1911296.0 / 11151.0 * t4 + 813056.0 / 11151.0 * t5;
        wp[1] = 777.0 / 295.0 * t - 2534158.0 / 78057.0 * t2 + 8354320.0 / 78057.0 * t3 -
                10479104.0 / 78057.0 * t4 + 22657024.0 / 390285.0 * t5;
        wp[2] = -2016.0 / 59.0 * t + 222176.0 / 531.0 * t2 - 720128.0 / 531.0 * t3 +
                876544.0 / 531.0 * t4 - 360448.0 / 531.0 * t5;
        wp[3] = wp[2];
        wp[4] = -756.0 / 59.0 * t + 27772.0 / 177.0 * t2 - 90016.0 / 177.0 * t3 +
                109568.0 / 177.0 * t4 - 45056.0 / 177.0 * t5;
        wp[5] = -190464.0 / 413.0 * t + 62384128.0 / 11151.0 * t2 - 197718016.0 / 11151.0 * t3 +
                233799680.0 / 11151.0 * t4 - 93323264.0 / 11151.0 * t5;
        wp[6] = 1792.0 / 5.0 * t - 4352.0 * t2 + 13824.0 * t3 - 16384.0 * t4 + 32768.0 / 5.0 * t5;
        wp[7] = 100352.0 / 531.0 * t - 179554304.0 / 78057.0 * t2 + 573452288.0 / 78057.0 * t3 -
                683376640.0 / 78057.0 * t4 + 274726912.0 / 78057.0 * t5;
        wp[8] = 16384.0 / 147.0 * t2 - 65536.0 / 147.0 * t3 + 81920.0 / 147.0 * t4 -
                32768.0 / 147.0 * t5;
    }
};

// MIRK6I (`order` code 7): the 6th-order tableau with irrational abscissae, s = 5, s* = 8, tau* = 0.4
//   lib/BoundaryValueDiffEqMIRK/src/mirk_tableaus.jl:154-194 (the second assignments of `b` and `x_star` are the
//   ones in force), weights lib/BoundaryValueDiffEqMIRK/src/interpolation.jl:582-710.
// The expressions keep the reference's evaluation order (several entries are differences of numbers 60x their
// size), with sqrt(21), sqrt(7), sqrt(3) as correctly rounded literals.
constexpr int kMIRK6I = 7;
template <> struct Tableau<kMIRK6I> {
    static constexpr int order = 6, s = 5, s_star = 8, si = 3;
    static constexpr double s21 = 4.58257569495584000658804719372800848898, s7 = 2.64575131106459059050161575363926042571,
                            s3 = 1.73205080756887729352744634150587236694;
    MIRK_HD static constexpr double c(int r) {
        return r == 0 ? 0.0 : r == 1 ? 1.0 : r == 2 ? 0.5 - s21 / 14.0 : r == 3 ? 0.5 + s21 / 14.0 : 0.5;
    }
    MIRK_HD static constexpr double v(int r) {
        return r == 0 ? 0.0 : r == 1 ? 1.0 : r == 2 ? 0.5 - 9.0 * s21 / 98.0 : r == 3 ? 0.5 + 9.0 * s21 / 98.0 : 0.5;
    }
    MIRK_HD static constexpr double b(int r) { return r < 2 ? 1.0 / 20.0 : r < 4 ? 49.0 / 180.0 : 16.0 / 45.0; }
    MIRK_HD static constexpr double x(int r, int j) {
        return r == 2 ? (j == 0 ? 1.0 / 14.0 + s21 / 98.0 : j == 1 ? -1.0 / 14.0 + s21 / 98.0 : 0.0)
             : r == 3 ? (j == 0 ? 1.0 / 14.0 - s21 / 98.0 : j == 1 ? -1.0 / 14.0 - s21 / 98.0 : 0.0)
             : r == 4 ? (j == 0 ? -5.0 / 128.0 : j == 1 ? 5.0 / 128.0 : j == 2 ? 7.0 * s21 / 128.0
                                                         : j == 3 ? -7.0 * s21 / 128.0 : 0.0)
                      : 0.0;
    }
    MIRK_HD static constexpr double c_star(int r) { return r == 0 ? 0.5 : r == 1 ? 0.5 - s7 / 14.0 : 87.0 / 100.0; }
    MIRK_HD static constexpr double v_star(int r) { return c_star(r); }
    MIRK_HD static constexpr double x_star(int r, int j) {
        return r == 0 ? (j == 0 ? 1.0 / 64.0 : j == 1 ? -1.0 / 64.0 : j == 2 ? 7.0 / 192.0 * s21
                         : j == 3 ? -7.0 / 192.0 * s21 : 0.0)
             : r == 1 ? (j == 0 ? 3.0 / 112.0 + 9.0 / 1960.0 * s7 : j == 1 ? -3.0 / 112.0 + 9.0 / 1960.0 * s7
                         : j == 2 ? 11.0 / 840.0 * s7 + 3.0 / 112.0 * s7 * s3
                         : j == 3 ? 11.0 / 840.0 * s7 - 3.0 / 112.0 * s7 * s3
                         : j == 4 ? 88.0 / 5145.0 * s7 : j == 5 ? -18.0 / 343.0 * s7 : 0.0)
                      : (j == 0 ? 2707592511.0 / 1000000000000.0 - 1006699707.0 / 1000000000000.0 * s7
                         : j == 1 ? -51527976591.0 / 1000000000000.0 - 1006699707.0 / 1000000000000.0 * s7
                         : j == 2 ? -610366393.0 / 75000000000.0 + 7046897949.0 / 1000000000000.0 * s7 +
                                        14508670449.0 / 1000000000000.0 * s7 * s3
                         : j == 3 ? -610366393.0 / 75000000000.0 + 7046897949.0 / 1000000000000.0 * s7 -
                                        14508670449.0 / 1000000000000.0 * s7 * s3
                         : j == 4 ? -12456457.0 / 1171875000.0 + 1006699707.0 / 109375000000.0 * s7
                         : j == 5 ? 3020099121.0 / 437500000000.0 * s7 + 47328957.0 / 625000000.0
                         : j == 6 ? -7046897949.0 / 250000000000.0 * s7 : 0.0);
    }
    MIRK_HD static constexpr double tau_star() { return 0.4; }
    MIRK_HD static void weights(double t, double* w, double* wp) {
        const double t2 = t * t, t3 = t2 * t, t4 = t2 * t2, t5 = t4 * t, tm1 = t - 1.0;
        // the quartic shared by the weights of stages 3, 4 and 5, and the common factor of their derivatives
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
    }
};

}  // namespace mirk
