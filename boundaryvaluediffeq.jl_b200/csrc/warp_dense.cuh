// warp_dense.cuh — one-warp dense solve and the pivot reciprocal shared by the ABD paths (abd.cuh, abd_warp.cuh) and
// the warp-per-trajectory ensemble kernel (ensemble_warp.cuh).  Header-only inline device code.
#pragma once
#include <cuda_runtime.h>

namespace mirk {

// 1/x for a pivot: hardware reciprocal seed (MUFU.RCP64H via rcp.approx.ftz.f64) + two Newton steps,
// branch-free (a full IEEE division drags a slow-path subroutine into the unrolled elimination).
// Pivots are finite and non-zero here; relative error <= ~2 ulp.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// Dense D x D solve (D <= 32) by one warp, in place on the shared-memory matrix [M | rhs] (ld doubles per row):
// Gauss-Jordan with row pivoting, lane r owns row r, pivot by one REDUX over the high words of |M[r][q]|, the
// pivot row is read by broadcast.  Rolled loops on purpose: the kernel runs this once, from a cold instruction
// cache, so code size counts more than instruction count (the unrolled register version took ~20 us).
// The lane that pivoted on column e writes unknown e to delta[kept[e / n] * n + e % n].
__device__ __forceinline__ void warp_dense_solve32(double* M, int D, int ld, const int* kept, int n, double* delta,
                                                   int* status) {
    const int lane = threadIdx.x & 31;
    double* row = M + (size_t)(lane < D ? lane : 0) * ld;
    bool elig = lane < D;
    int myq = -1;
    double myinv = 0.0;
    for (int q = 0; q < D; q++) {
        const double own = lane < D ? row[q] : 0.0;
        const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~31u) | (unsigned)(31 - lane)) : 0u;
        const unsigned mx = __reduce_max_sync(0xffffffffu, key);
        if ((mx >> 5) == 0u || mx >= 0x7ff00000u) {  // warp-uniform
            if (lane == 0) atomicExch(status, 1);
            return;
        }
        const int pr = 31 - (int)(mx & 31u);
        const double* prow = M + (size_t)pr * ld;
        const double inv = fast_rcp(prow[q]);
        if (lane == pr) { elig = false; myq = q; myinv = inv; }
        else if (lane < D) {
            const double f = -(own * inv);
#pragma unroll 4
            for (int c = q + 1; c <= D; c++) row[c] = fma(f, prow[c], row[c]);
        }
        __syncwarp();
    }
    if (myq >= 0) delta[(size_t)kept[myq / n] * n + myq % n] = row[D] * myinv;
}

}  // namespace mirk
