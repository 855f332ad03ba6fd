// abd_pair.cuh — register-resident ABD reduction for n = 32 (BASELINE config C5): the stacked
// 2n x (3n+1) = 64 x 97 working matrix is spread over TWO warps, one thread per row, 97 doubles of the
// row in registers.  Same algorithm, relation format and factor layout as abd.cuh / abd_warp.cuh; one
// CTA of 64 threads per group.  Per pivot: a REDUX per warp + one shared word per warp gives the pivot row
// (high word of |w|, low 6 bits = 63 - row), the pivot thread publishes its row through a double-buffered
// shared line, two __syncthreads per pivot.
#pragma once
#include <cuda_runtime.h>

#include "abd_warp.cuh"

namespace mirk {

template <int n> struct PairABD {
    static_assert(n == 32, "two warps hold exactly 64 rows");
    static constexpr int rows = 2 * n, cols = 3 * n + 1;
    static constexpr int pb_stride = (cols + 2) & ~1;

    __device__ __forceinline__ static bool eliminate(double (&w)[cols], int row, double* pb, volatile unsigned* skey,
                                                     int& myq, double& myinv) {
        myq = -1;
        myinv = 0.0;
        bool elig = true;
        const int warp = row >> 5, lane = row & 31;
#pragma unroll
        for (int q = 0; q < n; q++) {
            const double own = w[q];
            const double own_inv = fast_rcp(own);
            const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~63u) | (unsigned)(63 - row)) : 0u;
            const unsigned wmx = __reduce_max_sync(kFullMask, key);
            volatile unsigned* sk = skey + (q & 1) * 2;
            if (lane == 0) sk[warp] = wmx;
            __syncthreads();
            const unsigned k0 = sk[0], k1 = sk[1];
            const unsigned mx = k0 > k1 ? k0 : k1;
            if ((mx >> 6) == 0u || mx >= 0x7ff00000u) return false;
            const int pr = 63 - (int)(mx & 63u);
            double* line = pb + (q & 1) * pb_stride;
            if (row == pr) {
#pragma unroll
                for (int c = (q & ~1); c < cols; c += 2) {
                    const double lo = (c == q) ? own_inv : w[c];
                    const double hi = (c + 1 == q) ? own_inv : ((c + 1 < cols) ? w[(c + 1 < cols) ? c + 1 : c] : 0.0);
                    *reinterpret_cast<double2*>(line + c) = make_double2(lo, hi);
                }
                elig = false;
                myq = q;
                myinv = own_inv;
            }
            __syncthreads();
            const unsigned la = (unsigned)__cvta_generic_to_shared(line);
            const double inv = lds_f64(la + 8u * q);
            const double m = (row == pr) ? 0.0 : -(own * inv);
#pragma unroll
            for (int c = ((q + 1) & ~1); c < cols; c += 2) {
                const double2 v = lds_v2f64(la + 8u * c);
                if (c > q) w[c] = fma(m, v.x, w[c]);
                if (c + 1 < cols) w[(c + 1 < cols) ? c + 1 : c] = fma(m, v.y, w[(c + 1 < cols) ? c + 1 : c]);
            }
        }
        return true;
    }
};

// 64-bit mask of the rows for which `pred` holds (both warps), via one shared word per warp
__device__ __forceinline__ unsigned long long pair_ballot(bool pred, int row, volatile unsigned* sb) {
    const unsigned b = __ballot_sync(kFullMask, pred);
    __syncthreads();
    if ((row & 31) == 0) sb[row >> 5] = b;
    __syncthreads();
    return (unsigned long long)sb[0] | ((unsigned long long)sb[1] << 32);
}

// One level of the reduction for n = 32: CTA g (64 threads) collapses relations [gs[g], gs[g+1]).
template <int n>
__global__ void __launch_bounds__(64, 4)
k_reduce_pair(const double* __restrict__ inL, const double* __restrict__ inR, const double* __restrict__ inr,
              double* __restrict__ outL, double* __restrict__ outR, double* __restrict__ outr,
              const int* __restrict__ nodes, const int* __restrict__ gs, double* __restrict__ TL,
              double* __restrict__ TR, double* __restrict__ rt, int* __restrict__ status) {
    using PA = PairABD<n>;
    using WA = WarpABD<n>;  // row load / store helpers are layout-identical
    __shared__ __align__(16) double pbuf[2 * PA::pb_stride];
    __shared__ unsigned skey[4], sb[2];
    constexpr size_t nn = (size_t)n * n;
    const int row = threadIdx.x, g = blockIdx.x;
    const int k0 = gs[g], k1 = gs[g + 1];
    double w[PA::cols];
#pragma unroll
    for (int c = 0; c < PA::cols; c++) w[c] = 0.0;
    unsigned long long carried = (1ull << n) - 1ull;  // rows 0..n-1 hold relation k0
    if (row < n) WA::load_carried(w, inL + k0 * nn + (size_t)row * n, inR + k0 * nn + (size_t)row * n, inr[(size_t)k0 * n + row]);
    for (int j = k0 + 1; j < k1; j++) {
        const unsigned long long freem = ~carried;
        if ((freem >> row) & 1ull) {
            const int q = __popcll(freem & ((1ull << row) - 1ull));
            WA::load_incoming(w, inL + j * nn + (size_t)q * n, inR + j * nn + (size_t)q * n, inr[(size_t)j * n + q]);
        }
        int myq;
        double myinv;
        if (!PA::eliminate(w, row, pbuf, skey, myq, myinv)) {
            if (row == 0) atomicExch(status, 1);
            return;
        }
        const int c = nodes[j];
        // factors of node c from the pivot rows; survivors shift E <- B, B <- 0
        if (myq >= 0) {
            double* tl = TL + c * nn + (size_t)myq * n;
            double* tr = TR + c * nn + (size_t)myq * n;
#pragma unroll
            for (int k = 0; k < n; k += 2) {
                *reinterpret_cast<double2*>(tl + k) = make_double2(w[n + k] * myinv, w[n + k + 1] * myinv);
                *reinterpret_cast<double2*>(tr + k) = make_double2(w[2 * n + k] * myinv, w[2 * n + k + 1] * myinv);
            }
            rt[(size_t)c * n + myq] = w[3 * n] * myinv;
        } else {
#pragma unroll
            for (int k = 0; k < n; k++) { w[k] = w[2 * n + k]; w[2 * n + k] = 0.0; }
        }
        carried = ~pair_ballot(myq >= 0, row, sb);
    }
    if ((carried >> row) & 1ull) {
        const int idx = __popcll(carried & ((1ull << row) - 1ull));
        double* oR = outR + g * nn + (size_t)idx * n;
        double* oL = outL + g * nn + (size_t)idx * n;
#pragma unroll
        for (int k = 0; k < n; k += 2) {
            *reinterpret_cast<double2*>(oR + k) = make_double2(w[k], w[k + 1]);
            *reinterpret_cast<double2*>(oL + k) = make_double2(w[n + k], w[n + k + 1]);
        }
        outr[(size_t)g * n + idx] = w[3 * n];
    }
}

// Back substitution of one level for n = 32: warp 0 owns the rows of the TL product, warp 1 those of TR.
template <int n>
__global__ void __launch_bounds__(64)
k_backsub_pair(const int* __restrict__ nodes, const int* __restrict__ gs, const double* __restrict__ TL,
               const double* __restrict__ TR, const double* __restrict__ rt, double* __restrict__ delta) {
    __shared__ __align__(16) double da[n], dr[n], part[n];
    const int g = blockIdx.x, k0 = gs[g], k1 = gs[g + 1];
    if (k1 - k0 == 1) return;
    constexpr size_t nn = (size_t)n * n;
    const int half = threadIdx.x >> 5, q = threadIdx.x & 31;
    if (threadIdx.x < n) {
        da[q] = delta[(size_t)nodes[k0] * n + q];
        dr[q] = delta[(size_t)nodes[k1] * n + q];
    }
    __syncthreads();
    for (int j = k1 - 1; j > k0; j--) {
        const int c = nodes[j];
        const double* rowp = (half ? TR : TL) + c * nn + (size_t)q * n;
        const double* vec = half ? dr : da;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < n; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(rowp + k);
            acc = fma(v.x, vec[k], acc);
            acc = fma(v.y, vec[k + 1], acc);
        }
        if (half) part[q] = acc;
        __syncthreads();
        if (!half) {
            const double d = rt[(size_t)c * n + q] - acc - part[q];
            delta[(size_t)c * n + q] = d;
            dr[q] = d;
        }
        __syncthreads();
    }
}

}  // namespace mirk
#include "abd_mma32.cuh"  // four-warp DMMA-fragment merge: the default n = 32 reduction (MIRK_ABD_MMA32=0 selects k_reduce_pair)
namespace mirk {

inline bool pair_reduce_supported(int n) { return n == 32; }

inline void launch_pair_reduce(cudaStream_t st, int G, const double* inL, const double* inR, const double* inr,
                               double* outL, double* outR, double* outr, const int* nodes, const int* gs, double* TL,
                               double* TR, double* rt, int* status) {
    // default: the four-warp DMMA-fragment merge (abd_mma32.cuh); MIRK_ABD_MMA32=0 keeps the two-warp
    // thread-per-row kernel for A/B measurements (same pivots, relation and factor formats)
    static const bool use_mma = !(getenv("MIRK_ABD_MMA32") && atoi(getenv("MIRK_ABD_MMA32")) == 0);
    if (use_mma) k_reduce_mma32<<<G, 128, 0, st>>>(inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR, rt, status);
    else k_reduce_pair<32><<<G, 64, 0, st>>>(inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR, rt, status);
}
inline void launch_pair_backsub(cudaStream_t st, int G, const int* nodes, const int* gs, const double* TL,
                                const double* TR, const double* rt, double* delta) {
    k_backsub_pair<32><<<G, 64, 0, st>>>(nodes, gs, TL, TR, rt, delta);
}

}  // namespace mirk
