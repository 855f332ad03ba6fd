// abd_warp.cuh — warp-level paths of the ABD reduction for small blocks (2n <= 32) and the multi-level kernel.
//
// Same algorithm, relation format and factor layout as abd.cuh (Wright-style stable block cyclic reduction).
//   * n = 2, 4, 6, 8: one WARP per group of relations and one LANE per row of the stacked 2n x (3n+1) working
//     matrix [E | A | B | rhs] (WarpABD): each lane keeps its row in registers; the pivot of a column is one
//     REDUX.MAX over the high words of |w[q]| (partial pivoting to ~2^-20 relative, plenty for stability); pivot
//     rows stay unscaled and are scaled once when written out as factors.  MIRK_ELIM_PANEL (the build default)
//     takes the pivots in panels of 4: inside a panel the pivot lane's few entries travel by shuffle, and the 4
//     pivot rows are published once per panel through shared memory; without it every pivot row is broadcast
//     through a double-buffered shared line.
//   * n = 16: the merge runs in DMMA fragment layout (abd_mma.cuh, included below), MIRK_ABD_MMA (build default).
//   * k_tail_warp: several radix-2 levels per launch — one block for the tail (plus the closing solve and the
//     matching back substitutions), many blocks for a segment of upper levels, each walking its own sub-tree.
// The next relation of a group is staged with cp.async while the current merge runs.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include "abd.cuh"

namespace mirk {

constexpr unsigned kFullMask = 0xffffffffu;

// 1/x for a pivot: hardware reciprocal seed (MUFU.RCP64H via rcp.approx.ftz.f64) + two Newton steps,
// branch-free (a full IEEE division drags a slow-path subroutine into the unrolled elimination).
// Pivots are finite and non-zero here; relative error <= ~2 ulp.
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}

template <int n> struct WarpABD {
    static constexpr int rows = 2 * n, cols = 3 * n + 1;
    static constexpr int pb_stride = (cols + 2) & ~1;  // doubles per broadcast line, even
    // per-warp shared memory: the double-buffered pivot line + a staging buffer the NEXT relation of the
    // group is prefetched into with cp.async while the current merge runs (which lanes will be free to
    // receive it is only known after the merge, so it cannot be prefetched into registers)
    static constexpr int stage_stride = 2 * n + 2;  // doubles per staged row [L row | R row | r | pad], 16-byte multiple
    static constexpr int stage_doubles = n * stage_stride;
    // panel elimination (MIRK_ELIM_PANEL): PB pivots per panel; the PB pivot rows of a panel are published
    // once, as PB shared lines holding columns [SH, cols)
    static constexpr int PB = 4;
    static constexpr int SH = n < PB ? n : PB;
    static constexpr int LS = (cols - SH + 1) & ~1;  // doubles per published line, even
#if defined(MIRK_ELIM_PANEL)
    static constexpr int line_doubles = PB * LS;
#else
    static constexpr int line_doubles = 2 * pb_stride;
#endif
    static constexpr int smem_doubles_per_warp = line_doubles + stage_doubles;

    __device__ __forceinline__ static void stage_issue(double* st, const double* Lk, const double* Rk, const double* rk,
                                                       int lane) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(st);
        constexpr int chunks_per_row = n / 2;  // 16-byte chunks in one n-double row
        for (int e = lane; e < 2 * n * chunks_per_row; e += 32) {
            const int which = e / (n * chunks_per_row), rem = e % (n * chunks_per_row);
            const int q = rem / chunks_per_row, ch = rem % chunks_per_row;
            const double* src = (which ? Rk : Lk) + q * n + 2 * ch;
            const unsigned dst = sa + 8u * (q * stage_stride + which * n + 2 * ch);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        }
        if (lane < n) {
            const unsigned dst = sa + 8u * (lane * stage_stride + 2 * n);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(rk + lane));
        }
        asm volatile("cp.async.commit_group;");
    }
    __device__ __forceinline__ static void stage_wait() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    // staged relation row q into an incoming row  [E | A | B | rhs] = [L | 0 | R | r]
    __device__ __forceinline__ static void load_incoming_staged(double (&w)[cols], const double* st, int q) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(st) + 8u * (q * stage_stride);
#pragma unroll
        for (int k = 0; k < n; k += 2) {
            const double2 a = lds_v2f64(sa + 8u * k), b = lds_v2f64(sa + 8u * (n + k));
            w[k] = a.x; w[k + 1] = a.y;
            w[n + k] = 0.0; w[n + k + 1] = 0.0;
            w[2 * n + k] = b.x; w[2 * n + k + 1] = b.y;
        }
        w[3 * n] = lds_f64(sa + 8u * (2 * n));
    }

    // Gauss-Jordan on the n E-columns over all `rows` lanes.  On return lanes with myq >= 0 own unit
    // column myq up to the scale myinv (= 1 / pivot); lanes with myq < 0 are the survivors.
    // Returns false (warp-uniform) on a zero / non-finite pivot.  n must be even (vector accesses).
    __device__ __forceinline__ static bool eliminate(double (&w)[cols], int lane, double* __restrict__ pb,
                                                     int& myq, double& myinv) {
        myq = -1;
        myinv = 0.0;
        bool elig = lane < rows;
#if defined(MIRK_ELIM_PANEL)
        // Panel form of the same Gauss-Jordan elimination (same pivots).  Inside a panel of PB columns only the
        // panel entries are updated, from the pivot lane by warp shuffles (a short dependent chain per pivot:
        // REDUX -> SHFL -> DFMA), while every lane accumulates the coefficients g[j] of
        //     current row = row at panel start + sum_j g[j] * (pivot row j at panel start).
        // The PB pivot lanes then publish their rows at once and every lane applies its PB-term combination to
        // all later columns: one shared-memory round trip per panel instead of one per pivot.
        const unsigned la = (unsigned)__cvta_generic_to_shared(pb);
#pragma unroll
        for (int q0 = 0; q0 < n; q0 += PB) {
            const int pw = (n - q0) < PB ? (n - q0) : PB;
            double g[PB];
#pragma unroll
            for (int j = 0; j < PB; j++) g[j] = 0.0;
            int myrank = -1;
#pragma unroll
            for (int k = 0; k < PB; k++) {
                if (k < pw) {
                    const int q = q0 + k;
                    const double own = w[q];
                    const double own_inv = fast_rcp(own);
                    const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~31u) | (unsigned)(31 - lane)) : 0u;
                    const unsigned mx = __reduce_max_sync(kFullMask, key);
                    if ((mx >> 5) == 0u || mx >= 0x7ff00000u) return false;
                    const int pr = 31 - (int)(mx & 31u);
                    const bool isp = lane == pr;
                    const double inv = __shfl_sync(kFullMask, own_inv, pr);
                    const double m = isp ? 0.0 : -(own * inv);
#pragma unroll
                    for (int c = q + 1; c < q0 + pw; c++) w[c] = fma(m, __shfl_sync(kFullMask, w[c], pr), w[c]);
#pragma unroll
                    for (int j = 0; j < k; j++) g[j] = fma(m, __shfl_sync(kFullMask, g[j], pr), g[j]);
                    g[k] = m;
                    if (isp) { elig = false; myq = q; myinv = own_inv; myrank = k; }
                }
            }
            const int c0 = q0 + pw;  // first column behind the panel (even: n and PB are even)
            if (myrank >= 0) {
                double* line = pb + myrank * LS;
#pragma unroll
                for (int c = c0; c < cols; c += 2) {
                    const double hi = (c + 1 < cols) ? w[(c + 1 < cols) ? c + 1 : c] : 0.0;
                    *reinterpret_cast<double2*>(line + (c - SH)) = make_double2(w[c], hi);
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = c0; c < cols; c += 2) {
                double a0 = w[c], a1 = (c + 1 < cols) ? w[(c + 1 < cols) ? c + 1 : c] : 0.0;
#pragma unroll
                for (int j = 0; j < PB; j++) {
                    if (j < pw) {
                        const double2 v = lds_v2f64(la + 8u * (unsigned)(j * LS + c - SH));
                        a0 = fma(g[j], v.x, a0);
                        a1 = fma(g[j], v.y, a1);
                    }
                }
                w[c] = a0;
                if (c + 1 < cols) w[(c + 1 < cols) ? c + 1 : c] = a1;
            }
            __syncwarp();  // the lines are rewritten by the next panel
        }
        return true;
#else
#pragma unroll
        for (int q = 0; q < n; q++) {
            const double own = w[q];
            // speculative reciprocal of every lane's candidate: overlaps with the pivot search, so the
            // reciprocal is off the critical path of the pivot lane
            const double own_inv = fast_rcp(own);
            // one REDUX gives the pivot and its lane: high word of |w| (low 5 bits dropped) | (31 - lane)
            const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~31u) | (unsigned)(31 - lane)) : 0u;
            const unsigned mx = __reduce_max_sync(kFullMask, key);
            if ((mx >> 5) == 0u || mx >= 0x7ff00000u) return false;
            const int pr = 31 - (int)(mx & 31u);
#if defined(MIRK_BCAST_SHFL)
            // measured alternative: broadcast the pivot row by warp shuffles instead of shared memory
            if (lane == pr) { elig = false; myq = q; myinv = own_inv; }
            const double inv = __shfl_sync(kFullMask, own_inv, pr);
            const double m = (lane == pr) ? 0.0 : -(own * inv);
#pragma unroll
            for (int c = q + 1; c < cols; c++) w[c] = fma(m, __shfl_sync(kFullMask, w[c], pr), w[c]);
            (void)pb;
#else
            double* line = pb + (q & 1) * pb_stride;
            if (lane == pr) {
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
            __syncwarp();
            // volatile shared loads: nvcc otherwise forwards the pivot lane's own stores across the
            // __syncwarp and keeps a second copy of the row in registers (spills at n = 16)
            const unsigned la = (unsigned)__cvta_generic_to_shared(line);
            const double inv = lds_f64(la + 8u * q);
            const double m = (lane == pr) ? 0.0 : -(own * inv);
#pragma unroll
            for (int c = ((q + 1) & ~1); c < cols; c += 2) {
                const double2 v = lds_v2f64(la + 8u * c);
                if (c > q) w[c] = fma(m, v.x, w[c]);
                if (c + 1 < cols) w[(c + 1 < cols) ? c + 1 : c] = fma(m, v.y, w[(c + 1 < cols) ? c + 1 : c]);
            }
#endif
        }
        return true;
#endif  // MIRK_ELIM_PANEL
    }

    // relation row q of (L, R, r) into a carried row  [E | A | B | rhs] = [R | L | 0 | r]
    __device__ __forceinline__ static void load_carried(double (&w)[cols], const double* __restrict__ Lq,
                                                        const double* __restrict__ Rq, double rq) {
#pragma unroll
        for (int k = 0; k < n; k += 2) {
            const double2 a = *reinterpret_cast<const double2*>(Rq + k), b = *reinterpret_cast<const double2*>(Lq + k);
            w[k] = a.x; w[k + 1] = a.y;
            w[n + k] = b.x; w[n + k + 1] = b.y;
            w[2 * n + k] = 0.0; w[2 * n + k + 1] = 0.0;
        }
        w[3 * n] = rq;
    }
    // relation row q into an incoming row  [E | A | B | rhs] = [L | 0 | R | r]
    __device__ __forceinline__ static void load_incoming(double (&w)[cols], const double* __restrict__ Lq,
                                                         const double* __restrict__ Rq, double rq) {
#pragma unroll
        for (int k = 0; k < n; k += 2) {
            const double2 a = *reinterpret_cast<const double2*>(Lq + k), b = *reinterpret_cast<const double2*>(Rq + k);
            w[k] = a.x; w[k + 1] = a.y;
            w[n + k] = 0.0; w[n + k + 1] = 0.0;
            w[2 * n + k] = b.x; w[2 * n + k + 1] = b.y;
        }
        w[3 * n] = rq;
    }

    // after a merge: pivot lanes write the factors of the eliminated node c,
    //   d_c = rt - TL d_a - TR d_right ;  survivors shift E <- B, B <- 0
    __device__ __forceinline__ static void store_factors_and_shift(double (&w)[cols], int lane, int myq, double myinv,
                                                                   double* __restrict__ TLc, double* __restrict__ TRc,
                                                                   double* __restrict__ rtc) {
        if (myq >= 0) {
            double* tl = TLc + myq * n;
            double* tr = TRc + myq * n;
#pragma unroll
            for (int k = 0; k < n; k += 2) {
                *reinterpret_cast<double2*>(tl + k) = make_double2(w[n + k] * myinv, w[n + k + 1] * myinv);
                *reinterpret_cast<double2*>(tr + k) = make_double2(w[2 * n + k] * myinv, w[2 * n + k + 1] * myinv);
            }
            rtc[myq] = w[3 * n] * myinv;
        } else if (lane < rows) {
#pragma unroll
            for (int k = 0; k < n; k++) { w[k] = w[2 * n + k]; w[2 * n + k] = 0.0; }
        }
    }

    // the n carried rows, in lane order, as the collapsed relation g
    __device__ __forceinline__ static void store_relation(const double (&w)[cols], int lane, unsigned carried,
                                                          double* __restrict__ oL, double* __restrict__ oR,
                                                          double* __restrict__ orr) {
        if ((carried >> lane) & 1u) {
            const int idx = __popc(carried & ((1u << lane) - 1u));
#pragma unroll
            for (int k = 0; k < n; k += 2) {
                *reinterpret_cast<double2*>(oR + idx * n + k) = make_double2(w[k], w[k + 1]);
                *reinterpret_cast<double2*>(oL + idx * n + k) = make_double2(w[n + k], w[n + k + 1]);
            }
            orr[idx] = w[3 * n];
        }
    }
};

// One group of one reduction level, done by one warp (see k_reduce_generic for the argument meaning).
// Returns false (warp-uniform) on a singular block.
template <int n>
__device__ __forceinline__ bool warp_reduce_group(int g, const double* inL, const double* inR, const double* inr,
                                                  double* outL, double* outR, double* outr, const int* nodes,
                                                  const int* gs, double* TL, double* TR, double* rt, double* pbuf,
                                                  int lane) {
    using WA = WarpABD<n>;
    constexpr size_t nn = (size_t)n * n;
    const int k0 = gs[g], k1 = gs[g + 1];
    double w[WA::cols];
#pragma unroll
    for (int c = 0; c < WA::cols; c++) w[c] = 0.0;
    unsigned carried = (1u << n) - 1u;
    const unsigned rowmask = (WA::rows == 32) ? kFullMask : ((1u << WA::rows) - 1u);
    double* stage = pbuf + WA::line_doubles;
    if (k0 + 1 < k1) WA::stage_issue(stage, inL + (k0 + 1) * nn, inR + (k0 + 1) * nn, inr + (size_t)(k0 + 1) * n, lane);
    if (lane < n) WA::load_carried(w, inL + k0 * nn + (size_t)lane * n, inR + k0 * nn + (size_t)lane * n, inr[(size_t)k0 * n + lane]);
    for (int j = k0 + 1; j < k1; j++) {
        const unsigned freem = rowmask & ~carried;
        WA::stage_wait();
        if ((freem >> lane) & 1u) WA::load_incoming_staged(w, stage, __popc(freem & ((1u << lane) - 1u)));
        __syncwarp();  // every lane has its row before the buffer is refilled
        if (j + 1 < k1) WA::stage_issue(stage, inL + (j + 1) * nn, inR + (j + 1) * nn, inr + (size_t)(j + 1) * n, lane);
        int myq;
        double myinv;
        if (!WA::eliminate(w, lane, pbuf, myq, myinv)) return false;
        const int c = nodes[j];
        WA::store_factors_and_shift(w, lane, myq, myinv, TL + c * nn, TR + c * nn, rt + (size_t)c * n);
        carried = rowmask & ~__ballot_sync(kFullMask, myq >= 0);
        __syncwarp();
    }
    WA::store_relation(w, lane, carried, outL + g * nn, outR + g * nn, outr + (size_t)g * n);
    return true;
}

// Back substitution of one group: d_c = rt_c - TL_c d_a - TR_c d_right, right to left.  Lanes [0,n) own
// the rows of the TL product, lanes [16,16+n) those of the TR product; da/dr are 16-double shared lines.
template <int n>
__device__ __forceinline__ void warp_backsub_group(int g, const int* nodes, const int* gs, const double* TL,
                                                   const double* TR, const double* rt, double* delta, double* da,
                                                   double* dr, int lane, double* yup = nullptr) {
    const int k0 = gs[g], k1 = gs[g + 1];
    if (k1 - k0 == 1) return;
    constexpr size_t nn = (size_t)n * n;
    const int half = lane >> 4, q = lane & 15;
    const bool act = q < n;
    __syncwarp();
    if (lane < n) {
        da[lane] = delta[(size_t)nodes[k0] * n + lane];
        dr[lane] = delta[(size_t)nodes[k1] * n + lane];
    }
    __syncwarp();
    double va[n];
#pragma unroll
    for (int k = 0; k < n; k++) va[k] = da[k];
    // the factor rows of a node do not depend on the running solution: fetch node j-1's row (and rt) while
    // node j's products are in flight, so the right-to-left chain only carries the mat-vec, not the loads
    double rv[n], rtv = 0.0;
    if (act) {
        const int c = nodes[k1 - 1];
        const double* row = (half ? TR : TL) + c * nn + (size_t)q * n;
#pragma unroll
        for (int k = 0; k < n; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(row + k);
            rv[k] = v.x; rv[k + 1] = v.y;
        }
        if (lane < n) rtv = rt[(size_t)c * n + lane];
    }
    for (int j = k1 - 1; j > k0; j--) {
        const int c = nodes[j];
        double acc = 0.0;
        double nx[n], nrt = 0.0;
        if (act && j - 1 > k0) {
            const int cn = nodes[j - 1];
            const double* row = (half ? TR : TL) + cn * nn + (size_t)q * n;
#pragma unroll
            for (int k = 0; k < n; k += 2) {
                const double2 v = *reinterpret_cast<const double2*>(row + k);
                nx[k] = v.x; nx[k + 1] = v.y;
            }
            if (lane < n) nrt = rt[(size_t)cn * n + lane];
        }
        if (act) {
            // four partial sums: the chain is n/4 dependent FMAs instead of n
            double p4[4] = {0.0, 0.0, 0.0, 0.0};
            if (half == 0) {
#pragma unroll
                for (int k = 0; k < n; k++) p4[k & 3] = fma(rv[k], va[k], p4[k & 3]);
            } else {
#pragma unroll
                for (int k = 0; k < n; k++) p4[k & 3] = fma(rv[k], dr[k], p4[k & 3]);
            }
            acc = (p4[0] + p4[1]) + (p4[2] + p4[3]);
        }
        acc += __shfl_xor_sync(kFullMask, acc, 16);
        __syncwarp();
        if (lane < n) {
            const double d = rtv - acc;
            delta[(size_t)c * n + lane] = d;
            if (yup) yup[(size_t)c * n + lane] -= d;  // fused Newton update y -= delta (level 0 only)
            dr[lane] = d;
        }
        __syncwarp();
        if (j - 1 > k0) {
#pragma unroll
            for (int k = 0; k < n; k++) rv[k] = nx[k];
            rtv = nrt;
        }
    }
}

// Back substitution of one group for n = 16 with COALESCED factor loads.  warp_backsub_group gives every lane one factor
// row, so each of its 8 LDG.128 per node touches 32 different 128-byte lines and the level-0 pass (one factor pair per
// mesh node, 77 MB at C2) is bound by the L1 tag stage at ~3 TB/s.  Here lane i reads chunk i + 32k (k = 0..3) of the
// node's contiguous 2 KB TL block and the same of TR (four 128-byte lines per request), which leaves it with columns
// 2s, 2s+1 (s = i & 7) of rows 4k + (i >> 3): two-term partial dot products, summed over the 8 lanes of a row by three
// xor-shuffle steps.  The new node's solution reaches the lanes that need it as the next right-hand vector by two
// shuffles, so the right-to-left chain never touches shared memory.
__device__ __forceinline__ void warp_backsub_group16c(int g, const int* nodes, const int* gs, const double* TL, const double* TR,
                                                      const double* rt, double* delta, int lane, double* yup = nullptr) {
    constexpr int n = 16;
    constexpr size_t nn = (size_t)n * n;
    const int k0 = gs[g], k1 = gs[g + 1];
    if (k1 - k0 == 1) return;
    const int sub = lane & 7, rg = lane >> 3;
    const double2 da = *reinterpret_cast<const double2*>(delta + (size_t)nodes[k0] * n + 2 * sub);
    double2 dr = *reinterpret_cast<const double2*>(delta + (size_t)nodes[k1] * n + 2 * sub);
    // the factors do not depend on the running solution: nodes j-1 and j-2 are in flight while node j is computed
    // (two nodes = 8 KB per warp keep enough bytes in flight for the HBM latency at 11 warps per SM)
    double2 tl[4], tr[4], tl1[4], tr1[4];
    double rtv[4], rtv1[4];
    auto fetch = [&](int c, double2 (&l)[4], double2 (&r)[4], double (&t)[4]) {
        const double2* pl = reinterpret_cast<const double2*>(TL + c * nn) + lane;
        const double2* pr = reinterpret_cast<const double2*>(TR + c * nn) + lane;
#pragma unroll
        for (int k = 0; k < 4; k++) { l[k] = pl[32 * k]; r[k] = pr[32 * k]; t[k] = rt[(size_t)c * n + 4 * k + rg]; }
    };
    fetch(nodes[k1 - 1], tl, tr, rtv);
    if (k1 - 2 > k0) fetch(nodes[k1 - 2], tl1, tr1, rtv1);
    // sources of the next right-hand pair: rows 2s, 2s+1 live in the lanes of row groups 2(s&1), 2(s&1)+1, slot s >> 1
    const int src0 = 8 * (2 * (sub & 1)) + sub, src1 = src0 + 8;
    for (int j = k1 - 1; j > k0; j--) {
        const int c = nodes[j];
        double2 tl2[4], tr2[4];
        double rtv2[4];
        if (j - 2 > k0) fetch(nodes[j - 2], tl2, tr2, rtv2);
        double p[4];
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = fma(tl[k].x, da.x, tl[k].y * da.y) + fma(tr[k].x, dr.x, tr[k].y * dr.y);
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) p[k] += __shfl_xor_sync(kFullMask, p[k], m);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = rtv[k] - p[k];  // solution rows 4k + rg, in every lane of the row group
        const double own = (sub & 2) ? ((sub & 1) ? p[3] : p[2]) : ((sub & 1) ? p[1] : p[0]);  // row 4 (sub & 3) + rg
        if (sub < 4) {
            delta[(size_t)c * n + 4 * sub + rg] = own;
            if (yup) yup[(size_t)c * n + 4 * sub + rg] -= own;  // fused Newton update y -= delta (level 0 only)
        }
        const double v = (sub & 4) ? ((sub & 2) ? p[3] : p[2]) : ((sub & 2) ? p[1] : p[0]);  // slot sub >> 1
        dr.x = __shfl_sync(kFullMask, v, src0);
        dr.y = __shfl_sync(kFullMask, v, src1);
#pragma unroll
        for (int k = 0; k < 4; k++) { tl[k] = tl1[k]; tr[k] = tr1[k]; rtv[k] = rtv1[k]; tl1[k] = tl2[k]; tr1[k] = tr2[k]; rtv1[k] = rtv2[k]; }
    }
}

}  // namespace mirk
#include "abd_mma.cuh"  // n = 16 on the FP64 tensor path (needs WarpABD above; included from here only)
namespace mirk {

// the merge kernel of a given block size: DMMA-fragment path for n = 16, lane-per-row otherwise
template <int n, bool STAGE>
__device__ __forceinline__ bool reduce_group(int g, const double* inL, const double* inR, const double* inr, double* outL,
                                             double* outR, double* outr, const int* nodes, const int* gs, double* TL,
                                             double* TR, double* rt, double* pbuf, int lane) {
#if defined(MIRK_ABD_MMA)
    if constexpr (n == 16) return mma_reduce_group16<STAGE>(g, inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR, rt, pbuf, lane);
    else
#endif
        return warp_reduce_group<n>(g, inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR, rt, pbuf, lane);
}
template <int n, bool STAGE> __host__ __device__ constexpr int reduce_smem_doubles() {
#if defined(MIRK_ABD_MMA)
    if (n == 16) return MmaABD16::smem_doubles<STAGE>();
#endif
    return WarpABD<n>::smem_doubles_per_warp;
}

template <int n, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_reduce_warp(int G, const double* __restrict__ inL, const double* __restrict__ inR, const double* __restrict__ inr,
              double* __restrict__ outL, double* __restrict__ outR, double* __restrict__ outr,
              const int* __restrict__ nodes, const int* __restrict__ gs, double* __restrict__ TL,
              double* __restrict__ TR, double* __restrict__ rt, int* __restrict__ status) {
    __shared__ __align__(16) double pbuf[4][reduce_smem_doubles<n, true>()];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = blockIdx.x * (blockDim.x >> 5) + wib;
    if (g >= G) return;
    if (!reduce_group<n, true>(g, inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR, rt, pbuf[wib], lane))
        if (lane == 0) atomicExch(status, 1);
}

template <int n>
__global__ void __launch_bounds__(128)
k_backsub_warp(int G, const int* __restrict__ nodes, const int* __restrict__ gs, const double* __restrict__ TL,
               const double* __restrict__ TR, const double* __restrict__ rt, double* __restrict__ delta,
               double* __restrict__ yup) {
    __shared__ __align__(16) double dbuf[4][2][16];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = blockIdx.x * (blockDim.x >> 5) + wib;
    if (g >= G) return;
    if (yup && lane < n) {
        // fused update of the nodes this level does not recover: every group's left end, and the last node
        const int a = nodes[gs[g]];
        yup[(size_t)a * n + lane] -= delta[(size_t)a * n + lane];
        if (g == G - 1) {
            const int b = nodes[gs[g + 1]];
            yup[(size_t)b * n + lane] -= delta[(size_t)b * n + lane];
        }
    }
#if !defined(MIRK_BACKSUB_ROWS)
    if constexpr (n == 16) { warp_backsub_group16c(g, nodes, gs, TL, TR, rt, delta, lane, yup); return; }
#endif
    warp_backsub_group<n>(g, nodes, gs, TL, TR, rt, delta, dbuf[wib][0], dbuf[wib][1], lane, yup);
}

// ---- the tail: every remaining level, the closing solve and the matching back substitutions in ONE
// block, so the log-depth end of the reduction costs one launch instead of 2 x levels + 1.
constexpr int kMaxTail = 20;
struct TailArgs {
    int mode;  // bit 0: reduce the tail levels, bit 1: closing solve, bit 2: back-substitute the tail levels
    // multi = 1: a SEGMENT of <= 4 radix-2 levels on many blocks.  Block b owns groups [8b, 8b+8) of the segment's
    // first level, [4b, 4b+4) of the next, ... so each block walks its own sub-tree with block barriers only
    // (needs pure pairing after the first level: checked by the host planner).  multi = 0: one block, all groups.
    int multi;
    int nlev;
    int G[kMaxTail];
    const int* nodes[kMaxTail];
    const int* gs[kMaxTail];
    const double *inL[kMaxTail], *inR[kMaxTail], *inr[kMaxTail];
    double *outL[kMaxTail], *outR[kMaxTail], *outr[kMaxTail];
    double *TL, *TR, *rt;
    int* status;
    // closing system (see k_final_solve)
    int Q;
    const int* kept;
    const double *relL, *relR, *relr;
    int L, La;
    const int* m_ptr;
    const int* bc_nodes;
    const double* Bc;
    const double* resid;
    size_t tail_off;
    double* M;  // global scratch or nullptr when the closing matrix fits in shared memory
    double* delta;
};

constexpr int kTailWarps = 8;

// (A cooperative multi-CTA variant with one warp per SM and grid barriers between levels was measured
// slower: ~20 grid syncs cost more than the shared-memory contention they avoid — profiles/r01_notes.md.)
// shared-memory buffers of a tail / segment block (one set per kernel)
template <int n> struct TailSmem {
    double pbuf[kTailWarps][reduce_smem_doubles<n, false>()];
    double dbuf[kTailWarps][2][16];
};

template <int n>
__device__ __forceinline__ void tail_body(const TailArgs& a, TailSmem<n>& sm, double* tail_smem) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int l = 0; l < a.nlev && (a.mode & 1); l++) {
        const int per = a.multi ? ((kTailWarps >> l) > 0 ? (kTailWarps >> l) : 1) : a.G[l];
        const int gb = a.multi ? (int)blockIdx.x * per : 0;
        const int ge = gb + per < a.G[l] ? gb + per : a.G[l];
        for (int g = gb + wib; g < ge; g += kTailWarps) {
            if (!reduce_group<n, false>(g, a.inL[l], a.inR[l], a.inr[l], a.outL[l], a.outR[l], a.outr[l], a.nodes[l],
                                        a.gs[l], a.TL, a.TR, a.rt, sm.pbuf[wib], lane))
                if (lane == 0) atomicExch(a.status, 1);
        }
        __syncthreads();
    }
    if (a.mode & 2)
        final_solve_body(n, a.Q, a.kept, a.relL, a.relR, a.relr, a.L, a.La, a.m_ptr, a.bc_nodes, a.Bc, a.resid,
                         a.tail_off, a.M, a.delta, a.status, tail_smem);
    __syncthreads();
    for (int l = a.nlev - 1; l >= 0 && (a.mode & 4); l--) {
        const int per = a.multi ? ((kTailWarps >> l) > 0 ? (kTailWarps >> l) : 1) : a.G[l];
        const int gb = a.multi ? (int)blockIdx.x * per : 0;
        const int ge = gb + per < a.G[l] ? gb + per : a.G[l];
        for (int g = gb + wib; g < ge; g += kTailWarps)
            warp_backsub_group<n>(g, a.nodes[l], a.gs[l], a.TL, a.TR, a.rt, a.delta, sm.dbuf[wib][0], sm.dbuf[wib][1], lane);
        __syncthreads();
    }
}

// (A cooperative multi-CTA variant with one warp per SM and grid barriers between levels was measured
// slower: ~20 grid syncs cost more than the shared-memory contention they avoid — profiles/r01_notes.md.)
template <int n>
__global__ void __launch_bounds__(kTailWarps * 32, 1)
k_tail_warp(const TailArgs a) {
    extern __shared__ double tail_smem[];
    __shared__ __align__(16) TailSmem<n> sm;
    tail_body<n>(a, sm, tail_smem);
}

// ---- mesh-partitioned mode, peer-memory exchange (abd.cuh): the interface step of a Newton iteration in TWO
// kernels instead of seven graph nodes.
//   k_part_tail_push   the local tail levels of this segment (as k_tail_warp, mode 1) and, from the same block, the
//                      push of the collapsed relation into every peer's exchange buffer
//   k_part_interface   wait for all ranks' relations, unpack, reduce + close + back-substitute the interface system
//                      (k_tail_warp on the interface plan, mode 7), hand this rank's two end-node updates to the local
//                      system and back-substitute the local tail levels (mode 4)
struct PartPushArgs {
    int L, La;
    const double *relL, *relR, *relr, *Bc, *resid;
    size_t tail_off;
    XchgPeers peers;
    XchgLayout lay;
    int rank;
    unsigned long long* epoch;
};
struct PartIfaceArgs {
    int L, La, standard, rank, Nloc;
    double* xbuf;
    XchgLayout lay;
    double *if_L, *if_R, *if_r, *if_Bc, *if_resid, *if_delta, *delta;
    unsigned long long* epoch;
    int* status;
};

template <int n>
__global__ void __launch_bounds__(kTailWarps * 32, 1)
k_part_tail_push(const TailArgs a, const PartPushArgs ps) {
    extern __shared__ double tail_smem[];
    __shared__ __align__(16) TailSmem<n> sm;
    if (a.nlev > 0) tail_body<n>(a, sm, tail_smem);
    __syncthreads();  // the collapsed relation is complete (global writes + block barrier)
    part_push_body(n, ps.L, ps.La, ps.relL, ps.relR, ps.relr, ps.Bc, ps.resid, ps.tail_off, ps.peers, ps.lay, ps.rank, ps.epoch);
}

template <int n>
__global__ void __launch_bounds__(kTailWarps * 32, 1)
k_part_interface(const PartIfaceArgs w, const TailArgs ai, const TailArgs al) {
    extern __shared__ double tail_smem[];
    __shared__ __align__(16) TailSmem<n> sm;
    part_wait_unpack_body(n, w.L, w.La, w.standard, w.xbuf, w.lay, w.if_L, w.if_R, w.if_r, w.if_Bc, w.if_resid, w.epoch, w.status);
    __syncthreads();
    tail_body<n>(ai, sm, tail_smem);  // interface system: reduce, closing solve, back substitution
    __syncthreads();
    if (threadIdx.x < n) {  // this rank's two end nodes
        w.delta[threadIdx.x] = w.if_delta[(size_t)w.rank * n + threadIdx.x];
        w.delta[(size_t)(w.Nloc - 1) * n + threadIdx.x] = w.if_delta[(size_t)(w.rank + 1) * n + threadIdx.x];
    }
    __syncthreads();
    if (al.nlev > 0) tail_body<n>(al, sm, tail_smem);  // local tail levels, back substitution only (mode 4)
}

inline bool warp_reduce_supported(int n) { return n == 2 || n == 4 || n == 6 || n == 8 || n == 16; }

#define MIRK_WARP_DISPATCH(n, CALL) \
    switch (n) {                    \
    case 2: { constexpr int NN = 2; CALL; } break;   \
    case 4: { constexpr int NN = 4; CALL; } break;   \
    case 6: { constexpr int NN = 6; CALL; } break;   \
    case 8: { constexpr int NN = 8; CALL; } break;   \
    case 16: { constexpr int NN = 16; CALL; } break; \
    default: break;                 \
    }

inline void launch_warp_reduce(cudaStream_t st, int n, int G, const double* inL, const double* inR, const double* inr,
                               double* outL, double* outR, double* outr, const int* nodes, const int* gs, double* TL,
                               double* TR, double* rt, int* status) {
    // few groups: one warp per CTA so the merges spread over the SMs (each merge is bound by its SM's
    // shared-memory broadcast pipe); many groups: four warps per CTA
    const int wpb = G < 1024 ? 1 : 4;
    const int blocks = (G + wpb - 1) / wpb;
    static const int occ = getenv("MIRK_REDUCE_OCC") ? atoi(getenv("MIRK_REDUCE_OCC")) : 3;  // tuning knob
    if (occ >= 4) {
        MIRK_WARP_DISPATCH(n, (k_reduce_warp<NN, 4><<<blocks, 32 * wpb, 0, st>>>(G, inL, inR, inr, outL, outR, outr, nodes, gs,
                                                                           TL, TR, rt, status)));
    } else {
        MIRK_WARP_DISPATCH(n, (k_reduce_warp<NN, 3><<<blocks, 32 * wpb, 0, st>>>(G, inL, inR, inr, outL, outR, outr, nodes, gs,
                                                                           TL, TR, rt, status)));
    }
}
inline cudaError_t launch_part_tail_push(cudaStream_t st, int n, const TailArgs& a, const PartPushArgs& ps) {
    MIRK_WARP_DISPATCH(n, (k_part_tail_push<NN><<<1, kTailWarps * 32, 0, st>>>(a, ps)));
    return cudaGetLastError();
}
template <int NN> inline void launch_part_interface_n(cudaStream_t st, const PartIfaceArgs& w, const TailArgs& ai, const TailArgs& al,
                                                      int smem_bytes) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, k_part_interface<NN>) == cudaSuccess) {
            const int room = 227 * 1024 - (int)fa.sharedSizeBytes - 1024;
            cudaFuncSetAttribute(k_part_interface<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, room < 160 * 1024 ? room : 160 * 1024);
        }
        cudaGetLastError();
        attr_set = true;
    }
    k_part_interface<NN><<<1, kTailWarps * 32, smem_bytes, st>>>(w, ai, al);
}
inline cudaError_t launch_part_interface(cudaStream_t st, int n, const PartIfaceArgs& w, const TailArgs& ai, const TailArgs& al,
                                         int smem_bytes) {
    MIRK_WARP_DISPATCH(n, (launch_part_interface_n<NN>(st, w, ai, al, smem_bytes)));
    return cudaGetLastError();
}
inline cudaError_t launch_warp_tail(cudaStream_t st, int n, const TailArgs& a, int ctas, int smem_bytes) {
    MIRK_WARP_DISPATCH(n, (k_tail_warp<NN><<<ctas, kTailWarps * 32, smem_bytes, st>>>(a)));
    return cudaGetLastError();
}
// opt in to as much dynamic shared memory as fits beside the kernel's static allocation
inline void set_warp_tail_smem(int n, int bytes) {
    MIRK_WARP_DISPATCH(n, {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, k_tail_warp<NN>) == cudaSuccess) {
            const int room = 227 * 1024 - (int)fa.sharedSizeBytes - 1024;
            cudaFuncSetAttribute(k_tail_warp<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes < room ? bytes : room);
        }
        cudaGetLastError();
    });
}
inline void launch_warp_backsub(cudaStream_t st, int n, int G, const int* nodes, const int* gs, const double* TL,
                                const double* TR, const double* rt, double* delta, double* yup = nullptr) {
    const int wpb = G < 1024 ? 1 : 4;
    const int blocks = (G + wpb - 1) / wpb;
    MIRK_WARP_DISPATCH(n, (k_backsub_warp<NN><<<blocks, 32 * wpb, 0, st>>>(G, nodes, gs, TL, TR, rt, delta, yup)));
}

}  // namespace mirk
