// abd_block.cuh — blocked merge of the ABD reduction for LARGE blocks (n = 64, 128: BASELINE config C4).
//
// Same algorithm, pivots-by-magnitude rule, relation and factor formats as abd.cuh / abd_mma32.cuh: a group of
// consecutive relations is collapsed by row-pivoted Gauss-Jordan on the n E-columns of the stacked 2n x 3n matrix
// [E | A | B] (+ rhs) of each merge.  At n = 128 that matrix is 768 KB — neither registers nor shared memory
// hold it — so it lives in a per-CTA global slab (L2 resident: 148 CTAs x 768 KB) and the elimination is BLOCKED:
//
//   per panel of NB = 32 pivot columns
//     (1) thread r loads its row's 32 panel entries into registers (one thread per row, 2n threads per CTA);
//     (2) 32 pivot steps on the panel alone, ONE block barrier each: every warp's best candidate publishes its
//         record (1/pivot, its remaining panel entries, its coefficients so far, its rhs), the winner is the
//         maximum of the per-warp keys, and every row applies  row += m * (pivot row)  to its panel entries while
//         accumulating the coefficients g[j] of
//             current row = row at panel start + sum_j g[j] * (pivot row j at panel start);
//     (3) the coefficients G (2n x 32) and the 32 pivot rows P (32 x trailing columns, as they were at panel start)
//         go to shared memory in DMMA fragment order (bank-conflict free strides);
//     (4) trailing update  W[:, behind the panel] += G * P  as rank-32 DMMA products (mma.sync.m8n8k4.f64): each warp
//         streams the 8-column tiles of its own 32 rows through registers (next tile prefetched while the current
//         one multiplies), 32 DMMA per 32 x 8 tile.
//   After n / 32 panels the pivot rows, scaled by 1/pivot, are the factors TL, TR, rt of the eliminated node and the
//   survivors shift E <- B, B <- 0 for the next relation of the group.
//
// Cost per merge at n = 128: ~12.6 M FMA (DMMA pipe: ~100 us on one SM at peak) + 4 x 32 pivot steps of latency;
// L2 traffic 4 x 2 x 600 KB.  This replaces k_reduce_generic's one-column-at-a-time elimination with three block
// barriers per pivot (0.13 TFLOP/s at C4 in round 1).
#pragma once
#include "abd_warp.cuh"

namespace mirk {

template <int n> struct BlockABD {
    static_assert(n % 32 == 0 && n >= 64 && n <= 128, "block path: n = 64, 96, 128");
    static constexpr int rows = 2 * n, cols = 3 * n, NB = 32, NP = n / NB, warps = rows / 32;
    static constexpr int CS = rows + 4;             // Gs[k][row], stride = 4 (mod 16): conflict-free A fragments
    static constexpr int TMAX = cols - NB;          // most trailing columns behind a panel
    static constexpr int PS = (TMAX / 16) * 16 + 4 + ((TMAX % 16) > 4 ? 16 : 0);  // Ps[k][col], stride = 4 (mod 16)
    static constexpr int REC = 64;                  // record: [0] 1/pivot, [c] panel entry c (c > k), [32 + j] g[j] (j < k), [63] rhs
    static constexpr int oP = NB * CS, oREC = oP + NB * PS, smem_doubles = oREC + 2 * warps * REC;
    static constexpr int rows_per_warp_copy = NB / warps;  // pivot rows each warp stages
    static constexpr size_t slab_doubles = (size_t)rows * cols;
};

// Row-pivoted Gauss-Jordan on the first 32 * npanels columns of W (2n rows, pitch ld, ctot columns), one thread per
// row: the panel loop described at the top of this file.  rhs is the thread's right-hand side entry; on return
// rows with myq >= 0 pivoted on column myq (scale myinv = 1 / pivot), the others survive.  All 2n threads call it;
// returns false (block-uniform) on a zero / non-finite pivot.  bsm = BlockABD<n>::smem_doubles of shared memory.
template <int n>
__device__ __forceinline__ bool block_gj_panels(double* W, int ld, int ctot, int npanels, double& rhs, int& myq, double& myinv,
                                                double* bsm) {
    using BA = BlockABD<n>;
    constexpr int NB = BA::NB, CS = BA::CS, PS = BA::PS, REC = BA::REC, NW = BA::warps;
    __shared__ unsigned skey[2][NW];
    __shared__ int s_pr[NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(bsm);
    double* Gs = bsm;
    double* Ps = bsm + BA::oP;
    double* rec = bsm + BA::oREC;
myq = -1;
myinv = 0.0;
bool elig = true, bad = false;
for (int pn = 0; pn < npanels; pn++) {
        const int q0 = NB * pn, c0 = q0 + NB, Tc = ctot - c0;  // trailing columns [c0, ctot)
        // (1) my row's panel entries
        double pe[NB], gc[NB];
        {
            const double2* src = reinterpret_cast<const double2*>(W + (size_t)tid * ld + q0);
#pragma unroll
            for (int c = 0; c < NB; c += 2) {
                const double2 v = src[c >> 1];
                pe[c] = v.x; pe[c + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < NB; j++) gc[j] = 0.0;
        }
        // (2) 32 pivot steps, one block barrier each
#pragma unroll
        for (int k = 0; k < NB; k++) {
            const double own = pe[k];
            const double own_inv = fast_rcp(own);
            const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~255u) | (unsigned)(255 - tid)) : 0u;
            const unsigned wmx = __reduce_max_sync(kFullMask, key);
            double* recw = rec + ((k & 1) * NW + warp) * REC;
            if (wmx != 0u && key == wmx) {
                recw[0] = own_inv;
#pragma unroll
                for (int c = k + 1; c < NB; c++) recw[c] = pe[c];
#pragma unroll
                for (int j = 0; j < k; j++) recw[32 + j] = gc[j];
                recw[63] = rhs;
            }
            if (lane == 0) skey[k & 1][warp] = wmx;
            __syncthreads();
            unsigned mx = 0u;
            int ww = 0;
#pragma unroll
            for (int w2 = 0; w2 < NW; w2++) {
                const unsigned kv = skey[k & 1][w2];
                if (kv > mx) { mx = kv; ww = w2; }
            }
            bad |= (mx >> 8) == 0u || mx >= 0x7ff00000u;
            const int prow = 255 - (int)(mx & 255u);
            const unsigned ra = sa + 8u * (unsigned)(BA::oREC + ((k & 1) * NW + ww) * REC);
            const double inv = lds_f64(ra);
            const bool isp = tid == prow;
            const double m = isp ? 0.0 : -(own * inv);
#pragma unroll
            for (int c = k + 1; c < NB; c++) pe[c] = fma(m, lds_f64(ra + 8u * (unsigned)c), pe[c]);
#pragma unroll
            for (int j = 0; j < k; j++) gc[j] = fma(m, lds_f64(ra + 8u * (unsigned)(32 + j)), gc[j]);
            gc[k] = m;
            rhs = fma(m, lds_f64(ra + 8u * 63u), rhs);
            if (isp) { elig = false; myq = q0 + k; myinv = own_inv; }
            if (tid == 0) s_pr[k] = prow;
        }
        if (bad) {  // block-uniform: every thread read the same keys
            return false;
        }
        // (3) coefficients and the pivot rows (as at panel start) to shared memory
#pragma unroll
        for (int j = 0; j < NB; j++) Gs[j * CS + tid] = gc[j];
        __syncthreads();  // s_pr complete; every warp is past its last record read
#pragma unroll
        for (int i = 0; i < BA::rows_per_warp_copy; i++) {
            const int kk = warp * BA::rows_per_warp_copy + i;
            const double* src = W + (size_t)s_pr[kk] * ld + c0;
            for (int c = lane; c < Tc; c += 32) Ps[kk * PS + c] = src[c];
        }
        __syncthreads();
        // (4) trailing update of this warp's 32 rows: 8-column tiles, next tile prefetched
        {
            double a[4][8];
#pragma unroll
            for (int tr = 0; tr < 4; tr++)
#pragma unroll
                for (int ks = 0; ks < 8; ks++)
                    a[tr][ks] = lds_f64(sa + 8u * (unsigned)((4 * ks + t) * CS + 32 * warp + 8 * tr + g));
            double* wr[4];
#pragma unroll
            for (int tr = 0; tr < 4; tr++) wr[tr] = W + (size_t)(32 * warp + 8 * tr + g) * ld + c0 + 2 * t;
            // tiles come from L2 / DRAM (the 148 slabs do not all stay in L2): PF tiles are in flight ahead of the
            // one being multiplied, in a register ring
            const int ntile = Tc >> 3;
            constexpr int PF = 4;
            double2 ring[PF][4];
#pragma unroll
            for (int f = 0; f < PF; f++)
                if (f < ntile) {
#pragma unroll
                    for (int tr = 0; tr < 4; tr++) ring[f][tr] = *reinterpret_cast<const double2*>(wr[tr] + 8 * f);
                }
            for (int jt0 = 0; jt0 < ntile; jt0 += PF) {
#pragma unroll
                for (int f = 0; f < PF; f++) {
                    const int jt = jt0 + f;
                    if (jt < ntile) {
                        double cc[4][2];
#pragma unroll
                        for (int tr = 0; tr < 4; tr++) { cc[tr][0] = ring[f][tr].x; cc[tr][1] = ring[f][tr].y; }
                        if (jt + PF < ntile) {
#pragma unroll
                            for (int tr = 0; tr < 4; tr++) ring[f][tr] = *reinterpret_cast<const double2*>(wr[tr] + 8 * (jt + PF));
                        }
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) {
                            const double b = lds_f64(sa + 8u * (unsigned)(BA::oP + (4 * ks + t) * PS + 8 * jt + g));
#pragma unroll
                            for (int tr = 0; tr < 4; tr++) dmma_8x8x4(cc[tr], a[tr][ks], b);
                        }
#pragma unroll
                        for (int tr = 0; tr < 4; tr++) *reinterpret_cast<double2*>(wr[tr] + 8 * jt) = make_double2(cc[tr][0], cc[tr][1]);
                    }
                }
            }
        }
        __syncwarp();  // my row was updated by the other lanes of my warp
    }
    return true;
}

template <int n>
__global__ void __launch_bounds__(2 * n, 1)
k_reduce_block(const double* __restrict__ inL, const double* __restrict__ inR, const double* __restrict__ inr,
               double* __restrict__ outL, double* __restrict__ outR, double* __restrict__ outr,
               const int* __restrict__ nodes, const int* __restrict__ gs, double* __restrict__ TL,
               double* __restrict__ TR, double* __restrict__ rt, double* __restrict__ scratch, int* __restrict__ status) {
    using BA = BlockABD<n>;
    constexpr int rows = BA::rows, cols = BA::cols, NW = BA::warps;
    constexpr size_t nn = (size_t)n * n;
    extern __shared__ __align__(16) double bsm[];
    __shared__ unsigned sbal[NW];
    __shared__ int s_free[n], s_q[rows];
    __shared__ double s_inv[rows];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, T = rows;
    const int grp = blockIdx.x, k0 = gs[grp], k1 = gs[grp + 1];
    if (k1 - k0 == 1) {  // nothing to eliminate: pass the relation through
        for (int e = tid; e < (int)nn; e += T) {
            outL[grp * nn + e] = inL[k0 * nn + e];
            outR[grp * nn + e] = inR[k0 * nn + e];
        }
        for (int e = tid; e < n; e += T) outr[(size_t)grp * n + e] = inr[(size_t)k0 * n + e];
        return;
    }
    double* W = scratch + (size_t)grp * BA::slab_doubles;

    // carried rows 0..n-1 <- relation k0:  [E | A | B] = [R | L | 0]
    for (int e = tid; e < (int)nn; e += T) {
        const int q = e / n, c = e % n;
        double* row = W + (size_t)q * cols;
        row[c] = inR[k0 * nn + e];
        row[n + c] = inL[k0 * nn + e];
        row[2 * n + c] = 0.0;
    }
    double rhs = tid < n ? inr[(size_t)k0 * n + tid] : 0.0;
    bool carried = tid < n;

    for (int jrel = k0 + 1; jrel < k1; jrel++) {
        // ---- the free rows, in row order, receive the incoming relation:  [E | A | B] = [L | 0 | R] ----
        {
            const unsigned bal = __ballot_sync(kFullMask, !carried);
            if (lane == 0) sbal[warp] = bal;
            __syncthreads();
            int idx = __popc(bal & ((1u << lane) - 1u));
            for (int w2 = 0; w2 < warp; w2++) idx += __popc(sbal[w2]);
            if (!carried) {
                s_free[idx] = tid;
                rhs = inr[(size_t)jrel * n + idx];
            }
            __syncthreads();
            for (int e = tid; e < (int)nn; e += T) {
                const int q = e / n, c = e % n;
                double* row = W + (size_t)s_free[q] * cols;
                row[c] = inL[jrel * nn + e];
                row[n + c] = 0.0;
                row[2 * n + c] = inR[jrel * nn + e];
            }
            __syncthreads();
        }
        int myq;
        double myinv;
        if (!block_gj_panels<n>(W, cols, cols, BA::NP, rhs, myq, myinv, bsm)) {
            if (tid == 0) atomicExch(status, 1);
            return;
        }
        // ---- factors of the eliminated node; survivors shift E <- B, B <- 0 ------------------------------
        const int cnode = nodes[jrel];
        s_q[tid] = myq;
        s_inv[tid] = myinv;
        if (myq >= 0) rt[(size_t)cnode * n + myq] = rhs * myinv;
        __syncthreads();
        double* TLc = TL + cnode * nn;
        double* TRc = TR + cnode * nn;
        for (int rr = 0; rr < 32; rr++) {
            const int R = 32 * warp + rr, q = s_q[R];
            double* row = W + (size_t)R * cols;
            if (q >= 0) {
                const double inv = s_inv[R];
                for (int c = lane; c < n; c += 32) {
                    TLc[(size_t)q * n + c] = row[n + c] * inv;
                    TRc[(size_t)q * n + c] = row[2 * n + c] * inv;
                }
            } else {
                for (int c = lane; c < n; c += 32) {
                    row[c] = row[2 * n + c];
                    row[2 * n + c] = 0.0;
                }
            }
        }
        carried = myq < 0;
        __syncthreads();
    }
    // ---- the n carried rows, in row order, as the collapsed relation of the group -------------------------
    {
        const unsigned bal = __ballot_sync(kFullMask, carried);
        if (lane == 0) sbal[warp] = bal;
        __syncthreads();
        int idx = __popc(bal & ((1u << lane) - 1u));
        for (int w2 = 0; w2 < warp; w2++) idx += __popc(sbal[w2]);
        s_q[tid] = carried ? idx : -1;
        if (carried) outr[(size_t)grp * n + idx] = rhs;
        __syncthreads();
        double* oL = outL + grp * nn;
        double* oR = outR + grp * nn;
        for (int rr = 0; rr < 32; rr++) {
            const int R = 32 * warp + rr, q = s_q[R];
            if (q < 0) continue;
            const double* row = W + (size_t)R * cols;
            for (int c = lane; c < n; c += 32) {
                oR[(size_t)q * n + c] = row[c];
                oL[(size_t)q * n + c] = row[n + c];
            }
        }
    }
}

// Closing solve of a two-point problem on the block path: the 2n x 2n system of the last relation and the n boundary
// rows on the two kept nodes,
//     [ Ba  0  ] [d_first]   [bc_a]
//     [ L   R  ] [d_last ] = [ r  ]      (rows in any order: the elimination pivots over all of them)
//     [ 0   Bb ]             [bc_b]
// by the same blocked Gauss-Jordan (2n rows, 2n pivot columns, 2n / 32 panels).  One CTA of 2n threads; the matrix
// lives in the global scratch M (2n x 2n).  Replaces k_final_solve's column-at-a-time elimination (3.7 ms at n = 128).
template <int n>
__global__ void __launch_bounds__(2 * n, 1)
k_final_block(const int* __restrict__ kept, const double* __restrict__ relL, const double* __restrict__ relR,
              const double* __restrict__ relr, int L, int La, const int* __restrict__ m_ptr, const int* __restrict__ bc_nodes,
              const double* __restrict__ Bc, const double* __restrict__ resid, size_t tail_off, double* __restrict__ M,
              double* __restrict__ delta, int* __restrict__ status) {
    using BA = BlockABD<n>;
    constexpr int D = 2 * n;
    extern __shared__ __align__(16) double bsm[];
    const int tid = threadIdx.x, T = D;
    for (int e = tid; e < D * D; e += T) M[e] = 0.0;
    __syncthreads();
    // boundary rows [0, L): block k of Bc belongs to the kept node bc_nodes[k]
    const int m = *m_ptr;
    for (int e = tid; e < L * n; e += T) {
        const int q = e / n, c = e % n;
        for (int k = 0; k < m; k++) {
            const int bn = bc_nodes[k];
            const int slot = bn == kept[0] ? 0 : bn == kept[1] ? 1 : -1;
            if (slot < 0) { atomicExch(status, 2); continue; }
            M[(size_t)q * D + slot * n + c] += Bc[((size_t)k * L + q) * n + c];
        }
    }
    // relation rows [L, L + n)
    for (int e = tid; e < n * n; e += T) {
        const int q = e / n, c = e % n;
        M[(size_t)(L + q) * D + c] = relL[e];
        M[(size_t)(L + q) * D + n + c] = relR[e];
    }
    double rhs = tid < L ? (tid < La ? resid[tid] : resid[tail_off + (tid - La)]) : relr[tid - L];
    __syncthreads();
    int myq;
    double myinv;
    if (!block_gj_panels<n>(M, D, D, D / BA::NB, rhs, myq, myinv, bsm)) {
        if (tid == 0) atomicExch(status, 1);
        return;
    }
    if (myq >= 0) delta[(size_t)kept[myq / n] * n + myq % n] = rhs * myinv;
}

inline bool block_reduce_supported(int n) { return n == 64 || n == 128; }
template <int n> inline int block_reduce_smem_bytes() { return (int)(sizeof(double) * BlockABD<n>::smem_doubles); }
inline size_t block_reduce_slab_doubles(int n) { return (size_t)2 * n * 3 * n; }

// closing solve for Q = 2 kept nodes and L = n boundary rows (two-point problems)
inline cudaError_t launch_block_final(cudaStream_t st, int n, const int* kept, const double* relL, const double* relR,
                                      const double* relr, int L, int La, const int* m_ptr, const int* bc_nodes, const double* Bc,
                                      const double* resid, size_t tail_off, double* M, double* delta, int* status) {
    if (n == 128) {
        static const cudaError_t e = cudaFuncSetAttribute(k_final_block<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          block_reduce_smem_bytes<128>());
        if (e != cudaSuccess) return e;
        k_final_block<128><<<1, 256, block_reduce_smem_bytes<128>(), st>>>(kept, relL, relR, relr, L, La, m_ptr, bc_nodes, Bc, resid,
                                                                           tail_off, M, delta, status);
    } else if (n == 64) {
        static const cudaError_t e = cudaFuncSetAttribute(k_final_block<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          block_reduce_smem_bytes<64>());
        if (e != cudaSuccess) return e;
        k_final_block<64><<<1, 128, block_reduce_smem_bytes<64>(), st>>>(kept, relL, relR, relr, L, La, m_ptr, bc_nodes, Bc, resid,
                                                                         tail_off, M, delta, status);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

inline cudaError_t launch_block_reduce(cudaStream_t st, int n, int G, const double* inL, const double* inR, const double* inr,
                                       double* outL, double* outR, double* outr, const int* nodes, const int* gs, double* TL,
                                       double* TR, double* rt, double* scratch, int* status) {
    if (n == 128) {
        static const cudaError_t e = cudaFuncSetAttribute(k_reduce_block<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          block_reduce_smem_bytes<128>());
        if (e != cudaSuccess) return e;
        k_reduce_block<128><<<G, 256, block_reduce_smem_bytes<128>(), st>>>(inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR,
                                                                            rt, scratch, status);
    } else if (n == 64) {
        static const cudaError_t e = cudaFuncSetAttribute(k_reduce_block<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          block_reduce_smem_bytes<64>());
        if (e != cudaSuccess) return e;
        k_reduce_block<64><<<G, 128, block_reduce_smem_bytes<64>(), st>>>(inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR, rt,
                                                                          scratch, status);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace mirk
