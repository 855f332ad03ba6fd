// abd_mma32.cuh — the n = 32 merge of the ABD reduction in DMMA fragment layout, four warps per merge.
// The library's default n = 32 reduction (launch_pair_reduce; MIRK_ABD_MMA32=0 selects k_reduce_pair for A/B runs).
// Against k_reduce_pair<32> on synthetic relations (experiments/exp_mma32.cu, B200): factors equal to 7e-15, 1.8x
// faster at 100 groups, 1.25x at 2000 groups (profiles/r01_s3_final/exp_mma32.log); the C5 golden tests run through it.
//
// Extends abd_mma.cuh (n = 16, one warp): the 64 x 96 matrix [E | A | B] of a merge is 8 x 12 tiles of 8 x 8; warp w
// owns tile rows 2w, 2w+1 (rows 16w .. 16w+15) in the C-fragment layout of mma.sync.m8n8k4.f64 (48 doubles per lane,
// as at n = 16).  Lanes 0..15 of a warp "hold" the warp's 16 rows for the panel factorisation (panel entries,
// coefficients, rhs, pivot bookkeeping), so the panel gather, the coefficient scatter and the factor write-out stay
// inside the warp; only the pivot decision (every warp publishes its best candidate's record, one block barrier per
// pivot) and the four published pivot rows of a panel cross warps.  experiments/mma32_merge_emul.py is the
// lane-by-lane numpy emulation of this data movement.  Same relation and factor formats as abd_pair.cuh.
#pragma once
#include "abd_warp.cuh"

namespace mirk {

struct MmaABD32 {
    static constexpr int n = 32, TRL = 2, TJ = 12;   // tile rows per warp, tile columns
    static constexpr int CSW = 20;                   // column stride of a warp's gathered panel / coefficients
    static constexpr int PS = 100;                   // doubles per published pivot row (96 + pad)
    static constexpr int REC = 8;                    // candidate record: inv, pe[1..3], gc[0..2], rhs0
    // shared memory (doubles): per-warp panels [4][4*CSW] | records [2][4][REC] | pivot rows [4][PS]
    static constexpr int oREC = 4 * 4 * CSW, oP = oREC + 2 * 4 * REC, smem_doubles = oP + 4 * PS;
};

// One level of the reduction for n = 32: CTA g (128 threads) collapses relations [gs[g], gs[g+1]).
__global__ void __launch_bounds__(128, 2)
k_reduce_mma32(const double* __restrict__ inL, const double* __restrict__ inR, const double* __restrict__ inr,
               double* __restrict__ outL, double* __restrict__ outR, double* __restrict__ outr,
               const int* __restrict__ nodes, const int* __restrict__ gs, double* __restrict__ TL,
               double* __restrict__ TR, double* __restrict__ rt, int* __restrict__ status) {
    using MA = MmaABD32;
    constexpr int n = 32;
    constexpr size_t nn = (size_t)n * n;
    __shared__ __align__(16) double sm[MA::smem_doubles];
    __shared__ unsigned skey[2][4], sb[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const bool holder = lane < 16;
    const int hrow = 16 * warp + lane;  // the row a holder lane holds
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned swp = sa + 8u * (unsigned)(warp * 4 * MA::CSW);  // this warp's panel / coefficient area
    const int grp = blockIdx.x, k0 = gs[grp], k1 = gs[grp + 1];

    double w[MA::TRL][MA::TJ][2];
    double rhs = 0.0;
    // carried rows 0..31 (warps 0 and 1):  [E | A | B | rhs] = [R | L | 0 | r]
    {
        const double* Lk = inL + k0 * nn;
        const double* Rk = inR + k0 * nn;
#pragma unroll
        for (int trl = 0; trl < MA::TRL; trl++) {
#pragma unroll
            for (int j = 0; j < MA::TJ; j++) { w[trl][j][0] = 0.0; w[trl][j][1] = 0.0; }
            const int r = 16 * warp + 8 * trl + g;
            if (r < n) {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    const double2 e = *reinterpret_cast<const double2*>(Rk + r * n + 8 * jj + 2 * t);
                    const double2 a = *reinterpret_cast<const double2*>(Lk + r * n + 8 * jj + 2 * t);
                    w[trl][jj][0] = e.x; w[trl][jj][1] = e.y;
                    w[trl][4 + jj][0] = a.x; w[trl][4 + jj][1] = a.y;
                }
            }
        }
        if (holder && hrow < n) rhs = inr[(size_t)k0 * n + hrow];
    }
    unsigned long long carried = 0x00000000ffffffffull;
    for (int jrel = k0 + 1; jrel < k1; jrel++) {
        const unsigned long long freem = ~carried;
        // incoming rows into the free row slots, in row order:  [E | A | B | rhs] = [L | 0 | R | r]
        {
            const double* Lk = inL + jrel * nn;
            const double* Rk = inR + jrel * nn;
#pragma unroll
            for (int trl = 0; trl < MA::TRL; trl++) {
                const int r = 16 * warp + 8 * trl + g;
                if ((freem >> r) & 1ull) {
                    const int idx = __popcll(freem & ((1ull << r) - 1ull));
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        const double2 e = *reinterpret_cast<const double2*>(Lk + idx * n + 8 * jj + 2 * t);
                        const double2 b = *reinterpret_cast<const double2*>(Rk + idx * n + 8 * jj + 2 * t);
                        w[trl][jj][0] = e.x; w[trl][jj][1] = e.y;
                        w[trl][4 + jj][0] = 0.0; w[trl][4 + jj][1] = 0.0;
                        w[trl][8 + jj][0] = b.x; w[trl][8 + jj][1] = b.y;
                    }
                }
            }
            if (holder && ((freem >> hrow) & 1ull)) rhs = inr[(size_t)jrel * n + __popcll(freem & ((1ull << hrow) - 1ull))];
        }
        // ---- Gauss-Jordan on the 32 E columns, 8 panels of 4 ------------------------------------------
        int myq = -1;
        double myinv = 0.0;
        bool elig = holder;
        bool bad = false;
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            const int q0 = 4 * pn, jp = q0 >> 3, cq = q0 & 7, t0 = cq >> 1;
            // (A) the warp's 16 x 4 panel into lane-per-row form (warp-local)
            if (t == t0 || t == t0 + 1) {
#pragma unroll
                for (int trl = 0; trl < MA::TRL; trl++) {
                    sts_f64(swp + 8u * (unsigned)((2 * (t - t0)) * MA::CSW + 8 * trl + g), w[trl][jp][0]);
                    sts_f64(swp + 8u * (unsigned)((2 * (t - t0) + 1) * MA::CSW + 8 * trl + g), w[trl][jp][1]);
                }
            }
            __syncwarp();
            double pe[4] = {0.0, 0.0, 0.0, 0.0};
            if (holder) {
#pragma unroll
                for (int c = 0; c < 4; c++) pe[c] = lds_f64(swp + 8u * (unsigned)(c * MA::CSW + lane));
            }
            // (B) 4 pivot steps: every warp publishes the record of its best candidate, one barrier per pivot
            double gc[4] = {0.0, 0.0, 0.0, 0.0};
            int pr[4];
            double rhs0p[4];
            const double rhs0 = rhs;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double own = pe[k];
                const double own_inv = fast_rcp(own);
                const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~63u) | (unsigned)(63 - hrow)) : 0u;
                const unsigned wmx = __reduce_max_sync(kFullMask, key);
                const unsigned recw = sa + 8u * (unsigned)(MA::oREC + ((k & 1) * 4 + warp) * MA::REC);
                if (wmx != 0u && key == wmx) {
                    sts_v2f64(recw, own_inv, pe[1]);
                    sts_v2f64(recw + 16u, pe[2], pe[3]);
                    sts_v2f64(recw + 32u, gc[0], gc[1]);
                    sts_v2f64(recw + 48u, gc[2], rhs0);
                }
                if (lane == 0) skey[k & 1][warp] = wmx;
                __syncthreads();
                const unsigned ka = skey[k & 1][0], kb = skey[k & 1][1], kc = skey[k & 1][2], kd = skey[k & 1][3];
                const unsigned mab = ka > kb ? ka : kb, mcd = kc > kd ? kc : kd, mx = mab > mcd ? mab : mcd;
                const int ww = mx == ka ? 0 : mx == kb ? 1 : mx == kc ? 2 : 3;
                bad |= (mx >> 6) == 0u || mx >= 0x7ff00000u;
                pr[k] = 63 - (int)(mx & 63u);
                const unsigned rec = sa + 8u * (unsigned)(MA::oREC + ((k & 1) * 4 + ww) * MA::REC);
                const double2 r01 = lds_v2f64(rec), r23 = lds_v2f64(rec + 16u), r45 = lds_v2f64(rec + 32u),
                              r67 = lds_v2f64(rec + 48u);
                const double inv = r01.x;
                const double ppe[4] = {0.0, r01.y, r23.x, r23.y};
                const double pgc[3] = {r45.x, r45.y, r67.x};
                rhs0p[k] = r67.y;
                const bool isp = holder && hrow == pr[k];
                const double m = (holder && !isp) ? -(own * inv) : 0.0;
#pragma unroll
                for (int c = k + 1; c < 4; c++) pe[c] = fma(m, ppe[c], pe[c]);
#pragma unroll
                for (int j = 0; j < k; j++) gc[j] = fma(m, pgc[j], gc[j]);
                gc[k] = m;
                if (isp) { elig = false; myq = q0 + k; myinv = own_inv; }
            }
            if (bad) {  // block-uniform: every thread read the same keys
                if (tid == 0) atomicExch(status, 1);
                return;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) rhs = fma(gc[j], rhs0p[j], rhs);
            // (C) coefficients over the gathered panel (warp-local; every holder consumed it long ago)
            if (holder) {
#pragma unroll
                for (int j = 0; j < 4; j++) sts_f64(swp + 8u * (unsigned)(j * MA::CSW + lane), gc[j]);
            }
            // (D) the 4 pivot rows as they were at panel start into the block-wide lines
            const int jlo = cq == 0 ? jp : jp + 1;
#pragma unroll
            for (int trl = 0; trl < MA::TRL; trl++) {
                const int r = 16 * warp + 8 * trl + g;
                const int kk = r == pr[0] ? 0 : r == pr[1] ? 1 : r == pr[2] ? 2 : r == pr[3] ? 3 : -1;
                if (kk >= 0) {
                    const unsigned line = sa + 8u * (unsigned)(MA::oP + kk * MA::PS + 2 * t);
#pragma unroll
                    for (int j = jlo; j < MA::TJ; j++) sts_v2f64(line + 8u * (unsigned)(8 * j), w[trl][j][0], w[trl][j][1]);
                }
            }
            __syncthreads();
            // (E) fragments and the rank-4 update of the live tiles, every warp on its own two tile rows
            double a[MA::TRL];
#pragma unroll
            for (int trl = 0; trl < MA::TRL; trl++) a[trl] = lds_f64(swp + 8u * (unsigned)(t * MA::CSW + 8 * trl + g));
#pragma unroll
            for (int j = jlo; j < MA::TJ; j++) {
                const double b = lds_f64(sa + 8u * (unsigned)(MA::oP + t * MA::PS + 8 * j + g));
#pragma unroll
                for (int trl = 0; trl < MA::TRL; trl++) dmma_8x8x4(w[trl][j], a[trl], b);
            }
            __syncthreads();  // lines, records and panels are rewritten by the next panel
        }
        // ---- factors of the eliminated node; survivors shift E <- B, B <- 0 ----------------------------
        const int c = nodes[jrel];
        double* TLc = TL + c * nn;
        double* TRc = TR + c * nn;
#pragma unroll
        for (int trl = 0; trl < MA::TRL; trl++) {
            const int hl = 8 * trl + g;  // the holder lane of this row, same warp
            const int q = __shfl_sync(kFullMask, myq, hl);
            const double inv = __shfl_sync(kFullMask, myinv, hl);
            if (q >= 0) {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    *reinterpret_cast<double2*>(TLc + q * n + 8 * jj + 2 * t) = make_double2(w[trl][4 + jj][0] * inv, w[trl][4 + jj][1] * inv);
                    *reinterpret_cast<double2*>(TRc + q * n + 8 * jj + 2 * t) = make_double2(w[trl][8 + jj][0] * inv, w[trl][8 + jj][1] * inv);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    w[trl][jj][0] = w[trl][8 + jj][0]; w[trl][jj][1] = w[trl][8 + jj][1];
                    w[trl][8 + jj][0] = 0.0; w[trl][8 + jj][1] = 0.0;
                }
            }
        }
        if (holder && myq >= 0) rt[(size_t)c * n + myq] = rhs * myinv;
        // 64-bit mask of the pivot rows: 16 holder lanes per warp
        const unsigned bal = __ballot_sync(kFullMask, holder && myq >= 0) & 0xffffu;
        if (lane == 0) sb[warp] = bal;
        __syncthreads();
        carried = ~((unsigned long long)sb[0] | ((unsigned long long)sb[1] << 16) | ((unsigned long long)sb[2] << 32) |
                    ((unsigned long long)sb[3] << 48));
        __syncthreads();
    }
    // the 32 carried rows, in row order, as the collapsed relation of the group
    {
        double* oL = outL + grp * nn;
        double* oR = outR + grp * nn;
#pragma unroll
        for (int trl = 0; trl < MA::TRL; trl++) {
            const int r = 16 * warp + 8 * trl + g;
            if ((carried >> r) & 1ull) {
                const int idx = __popcll(carried & ((1ull << r) - 1ull));
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    *reinterpret_cast<double2*>(oR + idx * n + 8 * jj + 2 * t) = make_double2(w[trl][jj][0], w[trl][jj][1]);
                    *reinterpret_cast<double2*>(oL + idx * n + 8 * jj + 2 * t) = make_double2(w[trl][4 + jj][0], w[trl][4 + jj][1]);
                }
            }
        }
        if (holder && ((carried >> hrow) & 1ull)) outr[(size_t)grp * n + __popcll(carried & ((1ull << hrow) - 1ull))] = rhs;
    }
}

}  // namespace mirk
