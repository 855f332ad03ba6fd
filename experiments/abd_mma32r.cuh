// abd_mma32r.cuh — EXPERIMENT (not in the library): n = 32 merge with the panel factorisation replicated in every warp.
// Result on the B200 (experiments/exp_mma32r.cu): bit-identical factors and relations, but NOT faster than k_reduce_mma32
// (8.44 vs 8.08 ms for 250 000 relations, 12.3 vs 11.8 us per merge on an idle GPU): the barrier per pivot was not the
// bound.  Cycle stamps of one merge: per panel ~2100 cycles = gather + barrier 105 | 4 pivots 910 | coefficient + pivot-row
// publication 400-590 (24 STS.128 with 4 active lanes each) + barrier | 24 DMMA 350-420.
#pragma once
#include "abd_pair.cuh"
namespace mirk {

// ---- variant R: the panel factorisation REPLICATED in every warp, no barrier per pivot -----------------------------
// k_reduce_mma32 above spends ~560 cycles per pivot: per-warp candidate records through shared memory and ONE BLOCK
// BARRIER PER PIVOT (32 per merge).  Here every warp gathers the whole 64 x 4 panel (lane l holds rows l and l + 32) and
// runs the 4 pivot steps itself — identical arithmetic in all four warps, one REDUX per pivot, the pivot lane's entries
// by shuffle — exactly the chain of the n = 16 merge (abd_mma.cuh), ~150 cycles per pivot.  A panel then costs TWO block
// barriers (after the gather, after the publication of the pivot rows) and the merge needs none for its bookkeeping
// (every warp derives the same carried mask from its own ballots).  Same pivots, factors and relations as
// k_reduce_mma32 / k_reduce_pair<32>.
#if defined(MIRK_R32_PROF)  // experiments: cycle stamps of one CTA's first merge
__device__ long long g_r32_prof[96];
__device__ int g_r32_prof_n;
#define R32_STAMP() do { if (tid == 0 && blockIdx.x == 0 && g_r32_prof_n < 96) g_r32_prof[g_r32_prof_n++] = clock64(); } while (0)
#else
#define R32_STAMP() do { } while (0)
#endif
struct MmaABD32R {
    static constexpr int n = 32, TRL = 2, TJ = 12;
    static constexpr int CS = 68;    // column stride of the gathered 64 x 4 panel (2 * CS doubles = 16 banks: conflict-free scatter)
    static constexpr int CSG = 20;   // column stride of a warp's coefficients
    static constexpr int PS = 100;   // doubles per published pivot row (96 + pad)
    static constexpr int oG = 4 * CS, oP = oG + 4 * 4 * CSG, smem_doubles = oP + 4 * PS;
};

__global__ void __launch_bounds__(128, 2)
k_reduce_mma32r(const double* __restrict__ inL, const double* __restrict__ inR, const double* __restrict__ inr,
                double* __restrict__ outL, double* __restrict__ outR, double* __restrict__ outr,
                const int* __restrict__ nodes, const int* __restrict__ gs, double* __restrict__ TL,
                double* __restrict__ TR, double* __restrict__ rt, int* __restrict__ status) {
    using MA = MmaABD32R;
    constexpr int n = 32;
    constexpr size_t nn = (size_t)n * n;
    __shared__ __align__(16) double sm[MA::smem_doubles];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wslot = warp >> 1;                 // the row slot this warp's own 16 rows live in (rows 16w .. 16w + 15)
    const bool wown = (lane >> 4) == (warp & 1); // this lane holds one of the warp's own rows (in slot wslot)
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned sg = sa + 8u * (unsigned)(MA::oG + warp * 4 * MA::CSG);
    const int grp = blockIdx.x, k0 = gs[grp], k1 = gs[grp + 1];

    double w[MA::TRL][MA::TJ][2];
    double rhs[2] = {0.0, 0.0};  // rows lane, lane + 32
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
        rhs[0] = inr[(size_t)k0 * n + lane];
    }
    unsigned long long carried = 0x00000000ffffffffull;
    bool bad = false;
    for (int jrel = k0 + 1; jrel < k1; jrel++) {
        const unsigned long long freem = ~carried;
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
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int row = lane + 32 * sl;
                if ((freem >> row) & 1ull) rhs[sl] = inr[(size_t)jrel * n + __popcll(freem & ((1ull << row) - 1ull))];
            }
        }
        int myq[2] = {-1, -1};
        double myinv[2] = {0.0, 0.0};
        bool elig[2] = {true, true};
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            const int q0 = 4 * pn, jp = q0 >> 3, cq = q0 & 7, t0 = cq >> 1;
            // (A) this warp's 16 rows of the panel into the block's lane-per-row copy
            R32_STAMP();
            if (t == t0 || t == t0 + 1) {
#pragma unroll
                for (int trl = 0; trl < MA::TRL; trl++) {
                    const int row = 16 * warp + 8 * trl + g;
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0)) * MA::CS + row), w[trl][jp][0]);
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0) + 1) * MA::CS + row), w[trl][jp][1]);
                }
            }
            __syncthreads();
            R32_STAMP();
            double pe[2][4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                pe[0][c] = lds_f64(sa + 8u * (unsigned)(c * MA::CS + lane));
                pe[1][c] = lds_f64(sa + 8u * (unsigned)(c * MA::CS + 32 + lane));
            }
            // (B) 4 pivot steps, replicated in every warp
            double gc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
            int pr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double own0 = pe[0][k], own1 = pe[1][k];
                const double inv0 = fast_rcp(own0), inv1 = fast_rcp(own1);
                const unsigned key0 = elig[0] ? (((unsigned)__double2hiint(fabs(own0)) & ~63u) | (unsigned)(63 - lane)) : 0u;
                const unsigned key1 = elig[1] ? (((unsigned)__double2hiint(fabs(own1)) & ~63u) | (unsigned)(31 - lane)) : 0u;
                const unsigned mx = __reduce_max_sync(kFullMask, key0 > key1 ? key0 : key1);
                bad |= (mx >> 6) == 0u || mx >= 0x7ff00000u;
                pr[k] = 63 - (int)(mx & 63u);
                const int src = pr[k] & 31;
                const bool s1 = pr[k] >= 32;  // warp-uniform: which slot of lane `src` holds the pivot row
                const bool isp0 = !s1 && lane == src, isp1 = s1 && lane == src;
                const double inv = __shfl_sync(kFullMask, s1 ? inv1 : inv0, src);
                const double m0 = isp0 ? 0.0 : -(own0 * inv), m1 = isp1 ? 0.0 : -(own1 * inv);
#pragma unroll
                for (int c = k + 1; c < 4; c++) {
                    const double pv = __shfl_sync(kFullMask, s1 ? pe[1][c] : pe[0][c], src);
                    pe[0][c] = fma(m0, pv, pe[0][c]);
                    pe[1][c] = fma(m1, pv, pe[1][c]);
                }
#pragma unroll
                for (int j = 0; j < k; j++) {
                    const double gv = __shfl_sync(kFullMask, s1 ? gc[1][j] : gc[0][j], src);
                    gc[0][j] = fma(m0, gv, gc[0][j]);
                    gc[1][j] = fma(m1, gv, gc[1][j]);
                }
                gc[0][k] = m0;
                gc[1][k] = m1;
                if (isp0) { elig[0] = false; myq[0] = q0 + k; myinv[0] = inv0; }
                if (isp1) { elig[1] = false; myq[1] = q0 + k; myinv[1] = inv1; }
            }
            R32_STAMP();
            {
                const double r0 = rhs[0], r1 = rhs[1];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double rv = __shfl_sync(kFullMask, pr[j] >= 32 ? r1 : r0, pr[j] & 31);
                    rhs[0] = fma(gc[0][j], rv, rhs[0]);
                    rhs[1] = fma(gc[1][j], rv, rhs[1]);
                }
            }
            // (C) the coefficients of this warp's own rows (its private area)
            if (wown) {
#pragma unroll
                for (int j = 0; j < 4; j++) sts_f64(sg + 8u * (unsigned)(j * MA::CSG + (lane & 15)), wslot ? gc[1][j] : gc[0][j]);
            }
            // (D) the pivot rows this warp owns, as they were at panel start, into the block-wide lines
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
            R32_STAMP();
            __syncthreads();
            R32_STAMP();
            // (E) fragments and the rank-4 update of the live tiles, every warp on its own two tile rows
            double a[MA::TRL];
#pragma unroll
            for (int trl = 0; trl < MA::TRL; trl++) a[trl] = lds_f64(sg + 8u * (unsigned)(t * MA::CSG + 8 * trl + g));
#pragma unroll
            for (int j = jlo; j < MA::TJ; j++) {
                const double b = lds_f64(sa + 8u * (unsigned)(MA::oP + t * MA::PS + 8 * j + g));
#pragma unroll
                for (int trl = 0; trl < MA::TRL; trl++) dmma_8x8x4(w[trl][j], a[trl], b);
            }
            // (Wp is rewritten behind barrier 2 of this panel, Gs / P behind barrier 1 of the next one)
        }
        R32_STAMP();
        // ---- factors of the eliminated node; survivors shift E <- B, B <- 0 ----------------------------
        const int c = nodes[jrel];
        double* TLc = TL + c * nn;
        double* TRc = TR + c * nn;
#pragma unroll
        for (int trl = 0; trl < MA::TRL; trl++) {
            const int hl = (16 * warp + 8 * trl + g) & 31;  // the lane that holds this row (slot wslot)
            const int q = __shfl_sync(kFullMask, wslot ? myq[1] : myq[0], hl);
            const double inv = __shfl_sync(kFullMask, wslot ? myinv[1] : myinv[0], hl);
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
        if (warp == 0) {
            if (myq[0] >= 0) rt[(size_t)c * n + myq[0]] = rhs[0] * myinv[0];
            if (myq[1] >= 0) rt[(size_t)c * n + myq[1]] = rhs[1] * myinv[1];
        }
        carried = ~((unsigned long long)__ballot_sync(kFullMask, myq[0] >= 0) |
                    ((unsigned long long)__ballot_sync(kFullMask, myq[1] >= 0) << 32));
    }
    if (bad && tid == 0) atomicExch(status, 1);  // block-uniform: every warp computed the same keys
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
        if (warp == 0) {
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int row = lane + 32 * sl;
                if ((carried >> row) & 1ull) outr[(size_t)grp * n + __popcll(carried & ((1ull << row) - 1ull))] = rhs[sl];
            }
        }
    }
}


}  // namespace mirk
