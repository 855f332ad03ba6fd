// abd_team.cuh — the n = 16 merge of the ABD reduction by a TEAM of NW warps (NW = 1, 2, 4), DMMA fragment layout, and the
// upper levels of the reduction tree as THREAD-BLOCK CLUSTERS (k_seg_cluster16).
//
// Same algorithm, pivots, relation and factor formats as abd_mma.cuh (one warp per merge).  There a merge is a serial
// stream of ~1 650 warp instructions that one warp retires at one per ~5 cycles (4.1 us, profiles/r01_notes.md): the
// upper levels of the reduction tree, 11 merges deep above level 0, are pure merge latency.  Here the 32 x 48 matrix
// [E | A | B] is split by TILE ROWS: warp wv of a team owns tile rows [wv*TRW, (wv+1)*TRW) (TRW = 4 / NW; 8 rows x 48
// columns = 12 doubles per lane at NW = 4), so the panel gather, the publication of the pivot rows, the rank-4 DMMA
// update and the factor write-out are NW-way parallel.  The 4-pivot panel factorisation is the serial part: every
// warp of the team runs it REDUNDANTLY on the gathered 32 x 4 panel (lane r = row r, one REDUX per pivot, shuffles for
// the pivot lane's entries) — identical arithmetic in every warp, so no decision crosses warps and a panel costs two
// team barriers (named barriers, bar.sync id, 32*NW):
//     (A) every warp scatters its rows of the panel into the team's shared Wp[c][row]             -> barrier 1
//     (B) every warp: 4 pivot steps on lane-per-row copies; coefficients gc[4], pivot rows pr[4], rhs
//     (C) coefficients into the warp's private Gs[j][row]; (D) owners publish the 4 pivot rows P   -> barrier 2
//     (E) W[:, live columns] += G (rows of this warp x 4) * P (4 x 48): TRW DMMA per live tile column
// The per-row bookkeeping (rhs, pivot column myq, 1/pivot) is lane-per-row and replicated in every warp.
#pragma once
#include <cooperative_groups.h>

#include "abd_warp.cuh"

namespace mirk {

#if defined(MIRK_TEAM_PROF)  // experiments/exp_team.cu: cycle stamps of the phases of a merge (lane 0 of warp 0 of team 0)
__device__ long long g_team_prof[64];
__device__ int g_team_prof_n;
#define TEAM_STAMP() do { if (lane == 0 && wv == 0 && blockIdx.x == 0 && bar == 1 && g_team_prof_n < 64) g_team_prof[g_team_prof_n++] = clock64(); } while (0)
#else
#define TEAM_STAMP() do { } while (0)
#endif

__device__ __forceinline__ double2 ldcg_v2f64(const double* p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ldcg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <int NW> struct TeamABD16 {
    static_assert(NW == 1 || NW == 2 || NW == 4, "team of 1, 2 or 4 warps");
    static constexpr int n = 16, TRW = 4 / NW, TJ = 6;
    static constexpr int PS = 52;  // doubles per published pivot row (48 + pad: conflict-free B-fragment loads)
    static constexpr int CS = 36;  // column stride of the gathered panel / the coefficients (conflict-free)
    // shared memory of one team (doubles): Wp[4][CS] | Gs[NW][4][CS] | P[4][PS]
    static constexpr int oG = 4 * CS, oP = oG + NW * 4 * CS, smem_doubles = oP + 4 * PS;

    // barrier of the team: named barrier `bar` (1..15) over 32*NW threads; a warp barrier for a one-warp team
    __device__ __forceinline__ static void sync(int bar) {
        if (NW == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(32 * NW) : "memory");
    }

    // Gauss-Jordan on the 16 E columns (see the file header).  w: this lane's fragments of the warp's tile rows;
    // rhs / myq / myinv: row `lane` (replicated in every warp).  Returns false (team-uniform) on a zero / non-finite pivot.
    __device__ __forceinline__ static bool eliminate(double (&w)[TRW][TJ][2], double& rhs, int lane, int wv, int bar, double* sm,
                                                     int& myq, double& myinv) {
        const int g = lane >> 2, t = lane & 3;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
        const unsigned sg = sa + 8u * (unsigned)(oG + wv * 4 * CS);
        myq = -1;
        myinv = 0.0;
        bool elig = true;
#pragma unroll
        for (int pn = 0; pn < 4; pn++) {
            const int q0 = 4 * pn, jp = q0 >> 3, cq = q0 & 7, t0 = cq >> 1;
            // (A) this warp's rows of the panel into lane-per-row form
            TEAM_STAMP();
            if (t == t0 || t == t0 + 1) {
#pragma unroll
                for (int trl = 0; trl < TRW; trl++) {
                    const int row = 8 * (wv * TRW + trl) + g;
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0)) * CS + row), w[trl][jp][0]);
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0) + 1) * CS + row), w[trl][jp][1]);
                }
            }
            sync(bar);
            TEAM_STAMP();
            double pe[4];
#pragma unroll
            for (int c = 0; c < 4; c++) pe[c] = lds_f64(sa + 8u * (unsigned)(c * CS + lane));
            // (B) 4 pivot steps; gc[j] = coefficient of (pivot row j at panel start) in this row
            double gc[4] = {0.0, 0.0, 0.0, 0.0};
            int pr[4];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double own = pe[k];
                const double own_inv = fast_rcp(own);
                const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~31u) | (unsigned)(31 - lane)) : 0u;
                const unsigned mx = __reduce_max_sync(kFullMask, key);
                bad |= (mx >> 5) == 0u || mx >= 0x7ff00000u;
                pr[k] = 31 - (int)(mx & 31u);
                const bool isp = lane == pr[k];
                const double inv = __shfl_sync(kFullMask, own_inv, pr[k]);
                const double m = isp ? 0.0 : -(own * inv);
#pragma unroll
                for (int c = k + 1; c < 4; c++) pe[c] = fma(m, __shfl_sync(kFullMask, pe[c], pr[k]), pe[c]);
#pragma unroll
                for (int j = 0; j < k; j++) gc[j] = fma(m, __shfl_sync(kFullMask, gc[j], pr[k]), gc[j]);
                gc[k] = m;
                if (isp) { elig = false; myq = q0 + k; myinv = own_inv; }
            }
            if (bad) return false;  // identical in every warp of the team
            TEAM_STAMP();
            {
                const double rhs0 = rhs;
#pragma unroll
                for (int j = 0; j < 4; j++) rhs = fma(gc[j], __shfl_sync(kFullMask, rhs0, pr[j]), rhs);
            }
            // (C) the coefficients of this warp's rows (its private area)
            if ((lane >> 3) / TRW == wv) {
#pragma unroll
                for (int j = 0; j < 4; j++) sts_f64(sg + 8u * (unsigned)(j * CS + lane), gc[j]);
            }
            // (D) the pivot rows this warp owns, as they are (= as they were at panel start), into the team's lines
            const int jlo = cq == 0 ? jp : jp + 1;  // first tile column with live entries behind the panel
#pragma unroll
            for (int trl = 0; trl < TRW; trl++) {
                const int r = 8 * (wv * TRW + trl) + g;
                const int kk = r == pr[0] ? 0 : r == pr[1] ? 1 : r == pr[2] ? 2 : r == pr[3] ? 3 : -1;
                if (kk >= 0) {
                    const unsigned line = sa + 8u * (unsigned)(oP + kk * PS + 2 * t);
#pragma unroll
                    for (int j = jlo; j < TJ; j++) sts_v2f64(line + 8u * (unsigned)(8 * j), w[trl][j][0], w[trl][j][1]);
                }
            }
            TEAM_STAMP();
            sync(bar);
            TEAM_STAMP();
            // (E) fragments and the rank-4 update of the live tiles
            double a[TRW];
#pragma unroll
            for (int trl = 0; trl < TRW; trl++) a[trl] = lds_f64(sg + 8u * (unsigned)(t * CS + 8 * (wv * TRW + trl) + g));
#pragma unroll
            for (int j = jlo; j < TJ; j++) {
                const double b = lds_f64(sa + 8u * (unsigned)(oP + t * PS + 8 * j + g));
#pragma unroll
                for (int trl = 0; trl < TRW; trl++) dmma_8x8x4(w[trl][j], a[trl], b);
            }
            // Wp is rewritten after every warp passed barrier 2 (its reads precede it); Gs (private) and P are
            // rewritten behind the next panel's barrier 1, which every warp reaches after these loads
        }
        return true;
    }

    // carried rows 0..15 of a fresh group:  [E | A | B | rhs] = [R | L | 0 | r]
    template <bool CG>
    __device__ __forceinline__ static void load_carried(double (&w)[TRW][TJ][2], double& rhs, const double* Lk, const double* Rk,
                                                        const double* rk, int lane, int wv) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int trl = 0; trl < TRW; trl++) {
#pragma unroll
            for (int j = 0; j < TJ; j++) { w[trl][j][0] = 0.0; w[trl][j][1] = 0.0; }
            const int r = 8 * (wv * TRW + trl) + g;
            if (r < n) {
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    const double2 e = CG ? ldcg_v2f64(Rk + r * n + 8 * jj + 2 * t) : *reinterpret_cast<const double2*>(Rk + r * n + 8 * jj + 2 * t);
                    const double2 a = CG ? ldcg_v2f64(Lk + r * n + 8 * jj + 2 * t) : *reinterpret_cast<const double2*>(Lk + r * n + 8 * jj + 2 * t);
                    w[trl][jj][0] = e.x; w[trl][jj][1] = e.y;
                    w[trl][2 + jj][0] = a.x; w[trl][2 + jj][1] = a.y;
                }
            }
        }
        rhs = lane < n ? (CG ? ldcg_f64(rk + lane) : rk[lane]) : 0.0;
    }
    // incoming relation into the free row slots, in row order:  [E | A | B | rhs] = [L | 0 | R | r]
    template <bool CG>
    __device__ __forceinline__ static void load_incoming(double (&w)[TRW][TJ][2], double& rhs, unsigned freem, const double* Lk,
                                                         const double* Rk, const double* rk, int lane, int wv) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int trl = 0; trl < TRW; trl++) {
            const int r = 8 * (wv * TRW + trl) + g;
            if ((freem >> r) & 1u) {
                const int idx = __popc(freem & ((1u << r) - 1u));
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    const double2 e = CG ? ldcg_v2f64(Lk + idx * n + 8 * jj + 2 * t) : *reinterpret_cast<const double2*>(Lk + idx * n + 8 * jj + 2 * t);
                    const double2 b = CG ? ldcg_v2f64(Rk + idx * n + 8 * jj + 2 * t) : *reinterpret_cast<const double2*>(Rk + idx * n + 8 * jj + 2 * t);
                    w[trl][jj][0] = e.x; w[trl][jj][1] = e.y;
                    w[trl][2 + jj][0] = 0.0; w[trl][2 + jj][1] = 0.0;
                    w[trl][4 + jj][0] = b.x; w[trl][4 + jj][1] = b.y;
                }
            }
        }
        if ((freem >> lane) & 1u) {
            const int idx = __popc(freem & ((1u << lane) - 1u));
            rhs = CG ? ldcg_f64(rk + idx) : rk[idx];
        }
    }
    // factors of the eliminated node c (d_c = rt - TL d_a - TR d_right); survivors shift E <- B, B <- 0.
    // Returns the new carried mask.
    __device__ __forceinline__ static unsigned store_factors_and_shift(double (&w)[TRW][TJ][2], double rhs, int myq, double myinv,
                                                                       double* TLc, double* TRc, double* rtc, int lane, int wv) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int trl = 0; trl < TRW; trl++) {
            const int r = 8 * (wv * TRW + trl) + g;
            const int q = __shfl_sync(kFullMask, myq, r);
            const double inv = __shfl_sync(kFullMask, myinv, r);
            if (q >= 0) {
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    *reinterpret_cast<double2*>(TLc + q * n + 8 * jj + 2 * t) = make_double2(w[trl][2 + jj][0] * inv, w[trl][2 + jj][1] * inv);
                    *reinterpret_cast<double2*>(TRc + q * n + 8 * jj + 2 * t) = make_double2(w[trl][4 + jj][0] * inv, w[trl][4 + jj][1] * inv);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    w[trl][jj][0] = w[trl][4 + jj][0]; w[trl][jj][1] = w[trl][4 + jj][1];
                    w[trl][4 + jj][0] = 0.0; w[trl][4 + jj][1] = 0.0;
                }
            }
        }
        if (wv == 0 && myq >= 0) rtc[myq] = rhs * myinv;
        return ~__ballot_sync(kFullMask, myq >= 0);
    }
    // the 16 carried rows, in row order, as a relation
    __device__ __forceinline__ static void store_relation(const double (&w)[TRW][TJ][2], double rhs, unsigned carried, double* oL,
                                                          double* oR, double* orr, int lane, int wv) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int trl = 0; trl < TRW; trl++) {
            const int r = 8 * (wv * TRW + trl) + g;
            if ((carried >> r) & 1u) {
                const int idx = __popc(carried & ((1u << r) - 1u));
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    *reinterpret_cast<double2*>(oR + idx * n + 8 * jj + 2 * t) = make_double2(w[trl][jj][0], w[trl][jj][1]);
                    *reinterpret_cast<double2*>(oL + idx * n + 8 * jj + 2 * t) = make_double2(w[trl][2 + jj][0], w[trl][2 + jj][1]);
                }
            }
        }
        if (wv == 0 && ((carried >> lane) & 1u)) orr[__popc(carried & ((1u << lane) - 1u))] = rhs;
    }
    // incoming relation from a staged copy in shared memory (rows [L | R | r | pad], WarpABD<16>::stage_stride doubles each)
    __device__ __forceinline__ static void load_incoming_staged(double (&w)[TRW][TJ][2], double& rhs, unsigned freem, const double* st,
                                                                int lane, int wv) {
        const int g = lane >> 2, t = lane & 3;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(st);
        constexpr int SS = WarpABD<16>::stage_stride;
#pragma unroll
        for (int trl = 0; trl < TRW; trl++) {
            const int r = 8 * (wv * TRW + trl) + g;
            if ((freem >> r) & 1u) {
                const int idx = __popc(freem & ((1u << r) - 1u));
                const unsigned row = sa + 8u * (unsigned)(idx * SS + 2 * t);
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    const double2 e = lds_v2f64(row + 8u * (unsigned)(8 * jj)), b = lds_v2f64(row + 8u * (unsigned)(n + 8 * jj));
                    w[trl][jj][0] = e.x; w[trl][jj][1] = e.y;
                    w[trl][2 + jj][0] = 0.0; w[trl][2 + jj][1] = 0.0;
                    w[trl][4 + jj][0] = b.x; w[trl][4 + jj][1] = b.y;
                }
            }
        }
        if ((freem >> lane) & 1u) rhs = lds_f64(sa + 8u * (unsigned)(__popc(freem & ((1u << lane) - 1u)) * SS + 2 * n));
    }
    // the 16 carried rows, in row order, into a staged copy (dst: generic pointer, possibly another CTA's shared memory):
    // the relation (L, R, r) = (A part, E part, rhs) of the carried rows
    __device__ __forceinline__ static void store_relation_staged(const double (&w)[TRW][TJ][2], double rhs, unsigned carried, double* dst,
                                                                 int lane, int wv) {
        const int g = lane >> 2, t = lane & 3;
        constexpr int SS = WarpABD<16>::stage_stride;
#pragma unroll
        for (int trl = 0; trl < TRW; trl++) {
            const int r = 8 * (wv * TRW + trl) + g;
            if ((carried >> r) & 1u) {
                double* row = dst + __popc(carried & ((1u << r) - 1u)) * SS + 2 * t;
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    *reinterpret_cast<double2*>(row + 8 * jj) = make_double2(w[trl][2 + jj][0], w[trl][2 + jj][1]);
                    *reinterpret_cast<double2*>(row + n + 8 * jj) = make_double2(w[trl][jj][0], w[trl][jj][1]);
                }
            }
        }
        if (wv == 0 && ((carried >> lane) & 1u)) dst[__popc(carried & ((1u << lane) - 1u)) * SS + 2 * n] = rhs;
    }
};

// One group of one reduction level by a team (arguments as warp_reduce_group; CG: read the relations past L1 —
// they were written by other CTAs of the same launch).  Returns false (team-uniform) on a singular block.
template <int NW, bool CG>
__device__ __forceinline__ bool team_reduce_group16(int grp, const double* inL, const double* inR, const double* inr, double* outL,
                                                    double* outR, double* outr, const int* nodes, const int* gs, double* TL,
                                                    double* TR, double* rt, double* sm, int lane, int wv, int bar) {
    using TA = TeamABD16<NW>;
    constexpr int n = 16;
    constexpr size_t nn = (size_t)n * n;
    const int k0 = gs[grp], k1 = gs[grp + 1];
    double w[TA::TRW][TA::TJ][2];
    double rhs;
    TA::template load_carried<CG>(w, rhs, inL + k0 * nn, inR + k0 * nn, inr + (size_t)k0 * n, lane, wv);
    unsigned carried = 0x0000ffffu;
    for (int j = k0 + 1; j < k1; j++) {
        TA::template load_incoming<CG>(w, rhs, ~carried, inL + j * nn, inR + j * nn, inr + (size_t)j * n, lane, wv);
        int myq;
        double myinv;
        if (!TA::eliminate(w, rhs, lane, wv, bar, sm, myq, myinv)) return false;
        TEAM_STAMP();
        const int c = nodes[j];
        carried = TA::store_factors_and_shift(w, rhs, myq, myinv, TL + c * nn, TR + c * nn, rt + (size_t)c * n, lane, wv);
        TEAM_STAMP();
    }
    TA::store_relation(w, rhs, carried, outL + grp * nn, outR + grp * nn, outr + (size_t)grp * n, lane, wv);
    return true;
}

// One level of the reduction, one team per group, TPB teams per CTA.
template <int NW, int TPB>
__global__ void __launch_bounds__(32 * NW * TPB)
k_reduce_team16(int G, const double* __restrict__ inL, const double* __restrict__ inR, const double* __restrict__ inr,
                double* __restrict__ outL, double* __restrict__ outR, double* __restrict__ outr, const int* __restrict__ nodes,
                const int* __restrict__ gs, double* __restrict__ TL, double* __restrict__ TR, double* __restrict__ rt,
                int* __restrict__ status) {
    __shared__ __align__(16) double sm[TPB][TeamABD16<NW>::smem_doubles];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, team = warp / NW, wv = warp % NW;
    const int g = blockIdx.x * TPB + team;
    if (g >= G) return;
    if (!team_reduce_group16<NW, false>(g, inL, inR, inr, outL, outR, outr, nodes, gs, TL, TR, rt, sm[team], lane, wv, 1 + team))
        if (lane == 0 && wv == 0) atomicExch(status, 1);
}

// ---- a SEGMENT of <= 4 radix-2 levels of the tree by thread-block clusters of 8 CTAs (one team of 4 warps each) -------
// Same level tables and block-to-sub-tree mapping as k_tail_warp with multi = 1 (TailArgs: cluster c owns groups
// [8c, 8c + 8) of the segment's first level), other execution: the 8 merges of a sub-tree's first level run on 8 DIFFERENT SMs instead of 8 warps of one (no contention for
// one SM's shared-memory pipe), every merge is the 5 300-cycle team merge instead of the 8 000-cycle one-warp merge, and a
// level boundary is ONE cluster barrier: the right partner writes its collapsed relation straight into the left partner's
// shared memory (DSMEM, staged format of the level-0 kernel) while the left partner's relation never leaves its registers.
// Factors of the eliminated nodes go to global memory as everywhere.  Requires pure pairing (gs[g] = 2g) inside the segment.
// Measured (B200, profiles/r02/notes.md): 3.0 us per level (2.4 merge + 0.6 cluster barrier) against 5.5 us in
// k_tail_warp; a cluster launch costs ~4 us more than a plain one and a segment of 834 groups (105 clusters, 840 CTAs)
// does not fit one wave of co-scheduled clusters, so the host uses this kernel for segments of <= 148 groups only
// (C2: the 105 -> 7 segment, 21.7 -> 15.9 us) and the one-SM kernels elsewhere (the one-block tail is dominated by its
// closing solve and back substitution: no gain there).
constexpr int kClusterCTAs = 8;
#if defined(MIRK_SEG_PROF)  // experiments: globaltimer stamps of cluster 0, rank 0
__device__ unsigned long long g_seg_prof[64];
#define SEG_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); g_seg_prof[i] = t_; } } while (0)
#else
#define SEG_STAMP(i) do { } while (0)
#endif

__global__ void __cluster_dims__(kClusterCTAs, 1, 1) __launch_bounds__(128)
k_seg_cluster16(const TailArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    using TA = TeamABD16<4>;
    constexpr int n = 16;
    constexpr size_t nn = (size_t)n * n;
    __shared__ __align__(16) double sm[TA::smem_doubles];
    __shared__ __align__(16) double stage[2][WarpABD<16>::stage_doubles];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const int rank = (int)cluster.block_rank(), cl = (int)blockIdx.x / kClusterCTAs;
    const int leaf = kClusterCTAs * cl + rank;  // this CTA's group of the segment's first level
    double w[TA::TRW][TA::TJ][2];
    double rhs = 0.0;
    unsigned carried = 0x0000ffffu;
    bool have = false, ok = true;
    SEG_STAMP(0);
    // A CTA's shared memory may only be written remotely once that CTA runs: every CTA arrives at the cluster barrier
    // now and waits for it behind its first-level merge (split barrier: the wait is free by then)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    {
        if (leaf < a.G[0]) {
            const int k0 = a.gs[0][leaf], k1 = a.gs[0][leaf + 1];
            TA::load_carried<false>(w, rhs, a.inL[0] + k0 * nn, a.inR[0] + k0 * nn, a.inr[0] + (size_t)k0 * n, lane, wv);
            for (int j = k0 + 1; j < k1; j++) {
                TA::load_incoming<false>(w, rhs, ~carried, a.inL[0] + j * nn, a.inR[0] + j * nn, a.inr[0] + (size_t)j * n, lane, wv);
                int myq;
                double myinv;
                ok = TA::eliminate(w, rhs, lane, wv, 1, sm, myq, myinv) && ok;
                const int c = a.nodes[0][j];
                carried = TA::store_factors_and_shift(w, rhs, myq, myinv, a.TL + c * nn, a.TR + c * nn, a.rt + (size_t)c * n, lane, wv);
            }
            have = true;
        }
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        SEG_STAMP(1);
        for (int t = 1; t < a.nlev; t++) {
            const int half = 1 << (t - 1);
            double* buf = stage[t & 1];
            if (have && (rank & (2 * half - 1)) == half) {  // right partner: hand the relation over and retire
                TA::store_relation_staged(w, rhs, carried, cluster.map_shared_rank(buf, rank - half), lane, wv);
                have = false;
            }
            cluster.sync();
            SEG_STAMP(2 * t);
            if (have && (rank & (2 * half - 1)) == 0) {
                const int g = leaf >> t;
                const int k0 = a.gs[t][g], k1 = a.gs[t][g + 1];
                if (k1 - k0 == 2) {  // (a lone relation at the end of a level passes through)
                    TA::load_incoming_staged(w, rhs, ~carried, buf, lane, wv);
                    int myq;
                    double myinv;
                    ok = TA::eliminate(w, rhs, lane, wv, 1, sm, myq, myinv) && ok;
                    const int c = a.nodes[t][k0 + 1];
                    carried = TA::store_factors_and_shift(w, rhs, myq, myinv, a.TL + c * nn, a.TR + c * nn, a.rt + (size_t)c * n, lane, wv);
                }
            }
            SEG_STAMP(2 * t + 1);
            // (the buffers alternate, so the next level's writer cannot overtake this level's reader: one barrier per level)
        }
        if (have) {
            const int g = leaf >> (a.nlev - 1), l = a.nlev - 1;
            TA::store_relation(w, rhs, carried, a.outL[l] + g * nn, a.outR[l] + g * nn, a.outr[l] + (size_t)g * n, lane, wv);
        }
        if (!ok && threadIdx.x == 0) atomicExch(a.status, 1);
    }
}

inline cudaError_t launch_seg_cluster16(cudaStream_t st, const TailArgs& a, int clusters) {
    k_seg_cluster16<<<clusters * kClusterCTAs, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace mirk
