// abd_tree_exp.cuh — EXPERIMENT (not in the library): flag-driven tree kernels over the team merge of csrc/abd_team.cuh.
// Measured slower than the segment kernels (profiles/r02/notes.md); kept for the record with its harness exp_team.cu.
#pragma once
#include "abd_team.cuh"

namespace mirk {
// ---- the reduction TREE above level 0 in two launches (two-point problems: pure radix-2 pairing, one root) ----------
// k_tree_up16    one team per group of the tree's first level.  A team merges its group, publishes the collapsed relation
//                and arrives at the parent group's counter; the LAST arriver of a parent continues with the parent's merge
//                (the relations of its sibling were written by another CTA: read past L1), every other team exits — no
//                spinning, no co-residency requirement, no launch or grid barrier between levels.  The team that merges
//                the root also runs the closing solve on the two kept nodes + the boundary rows.
// k_tree_down16  back substitution of the tree, one warp per group of the first level.  Warp b owns the merges
//                (l, b >> l) for l <= ctz(b) — the left spine of the sub-tree it is the leftmost leaf of — and walks them
//                top down with the running solution in registers; the right child of a merge waits for the merge's flag
//                (one waiter per flag, which resets it).  Launched cooperatively: all warps are co-resident by contract.
#if defined(MIRK_TREE_PROF)  // experiments: latest start time (globaltimer, ns) of a merge of each tree level
__device__ unsigned long long g_tree_prof[4 * (kMaxTail + 2)];
#endif
struct TreeSync {
    unsigned* cnt;           // arrival counters, one per group of levels >= 1 (self-resetting)
    unsigned* flag;          // completion flags of the down sweep, same indexing (reset by their one waiter)
    int off[kMaxTail + 1];   // offset of level l
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NW>
__global__ void __launch_bounds__(32 * NW)
k_tree_up16(const TailArgs a, const TreeSync ts) {
    extern __shared__ double tail_smem[];
    using TA = TeamABD16<NW>;
    __shared__ __align__(16) double sm[TA::smem_doubles];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    int g = blockIdx.x;
    for (int l = 0;; l++) {
#if defined(MIRK_TREE_PROF)
#define TREE_STAMP(ph) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); atomicMax(&g_tree_prof[4 * l + (ph)], t_); } } while (0)
#else
#define TREE_STAMP(ph) do { } while (0)
#endif
        TREE_STAMP(0);
        // a singular block is reported and the protocol goes on (garbage flows up): the counters stay consistent
        if (!team_reduce_group16<NW, true>(g, a.inL[l], a.inR[l], a.inr[l], a.outL[l], a.outR[l], a.outr[l], a.nodes[l], a.gs[l],
                                           a.TL, a.TR, a.rt, sm, lane, wv, 1))
            if (threadIdx.x == 0) atomicExch(a.status, 1);
        TREE_STAMP(1);
        if (l + 1 == a.nlev) break;  // that was the root
        __threadfence();             // this thread's part of the relation is visible device-wide ...
        TA::sync(1);                 // ... and so is every other thread's, before the arrival is counted
        TREE_STAMP(2);
        if (threadIdx.x == 0) {
            const int p = g >> 1;
            const int nchild = a.gs[l + 1][p + 1] - a.gs[l + 1][p];
            unsigned* c = ts.cnt + ts.off[l + 1] + p;
            const unsigned old = atomicAdd(c, 1u);
            const int last = old + 1u == (unsigned)nchild;
            if (last) *c = 0u;  // nobody touches it again in this solve
            s_last = last;     // (the sibling's relation is read past L1 — ld.global.cg — after this barrier)
        }
        TA::sync(1);
        TREE_STAMP(3);
        if (!s_last) return;
        g >>= 1;
    }
    __syncthreads();
    if (a.Q > 0)
    final_solve_body(16, a.Q, a.kept, a.relL, a.relR, a.relr, a.L, a.La, a.m_ptr, a.bc_nodes, a.Bc, a.resid, a.tail_off, a.M, a.delta,
                     a.status, tail_smem);
}

__global__ void __launch_bounds__(32)
k_tree_down16(const TailArgs a, const TreeSync ts) {
    constexpr int n = 16;
    constexpr size_t nn = (size_t)n * n;
    const int lane = threadIdx.x, b = blockIdx.x;
    const int half = lane >> 4, q = lane & 15;
    int lt = a.nlev - 1;
    if (b != 0 && __ffs(b) - 1 < lt) lt = __ffs(b) - 1;
    // factor row q of (half ? TR : TL) and rt of the eliminated node of merge (l, b >> l); has = the group is a pair
    double rv[n], rtv = 0.0;
    int c = 0, na = 0, nb = 0;
    bool has = false;
    auto fetch = [&](int l) {
        const int g = b >> l;
        const int k0 = a.gs[l][g], k1 = a.gs[l][g + 1];
        has = k1 - k0 == 2;
        na = a.nodes[l][k0];
        nb = a.nodes[l][k1];
        if (has) {
            c = a.nodes[l][k0 + 1];
            const double* row = (half ? a.TR : a.TL) + c * nn + (size_t)q * n;
#pragma unroll
            for (int k = 0; k < n; k += 2) {
                const double2 v = ldcg_v2f64(row + k);
                rv[k] = v.x; rv[k + 1] = v.y;
            }
            rtv = ldcg_f64(a.rt + (size_t)c * n + q);
        }
    };
    fetch(lt);
    if (lt + 1 < a.nlev) {  // the right child of merge (lt + 1, b >> (lt + 1)): wait for its solution
        unsigned* f = ts.flag + ts.off[lt + 1] + (b >> (lt + 1));
        if (lane == 0) {
            while (*(volatile unsigned*)f == 0u) { }
            *(volatile unsigned*)f = 0u;
            __threadfence();
        }
        __syncwarp();
    }
    // x = element q of the solution at (half ? nb : na)
    double x = ldcg_f64(a.delta + (size_t)(half ? nb : na) * n + q);
    for (int l = lt; l >= 0; l--) {
        double d16 = 0.0;
        const int cc = c;
        const bool had = has;
        if (has) {
            double p4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < n; k++) p4[k & 3] = fma(rv[k], __shfl_sync(kFullMask, x, 16 * half + k), p4[k & 3]);
            double acc = (p4[0] + p4[1]) + (p4[2] + p4[3]);
            acc += __shfl_xor_sync(kFullMask, acc, 16);
            d16 = rtv - acc;  // element q of the eliminated node's solution, in both halves
            if (half == 0) a.delta[(size_t)c * n + q] = d16;
        }
        if (l >= 1) {
            const int g = b >> l;
            if (2 * g + 1 < a.G[l - 1]) {  // a right child waits for this merge
                __threadfence();
                __syncwarp();
                if (lane == 0) *(volatile unsigned*)(ts.flag + ts.off[l] + g) = 1u;
            }
            // descend into the left child (l - 1, 2g): its ends are (na, c) when this merge was a pair, else the same ends
            fetch(l - 1);
            if (had) {
                if (nb == cc && half == 1) x = d16;
                if (na == cc && half == 0) x = d16;
            }
        }
    }
}

}  // namespace mirk
