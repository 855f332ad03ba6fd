// EXPERIMENT (not kept, profiles/r02/notes.md "look-ahead"): abd_mma.cuh with -DMIRK_MMA_LOOKAHEAD — the rank-4 products of a
// panel deferred into the next panel's pivot chain.  Bit-identical factors, parity tests green, 0.2503 -> 0.2497 ms per C2 step.
// abd_mma.cuh — the n = 16 merge of the ABD reduction on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64).
//
// Same algorithm, pivots, relation and factor formats as abd_warp.cuh (row-pivoted Gauss-Jordan on the n
// E-columns of the stacked 2n x (3n+1) matrix of one merge), other data layout.  With one lane per row every
// lane needs every element of every pivot row, so each pivot costs (3n+1) x 32 x 8 B through the 128 B/clk
// shared-memory/shuffle crossbar — that, not arithmetic, bounded the lane-per-row kernel (profiles/r01_notes.md).
// Here the 32 x 48 matrix [E | A | B] lives in the C-fragment layout of m8n8k4 (4 x 6 tiles of 8 x 8; lane
// (g = lane/4, t = lane%4) holds row 8*tr+g, columns 8*j+2t, 8*j+2t+1 of tile (tr, j)), pivots are taken in
// panels of 4 columns, and the trailing update of a panel is ONE rank-4 product per tile,
//     W[:, later columns] += G (32 x 4) * P (4 x 48),
// where P are the 4 pivot rows as they were at panel start and G the per-row combination coefficients the
// panel factorisation accumulates (see WarpABD::eliminate, MIRK_ELIM_PANEL).  The fragments cost one double
// per lane and tile row/column: 4 + 6 doubles per lane and panel instead of 4 x 45.
//
// Panel factorisation: the 32 x 4 panel is gathered (through shared memory) into lane-per-row form, lane r
// owning row r — pivot search is then one REDUX, and the pivot lane's <= 3 panel entries and coefficients travel
// by warp shuffle.  Lane r also owns rhs[r], the pivot bookkeeping of row r (myq, 1/pivot) and its eligibility.
// experiments/mma_merge_emul.py is a lane-by-lane numpy emulation of this file's data movement.
// Included by abd_warp.cuh (after WarpABD), never directly.
#pragma once

namespace mirk {

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void sts_f64(unsigned addr, double x) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory");
}
__device__ __forceinline__ void sts_v2f64(unsigned addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

struct MmaABD16 {
    static constexpr int n = 16, TR = 4, TJ = 6;  // 4 tile rows x 6 tile columns of 8 x 8
    static constexpr int PS = 52;                 // doubles per published pivot row (48 + pad: conflict-free B loads)
    // per-warp shared memory (doubles): [0, 4*CS) the gathered panel, column-major Wp[c][row] (column stride
    // CS = 36: every access below is bank-conflict free), later the coefficients Gs[j][row]; then the 4 published
    // pivot rows; then (optional) the cp.async staging buffer of WarpABD<16>
    static constexpr int CS = 36, oP = 4 * CS, BUF = 4 * CS + 4 * PS;
#if defined(MIRK_MMA_LOOKAHEAD)
    static constexpr int core_doubles = 2 * BUF;  // look-ahead: the buffers of two consecutive panels are live
#else
    static constexpr int core_doubles = BUF;
#endif
    static constexpr int stage_stride = WarpABD<16>::stage_stride;
    template <bool STAGE> __host__ __device__ static constexpr int smem_doubles() { return core_doubles + (STAGE ? WarpABD<16>::stage_doubles : 0); }

#if defined(MIRK_MMA_LOOKAHEAD)
    // Gauss-Jordan on the 16 E columns with LOOK-AHEAD (arguments and results as below; the same products on the same
    // operands, so the factors are bit-identical): after a panel's factorisation only the tile column that holds the
    // NEXT panel is updated at once; the other live tiles' rank-4 products are issued between the dependent steps of the
    // next panel's pivot chain (~150 cycles per pivot of pure latency), where the DMMA pipe would otherwise idle.
    // Their fragments are read from the previous panel's buffers, hence two buffer sets per warp.
    __device__ __forceinline__ static bool eliminate(double (&w)[TR][TJ][2], double& rhs, int lane, double* sm, int& myq,
                                                     double& myinv) {
        const int g = lane >> 2, t = lane & 3;
        const unsigned sa0 = (unsigned)__cvta_generic_to_shared(sm);
        myq = -1;
        myinv = 0.0;
        bool elig = true;
        double a[TR];  // A fragments of the last factorised panel
#pragma unroll
        for (int pn = 0; pn < 4; pn++) {
            const int q0 = 4 * pn, jp = q0 >> 3, cq = q0 & 7, t0 = cq >> 1;
            const unsigned sa = sa0 + 8u * (unsigned)((pn & 1) * BUF);        // this panel's buffers
            const unsigned sp = sa0 + 8u * (unsigned)(((pn & 1) ^ 1) * BUF);  // the previous panel's
            // live tiles of the previous panel still to be updated: [pjlo, TJ) without tile column jp (done at once)
            const int pjlo = pn == 0 ? TJ : (((4 * (pn - 1)) & 7) == 0 ? (4 * (pn - 1)) >> 3 : ((4 * (pn - 1)) >> 3) + 1);
            // (A) the panel into lane-per-row form
            if (t == t0 || t == t0 + 1) {
#pragma unroll
                for (int tr = 0; tr < TR; tr++) {
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0)) * CS + 8 * tr + g), w[tr][jp][0]);
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0) + 1) * CS + 8 * tr + g), w[tr][jp][1]);
                }
            }
            __syncwarp();
            double pe[4];
#pragma unroll
            for (int c = 0; c < 4; c++) pe[c] = lds_f64(sa + 8u * (unsigned)(c * CS + lane));
            // (B) 4 pivot steps on the panel, the deferred products of the previous panel in between
            double gc[4] = {0.0, 0.0, 0.0, 0.0};
            int pr[4];
            bool bad = false;
            const int nd = TJ - pjlo - (jp >= pjlo ? 1 : 0);  // deferred tile columns: 0 (first panel), 5 or 4
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double own = pe[k];
                const double own_inv = fast_rcp(own);
                const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~31u) | (unsigned)(31 - lane)) : 0u;
                const unsigned mx = __reduce_max_sync(kFullMask, key);
                // the deferred tile columns are spread over the 4 steps (5 of them: 2 + 1 + 1 + 1); every index below is a
                // constant after unrolling
#pragma unroll
                for (int j = 0; j < TJ; j++) {
                    const int idx = j - pjlo - ((jp >= pjlo && j > jp) ? 1 : 0);
                    const int step = nd == 5 ? (idx == 0 ? 0 : idx - 1) : idx;
                    if (pn > 0 && j >= pjlo && j != jp && step == k) {
                        const double b = lds_f64(sp + 8u * (unsigned)(oP + t * PS + 8 * j + g));
#pragma unroll
                        for (int tr = 0; tr < TR; tr++) dmma_8x8x4(w[tr][j], a[tr], b);
                    }
                }
                bad |= (mx >> 5) == 0u || mx >= 0x7ff00000u;  // zero / non-finite pivot: checked once per panel
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
            if (bad) return false;  // warp-uniform
            {
                const double rhs0 = rhs;
#pragma unroll
                for (int j = 0; j < 4; j++) rhs = fma(gc[j], __shfl_sync(kFullMask, rhs0, pr[j]), rhs);
            }
            // (C) coefficients to shared memory (over the gathered panel)
#pragma unroll
            for (int j = 0; j < 4; j++) sts_f64(sa + 8u * (unsigned)(j * CS + lane), gc[j]);
            // (D) the 4 pivot rows as they are (every product of the earlier panels has been applied by now)
            const int jlo = cq == 0 ? jp : jp + 1;
#pragma unroll
            for (int tr = 0; tr < TR; tr++) {
                const int r = 8 * tr + g;
                const int kk = r == pr[0] ? 0 : r == pr[1] ? 1 : r == pr[2] ? 2 : r == pr[3] ? 3 : -1;
                if (kk >= 0) {
                    const unsigned line = sa + 8u * (unsigned)(oP + kk * PS + 2 * t);
#pragma unroll
                    for (int j = jlo; j < TJ; j++) sts_v2f64(line + 8u * (unsigned)(8 * j), w[tr][j][0], w[tr][j][1]);
                }
            }
            __syncwarp();
            // (E) fragments; at once only the tile column of the next panel (last panel: every live tile)
#pragma unroll
            for (int tr = 0; tr < TR; tr++) a[tr] = lds_f64(sa + 8u * (unsigned)(t * CS + 8 * tr + g));
            const int jn = (4 * (pn + 1)) >> 3;
#pragma unroll
            for (int j = jlo; j < TJ; j++) {
                if (pn == 3 || j == jn) {
                    const double b = lds_f64(sa + 8u * (unsigned)(oP + t * PS + 8 * j + g));
#pragma unroll
                    for (int tr = 0; tr < TR; tr++) dmma_8x8x4(w[tr][j], a[tr], b);
                }
            }
            // (no barrier here: the next panel gathers into the other buffer set, whose last readers are two barriers back)
        }
        return true;
    }
#else
    // Gauss-Jordan on the 16 E columns.  w: this lane's fragments, rhs: rhs of row `lane`.  On return lane r
    // knows whether row r was a pivot row (myq = its column, myinv = 1 / pivot) or survives (myq = -1).
    __device__ __forceinline__ static bool eliminate(double (&w)[TR][TJ][2], double& rhs, int lane, double* sm, int& myq,
                                                     double& myinv) {
        const int g = lane >> 2, t = lane & 3;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
        myq = -1;
        myinv = 0.0;
        bool elig = true;
#pragma unroll
        for (int pn = 0; pn < 4; pn++) {
            const int q0 = 4 * pn, jp = q0 >> 3, cq = q0 & 7, t0 = cq >> 1;
            // (A) the panel into lane-per-row form
            if (t == t0 || t == t0 + 1) {
#pragma unroll
                for (int tr = 0; tr < TR; tr++) {
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0)) * CS + 8 * tr + g), w[tr][jp][0]);
                    sts_f64(sa + 8u * (unsigned)((2 * (t - t0) + 1) * CS + 8 * tr + g), w[tr][jp][1]);
                }
            }
            __syncwarp();
            double pe[4];
#pragma unroll
            for (int c = 0; c < 4; c++) pe[c] = lds_f64(sa + 8u * (unsigned)(c * CS + lane));
            // (B) 4 pivot steps on the panel; gc[j] = coefficient of (pivot row j at panel start) in this row
            double gc[4] = {0.0, 0.0, 0.0, 0.0};
            int pr[4];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double own = pe[k];
                const double own_inv = fast_rcp(own);
                const unsigned key = elig ? (((unsigned)__double2hiint(fabs(own)) & ~31u) | (unsigned)(31 - lane)) : 0u;
                const unsigned mx = __reduce_max_sync(kFullMask, key);
                bad |= (mx >> 5) == 0u || mx >= 0x7ff00000u;  // zero / non-finite pivot: checked once per panel
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
            if (bad) return false;  // warp-uniform
            {
                const double rhs0 = rhs;
#pragma unroll
                for (int j = 0; j < 4; j++) rhs = fma(gc[j], __shfl_sync(kFullMask, rhs0, pr[j]), rhs);
            }
            // (C) coefficients to shared memory (over the gathered panel: every lane has consumed it, the
            //     last REDUX needed pe[3])
#pragma unroll
            for (int j = 0; j < 4; j++) sts_f64(sa + 8u * (unsigned)(j * CS + lane), gc[j]);
            // (D) the 4 pivot rows as they are (= as they were at panel start) into shared lines
            const int jlo = cq == 0 ? jp : jp + 1;  // first tile column with live entries behind the panel
#pragma unroll
            for (int tr = 0; tr < TR; tr++) {
                const int r = 8 * tr + g;
                const int kk = r == pr[0] ? 0 : r == pr[1] ? 1 : r == pr[2] ? 2 : r == pr[3] ? 3 : -1;
                if (kk >= 0) {
                    const unsigned line = sa + 8u * (unsigned)(oP + kk * PS + 2 * t);
#pragma unroll
                    for (int j = jlo; j < TJ; j++) sts_v2f64(line + 8u * (unsigned)(8 * j), w[tr][j][0], w[tr][j][1]);
                }
            }
            __syncwarp();
            // (E) fragments and the rank-4 update of the live tiles
            double a[TR];
#pragma unroll
            for (int tr = 0; tr < TR; tr++) a[tr] = lds_f64(sa + 8u * (unsigned)(t * CS + 8 * tr + g));
#pragma unroll
            for (int j = jlo; j < TJ; j++) {
                const double b = lds_f64(sa + 8u * (unsigned)(oP + t * PS + 8 * j + g));
#pragma unroll
                for (int tr = 0; tr < TR; tr++) dmma_8x8x4(w[tr][j], a[tr], b);
            }
            // (the next panel's gather is separated from these loads by its own __syncwarp only for Wp/Gs
            //  reuse: Gs is re-read above before any lane can pass the next panel's first barrier)
            __syncwarp();
        }
        return true;
    }
#endif
};

// One group of one reduction level by one warp, n = 16 (arguments as warp_reduce_group).
template <bool STAGE>
__device__ __forceinline__ bool mma_reduce_group16(int grp, const double* inL, const double* inR, const double* inr,
                                                   double* outL, double* outR, double* outr, const int* nodes,
                                                   const int* gs, double* TL, double* TR, double* rt, double* sm,
                                                   int lane) {
    using MA = MmaABD16;
    using WA = WarpABD<16>;
    constexpr int n = 16;
    constexpr size_t nn = (size_t)n * n;
    const int g = lane >> 2, t = lane & 3;
    const int k0 = gs[grp], k1 = gs[grp + 1];
    double* stage = sm + MA::core_doubles;
    if (STAGE && k0 + 1 < k1) WA::stage_issue(stage, inL + (k0 + 1) * nn, inR + (k0 + 1) * nn, inr + (size_t)(k0 + 1) * n, lane);

    double w[MA::TR][MA::TJ][2];
    double rhs = 0.0;
    // carried rows 0..15:  [E | A | B | rhs] = [R | L | 0 | r]
    {
        const double* Lk = inL + k0 * nn;
        const double* Rk = inR + k0 * nn;
#pragma unroll
        for (int tr = 0; tr < MA::TR; tr++) {
#pragma unroll
            for (int j = 0; j < MA::TJ; j++) { w[tr][j][0] = 0.0; w[tr][j][1] = 0.0; }
            if (tr < 2) {
                const int r = 8 * tr + g;
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    const double2 e = *reinterpret_cast<const double2*>(Rk + r * n + 8 * jj + 2 * t);
                    const double2 a = *reinterpret_cast<const double2*>(Lk + r * n + 8 * jj + 2 * t);
                    w[tr][jj][0] = e.x; w[tr][jj][1] = e.y;
                    w[tr][2 + jj][0] = a.x; w[tr][2 + jj][1] = a.y;
                }
            }
        }
        if (lane < n) rhs = inr[(size_t)k0 * n + lane];
    }
    unsigned carried = 0x0000ffffu;
    for (int j = k0 + 1; j < k1; j++) {
        const unsigned freem = ~carried;
        // incoming rows into the free row slots, in row order:  [E | A | B | rhs] = [L | 0 | R | r]
        if (STAGE) {
            WA::stage_wait();
            const unsigned sa = (unsigned)__cvta_generic_to_shared(stage);
#pragma unroll
            for (int tr = 0; tr < MA::TR; tr++) {
                const int r = 8 * tr + g;
                if ((freem >> r) & 1u) {
                    const int idx = __popc(freem & ((1u << r) - 1u));
                    const unsigned row = sa + 8u * (unsigned)(idx * MA::stage_stride + 2 * t);
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        const double2 e = lds_v2f64(row + 8u * (unsigned)(8 * jj)), b = lds_v2f64(row + 8u * (unsigned)(n + 8 * jj));
                        w[tr][jj][0] = e.x; w[tr][jj][1] = e.y;
                        w[tr][2 + jj][0] = 0.0; w[tr][2 + jj][1] = 0.0;
                        w[tr][4 + jj][0] = b.x; w[tr][4 + jj][1] = b.y;
                    }
                }
            }
            if ((freem >> lane) & 1u) rhs = lds_f64(sa + 8u * (unsigned)(__popc(freem & ((1u << lane) - 1u)) * MA::stage_stride + 2 * n));
            __syncwarp();  // every lane has its rows before the buffer is refilled
            if (j + 1 < k1) WA::stage_issue(stage, inL + (j + 1) * nn, inR + (j + 1) * nn, inr + (size_t)(j + 1) * n, lane);
        } else {
            const double* Lk = inL + j * nn;
            const double* Rk = inR + j * nn;
#pragma unroll
            for (int tr = 0; tr < MA::TR; tr++) {
                const int r = 8 * tr + g;
                if ((freem >> r) & 1u) {
                    const int idx = __popc(freem & ((1u << r) - 1u));
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        const double2 e = *reinterpret_cast<const double2*>(Lk + idx * n + 8 * jj + 2 * t);
                        const double2 b = *reinterpret_cast<const double2*>(Rk + idx * n + 8 * jj + 2 * t);
                        w[tr][jj][0] = e.x; w[tr][jj][1] = e.y;
                        w[tr][2 + jj][0] = 0.0; w[tr][2 + jj][1] = 0.0;
                        w[tr][4 + jj][0] = b.x; w[tr][4 + jj][1] = b.y;
                    }
                }
            }
            if ((freem >> lane) & 1u) rhs = inr[(size_t)j * n + __popc(freem & ((1u << lane) - 1u))];
        }
        int myq;
        double myinv;
        if (!MA::eliminate(w, rhs, lane, sm, myq, myinv)) return false;
        // factors of the eliminated node c:  d_c = rt - TL d_a - TR d_right ;  survivors shift E <- B, B <- 0
        const int c = nodes[j];
        double* TLc = TL + c * nn;
        double* TRc = TR + c * nn;
#pragma unroll
        for (int tr = 0; tr < MA::TR; tr++) {
            const int r = 8 * tr + g;
            const int q = __shfl_sync(kFullMask, myq, r);
            const double inv = __shfl_sync(kFullMask, myinv, r);
            if (q >= 0) {
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    *reinterpret_cast<double2*>(TLc + q * n + 8 * jj + 2 * t) = make_double2(w[tr][2 + jj][0] * inv, w[tr][2 + jj][1] * inv);
                    *reinterpret_cast<double2*>(TRc + q * n + 8 * jj + 2 * t) = make_double2(w[tr][4 + jj][0] * inv, w[tr][4 + jj][1] * inv);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    w[tr][jj][0] = w[tr][4 + jj][0]; w[tr][jj][1] = w[tr][4 + jj][1];
                    w[tr][4 + jj][0] = 0.0; w[tr][4 + jj][1] = 0.0;
                }
            }
        }
        if (myq >= 0) rt[(size_t)c * n + myq] = rhs * myinv;
        carried = ~__ballot_sync(kFullMask, myq >= 0);
    }
    // the 16 carried rows, in row order, as the collapsed relation of the group
    {
        double* oL = outL + grp * nn;
        double* oR = outR + grp * nn;
#pragma unroll
        for (int tr = 0; tr < MA::TR; tr++) {
            const int r = 8 * tr + g;
            if ((carried >> r) & 1u) {
                const int idx = __popc(carried & ((1u << r) - 1u));
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    *reinterpret_cast<double2*>(oR + idx * n + 8 * jj + 2 * t) = make_double2(w[tr][jj][0], w[tr][jj][1]);
                    *reinterpret_cast<double2*>(oL + idx * n + 8 * jj + 2 * t) = make_double2(w[tr][2 + jj][0], w[tr][2 + jj][1]);
                }
            }
        }
        if ((carried >> lane) & 1u) outr[(size_t)grp * n + __popc(carried & ((1u << lane) - 1u))] = rhs;
    }
    return true;
}

}  // namespace mirk
