"""Numerical check (numpy, CPU) of the FRACTION-FREE panel chain proposed in DESIGN.md §7 for the n = 16 merge.

The shipped merge (csrc/abd_mma.cuh) factorises a panel of 4 pivot columns on lane-per-row copies with true multipliers
    m = -own * (1 / p),  pe[c] <- pe[c] + m * rp[c],  gc[j] <- gc[j] + m * gc_r[j],  gc[k] = m
(one reciprocal on the serial chain of every pivot) and then applies  W[:, behind the panel] += G P  with the 4 pivot rows P as
they were at panel start.  The proposal keeps the reciprocal off the chain: inside the panel every non-pivot row is updated as
    pe[c] <- p * pe[c] - own * rp[c],   gc'[j] <- p * gc'[j] - own * gc'_r[j],   gc'[k] = -own * S      (S = product of the pivots so far)
so all rows still eligible carry the SAME scale S (the arg-max of the next pivot search is unchanged), and at panel end
    gc = gc' * f,   f = prod_k (1 / p_k)  for rows that did not pivot,   f * p_q for the row that pivoted at step q,
turns the coefficients back into true multipliers — the trailing update and everything behind it see unscaled rows.

This script runs both variants on the stacked 32 x 48 matrices of random ABD-like merges (and badly scaled ones) and reports
whether the pivot rows agree and how far the results differ.  It proves nothing about speed; the cycle estimate is in DESIGN.md.

    python experiments/fraction_free_panel.py
"""
import numpy as np


def panel_standard(W, rhs, elig, q0):
    rows = W.shape[0]
    pe = W[:, q0:q0 + 4].copy()
    gc = np.zeros((rows, 4))
    pr = []
    for k in range(4):
        cand = np.where(elig, np.abs(pe[:, k]), -1.0)
        r = int(np.argmax(cand))
        pr.append(r)
        inv = 1.0 / pe[r, k]
        m = -(pe[:, k] * inv)
        m[r] = 0.0
        rp, gr = pe[r].copy(), gc[r].copy()
        for c in range(k + 1, 4):
            pe[:, c] = pe[:, c] + m * rp[c]
        for j in range(k):
            gc[:, j] = gc[:, j] + m * gr[j]
        gc[:, k] = m
        elig[r] = False
    return gc, pr


def panel_fraction_free(W, rhs, elig, q0):
    rows = W.shape[0]
    pe = W[:, q0:q0 + 4].copy()
    gcp = np.zeros((rows, 4))
    scale = np.ones(rows)      # s_i: row i currently holds s_i * (row at panel start) + sum_j gc'[i][j] P_j
    pr, piv = [], []
    S = 1.0                    # common scale of the rows that have not pivoted yet (warp-uniform on the GPU)
    for k in range(4):
        cand = np.where(elig, np.abs(pe[:, k]), -1.0)
        r = int(np.argmax(cand))
        pr.append(r)
        p = pe[r, k]
        piv.append(p)
        own = pe[:, k].copy()
        rp, gr = pe[r].copy(), gcp[r].copy()
        mask = np.ones(rows, bool)
        mask[r] = False
        for c in range(k + 1, 4):
            pe[mask, c] = p * pe[mask, c] - own[mask] * rp[c]
        for j in range(k):
            gcp[mask, j] = p * gcp[mask, j] - own[mask] * gr[j]
        gcp[mask, k] = -own[mask] * S      # the pivot row's own term: s_r = S because it was still eligible
        scale[mask] *= p
        S *= p
        elig[r] = False
    # back to true multipliers: one reciprocal per pivot, all off the serial chain
    f = np.full(rows, np.prod([1.0 / p for p in piv]))
    for q, r in enumerate(pr):
        f[r] *= piv[q]
    gc = gcp * f[:, None]
    assert np.allclose(scale * f, 1.0, rtol=1e-12)
    return gc, pr


def eliminate(W, rhs, panel):
    """Gauss-Jordan on the first 16 columns of W (32 x 48), panels of 4, rank-4 trailing updates (as the kernel does)."""
    W, rhs = W.copy(), rhs.copy()
    elig = np.ones(W.shape[0], bool)
    order = []
    for pn in range(4):
        q0 = 4 * pn
        gc, pr = panel(W, rhs, elig, q0)
        P, r0 = W[pr].copy(), rhs[pr].copy()
        W = W + gc @ P
        rhs = rhs + gc @ r0
        order += pr
    return W, rhs, order


def stacked_merge(rng, h, scale_rows=False):
    n = 16
    J = [rng.standard_normal((n, n)) for _ in range(2)]
    W = np.zeros((32, 48))
    W[:16, :16] = np.eye(n) - h * J[0]        # carried rows  [E | A | 0] = [R | L | 0]
    W[:16, 16:32] = -np.eye(n) - h * J[0]
    W[16:, :16] = -np.eye(n) - h * J[1]       # incoming rows [E | 0 | B] = [L | 0 | R]
    W[16:, 32:] = np.eye(n) - h * J[1]
    if scale_rows:
        W *= 10.0 ** rng.uniform(-6, 6, (32, 1))
    return W, rng.standard_normal(32)


def main():
    rng = np.random.default_rng(0)
    for scaled in (False, True):
        worst, differ, trials = 0.0, 0, 0
        for h in (1e-4, 1e-2, 0.3, 3.0):
            for _ in range(50):
                W, rhs = stacked_merge(rng, h, scaled)
                Ws, rs, os_ = eliminate(W, rhs, panel_standard)
                Wf, rf, of_ = eliminate(W, rhs, panel_fraction_free)
                trials += 1
                if os_ != of_:
                    differ += 1
                    continue
                live = np.abs(Ws[:, 16:]).max()
                worst = max(worst, np.abs(Ws[:, 16:] - Wf[:, 16:]).max() / live, np.abs(rs - rf).max() / max(1.0, np.abs(rs).max()))
        kind = "rows scaled by 10^U(-6,6)" if scaled else "ABD-like blocks -I - hJ | I - hJ, h = 1e-4 ... 3"
        print(f"{kind}: {trials} merges, pivot rows differ in {differ}, max relative difference of the surviving columns / rhs {worst:.2e}")


if __name__ == "__main__":
    main()
