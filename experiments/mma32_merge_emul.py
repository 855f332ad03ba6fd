"""SIMT-style emulation (numpy, 4 warps x 32 explicit lanes) of the n = 32 DMMA-fragment merge in
boundaryvaluediffeq.jl_b200/csrc/abd_mma32.cuh: the 64 x 96 matrix [E | A | B] of a merge over FOUR warps, warp w owning tile rows
2w, 2w+1 (rows 16w .. 16w+15) in the C-fragment layout of mma.m8n8k4, lanes 0..15 of a warp "holding" the
panel entries / rhs / pivot bookkeeping of the warp's 16 rows.  Checked against a sequential Gauss-Jordan.

    python experiments/mma32_merge_emul.py
"""
import numpy as np

n = 32
NT = 128
TID = np.arange(NT)
WARP = TID >> 5
LANE = TID & 31
G_ = LANE >> 2
T_ = LANE & 3


def mma_m8n8k4(c, a, b):
    """one warp: c (32, 2), a (32,), b (32,) -> c + A @ B in fragment layout"""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    for l in range(32):
        g, t = l >> 2, l & 3
        A[g, t] = a[l]; B[t, g] = b[l]
        C[g, 2 * t] = c[l, 0]; C[g, 2 * t + 1] = c[l, 1]
    D = C + A @ B
    out = np.zeros_like(c)
    for l in range(32):
        g, t = l >> 2, l & 3
        out[l, 0] = D[g, 2 * t]; out[l, 1] = D[g, 2 * t + 1]
    return out


def to_frag(W):
    w = np.zeros((NT, 2, 12, 2))
    for tid in range(NT):
        for trl in range(2):
            for j in range(12):
                for e in range(2):
                    w[tid, trl, j, e] = W[16 * WARP[tid] + 8 * trl + G_[tid], 8 * j + 2 * T_[tid] + e]
    return w


def from_frag(w):
    W = np.zeros((64, 96))
    for tid in range(NT):
        for trl in range(2):
            for j in range(12):
                for e in range(2):
                    W[16 * WARP[tid] + 8 * trl + G_[tid], 8 * j + 2 * T_[tid] + e] = w[tid, trl, j, e]
    return W


def merge_mma32(W, rhs_rows):
    w = to_frag(W)
    holder = LANE < 16
    row = 16 * WARP + LANE                    # row held by a holder thread
    rhs = np.where(holder, rhs_rows[np.minimum(row, 63)], 0.0)
    elig = holder.copy()
    myq = -np.ones(NT, int); myinv = np.zeros(NT)
    CSW, PS = 20, 100
    for pn in range(8):
        q0 = 4 * pn; jp = q0 >> 3; cq = q0 & 7; t0 = cq >> 1
        # (A) warp-local gather of the panel
        Wp = np.zeros((4, 4 * CSW))
        for tid in range(NT):
            if T_[tid] in (t0, t0 + 1):
                for trl in range(2):
                    for e in range(2):
                        Wp[WARP[tid], (2 * (T_[tid] - t0) + e) * CSW + 8 * trl + G_[tid]] = w[tid, trl, jp, e]
        pe = np.zeros((NT, 4))
        for tid in range(NT):
            if holder[tid]:
                for c in range(4):
                    pe[tid, c] = Wp[WARP[tid], c * CSW + LANE[tid]]
        # (B) pivot steps: every warp publishes its best candidate's record, one block barrier per step
        gc = np.zeros((NT, 4)); prs = []; rhs0_pr = []
        rhs0 = rhs.copy()
        for k in range(4):
            own = pe[:, k]
            key = np.where(elig, np.abs(own), -1.0)
            rec = []
            for wv in range(4):
                sel = np.where(WARP == wv)[0]
                best = sel[np.argmax(key[sel])]
                rec.append((key[best], best))
            kb, best = max(rec, key=lambda r: (r[0], -r[1]))
            pr = int(row[best]); prs.append(pr); rhs0_pr.append(rhs0[best])
            inv = 1.0 / own[best]
            P_pe = pe[best].copy(); P_gc = gc[best].copy()
            m = np.where(holder, -(own * inv), 0.0); m[best] = 0.0
            for c in range(k + 1, 4):
                pe[:, c] = pe[:, c] + m * P_pe[c]
            for j in range(k):
                gc[:, j] = gc[:, j] + m * P_gc[j]
            gc[:, k] = m
            elig[best] = False; myq[best] = q0 + k; myinv[best] = inv
        for j in range(4):
            rhs = rhs + gc[:, j] * rhs0_pr[j]
        # (C) coefficients, warp-local, over the gathered panel
        Gs = np.zeros((4, 4 * CSW))
        for tid in range(NT):
            if holder[tid]:
                for j in range(4):
                    Gs[WARP[tid], j * CSW + LANE[tid]] = gc[tid, j]
        a = np.zeros((NT, 2))
        for tid in range(NT):
            for trl in range(2):
                a[tid, trl] = Gs[WARP[tid], T_[tid] * CSW + 8 * trl + G_[tid]]
        # (D) the 4 pivot rows (any warp) into the block-wide lines
        jlo = jp if cq == 0 else jp + 1
        P = np.zeros(4 * PS)
        for tid in range(NT):
            for trl in range(2):
                r = 16 * WARP[tid] + 8 * trl + G_[tid]
                if r in prs:
                    kk = prs.index(r)
                    for j in range(jlo, 12):
                        P[kk * PS + 8 * j + 2 * T_[tid]: kk * PS + 8 * j + 2 * T_[tid] + 2] = w[tid, trl, j, :]
        # (E) rank-4 update of the live tiles, every warp on its own two tile rows
        for wv in range(4):
            sel = np.where(WARP == wv)[0]
            for j in range(jlo, 12):
                b = np.array([P[T_[tid] * PS + 8 * j + G_[tid]] for tid in sel])
                for trl in range(2):
                    w[sel, trl, j, :] = mma_m8n8k4(w[sel, trl, j, :], a[sel, trl], b)
    rhs_rows_out = np.zeros(64); q_rows = -np.ones(64, int); inv_rows = np.zeros(64)
    for tid in range(NT):
        if holder[tid]:
            rhs_rows_out[row[tid]] = rhs[tid]; q_rows[row[tid]] = myq[tid]; inv_rows[row[tid]] = myinv[tid]
    return from_frag(w), rhs_rows_out, q_rows, inv_rows


def merge_seq(W, rhs):
    M = np.concatenate([W, rhs[:, None]], axis=1).copy()
    rows = M.shape[0]
    elig = np.ones(rows, bool); myq = -np.ones(rows, int); myinv = np.zeros(rows)
    for q in range(n):
        pr = int(np.argmax(np.where(elig, np.abs(M[:, q]), -1.0)))
        inv = 1.0 / M[pr, q]
        m = -(M[:, q] * inv); m[pr] = 0.0
        M[:, q + 1:] += np.outer(m, M[pr, q + 1:])
        elig[pr] = False; myq[pr] = q; myinv[pr] = inv
    return M[:, :-1], M[:, -1], myq, myinv


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    for trial in range(3):
        W = rng.standard_normal((64, 96)); rhs = rng.standard_normal(64)
        W[:32, 64:] = 0.0; W[32:, 32:64] = 0.0
        Wa, ra, qa, ia = merge_seq(W, rhs)
        Wb, rb, qb, ib = merge_mma32(W, rhs)
        print(trial, (qa == qb).all(), np.abs(Wa[:, 32:] - Wb[:, 32:]).max(), np.abs(ra - rb).max(), np.abs(ia - ib).max())
