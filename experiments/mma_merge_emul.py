"""SIMT-style emulation (numpy, 32 explicit lanes) of the DMMA-fragment merge in abd_mma.cuh, n = 16.

Every per-lane register is an array with a leading lane axis, shared memory is a plain array, shuffles
are gathers — so each line maps one to one to the CUDA code and the layout logic (who holds what, who
reads what) is checked on the CPU against a straightforward sequential Gauss-Jordan elimination.

    python experiments/mma_merge_emul.py
"""
import numpy as np

n = 16
LANES = np.arange(32)
G_ = LANES >> 2   # groupID
T_ = LANES & 3    # threadID_in_group


def mma_m8n8k4(c, a, b):
    """c: (32, 2) fragment, a: (32,), b: (32,) -> c + A @ B in fragment layout."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    for l in range(32):
        A[G_[l], T_[l]] = a[l]
        B[T_[l], G_[l]] = b[l]
        C[G_[l], 2 * T_[l]] = c[l, 0]; C[G_[l], 2 * T_[l] + 1] = c[l, 1]
    D = C + A @ B
    out = np.zeros_like(c)
    for l in range(32):
        out[l, 0] = D[G_[l], 2 * T_[l]]; out[l, 1] = D[G_[l], 2 * T_[l] + 1]
    return out


def to_frag(W):
    """W (32, 48) -> w[lane][tr][j][e]"""
    w = np.zeros((32, 4, 6, 2))
    for l in range(32):
        for tr in range(4):
            for j in range(6):
                for e in range(2):
                    w[l, tr, j, e] = W[8 * tr + G_[l], 8 * j + 2 * T_[l] + e]
    return w


def from_frag(w):
    W = np.zeros((32, 48))
    for l in range(32):
        for tr in range(4):
            for j in range(6):
                for e in range(2):
                    W[8 * tr + G_[l], 8 * j + 2 * T_[l] + e] = w[l, tr, j, e]
    return W


def merge_mma(W, rhs):
    """Gauss-Jordan on the 16 E columns of W (32 x 48) / rhs (32,), panel by panel, in fragment layout."""
    w = to_frag(W)
    rhs = rhs.copy()                      # lane i holds rhs of row i
    elig = np.ones(32, bool)
    myq = -np.ones(32, int); myinv = np.zeros(32)
    PS = 52
    for pn in range(4):
        q0 = 4 * pn; jp = q0 // 8; cq = q0 % 8; t0 = cq // 2
        # A: gather the panel into lane-per-row form through shared memory
        CS = 36
        Wp = np.zeros(4 * CS)             # column-major, column stride CS
        for l in range(32):
            if T_[l] in (t0, t0 + 1):
                for tr in range(4):
                    Wp[(2 * (T_[l] - t0)) * CS + 8 * tr + G_[l]] = w[l, tr, jp, 0]
                    Wp[(2 * (T_[l] - t0) + 1) * CS + 8 * tr + G_[l]] = w[l, tr, jp, 1]
        pe = np.stack([Wp[c * CS + LANES] for c in range(4)], axis=1)   # lane i reads row i
        # B: panel steps, lane per row
        g = np.zeros((32, 4)); prs = []
        for k in range(4):
            own = pe[:, k]
            key = np.where(elig, np.abs(own), -1.0)
            pr = int(np.argmax(key)); prs.append(pr)
            inv = (1.0 / own)[pr]
            m = -(own * inv); m[pr] = 0.0
            for c in range(k + 1, 4):
                pe[:, c] = pe[:, c] + m * pe[pr, c]
            for j in range(k):
                g[:, j] = g[:, j] + m * g[pr, j]
            g[:, k] = m
            elig[pr] = False; myq[pr] = q0 + k; myinv[pr] = 1.0 / own[pr]
        rhs0 = rhs.copy()
        for j in range(4):
            rhs = rhs + g[:, j] * rhs0[prs[j]]
        # C: coefficients into A-fragment layout through shared memory
        Gs = np.zeros(4 * CS)             # Gs[j][row], over the gathered panel
        for j in range(4):
            Gs[j * CS + LANES] = g[:, j]
        a = np.zeros((32, 4))
        for l in range(32):
            for tr in range(4):
                a[l, tr] = Gs[T_[l] * CS + 8 * tr + G_[l]]
        # D: the 4 pivot rows at panel start into shared lines P[k][col]
        P = np.zeros(4 * PS)
        for k in range(4):
            trk, gk = prs[k] >> 3, prs[k] & 7
            for l in range(32):
                if G_[l] == gk:
                    for j in range(6):
                        P[k * PS + 8 * j + 2 * T_[l]: k * PS + 8 * j + 2 * T_[l] + 2] = w[l, trk, j, :]
        # E: B fragments and the rank-4 update of the live tiles
        jlo = jp if cq == 0 else jp + 1
        for j in range(jlo, 6):
            b = np.array([P[T_[l] * PS + 8 * j + G_[l]] for l in range(32)])
            for tr in range(4):
                w[:, tr, j, :] = mma_m8n8k4(w[:, tr, j, :], a[:, tr], b)
    return from_frag(w), rhs, myq, myinv


def merge_seq(W, rhs):
    M = np.concatenate([W, rhs[:, None]], axis=1).copy()
    elig = np.ones(32, bool); myq = -np.ones(32, int); myinv = np.zeros(32)
    for q in range(n):
        pr = int(np.argmax(np.where(elig, np.abs(M[:, q]), -1.0)))
        inv = 1.0 / M[pr, q]
        m = -(M[:, q] * inv); m[pr] = 0.0
        M[:, q + 1:] += np.outer(m, M[pr, q + 1:])
        elig[pr] = False; myq[pr] = q; myinv[pr] = inv
    return M[:, :48], M[:, 48], myq, myinv


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    for trial in range(5):
        W = rng.standard_normal((32, 48)); rhs = rng.standard_normal(32)
        W[:16, 32:] = 0.0; W[16:, 16:32] = 0.0   # carried rows have no B part, incoming rows no A part
        Wa, ra, qa, ia = merge_seq(W, rhs)
        Wb, rb, qb, ib = merge_mma(W, rhs)
        live = slice(16, 48)
        print(trial, (qa == qb).all(), np.abs(Wa[:, live] - Wb[:, live]).max(), np.abs(ra - rb).max(), np.abs(ia - ib).max())
