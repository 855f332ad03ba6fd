"""ctypes binding of the CPU oracle (oracle/libmirk_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.  PARITY UNPINNED against a
live Julia run (see oracle/mirk_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmirk_oracle.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

RHS_FN = C.CFUNCTYPE(None, dp, dp, dp, C.c_double, C.c_void_p)
JAC_FN = C.CFUNCTYPE(None, dp, dp, dp, C.c_double, C.c_void_p)
TIMES_FN = C.CFUNCTYPE(C.c_int, dp, dp, C.c_double, C.c_double, C.c_void_p)
BC_FN = C.CFUNCTYPE(None, dp, dp, dp, C.c_void_p)
BCJ_FN = C.CFUNCTYPE(None, dp, dp, dp, C.c_void_p)

SUCCESS, FAILURE, MAXITERS, UNSTABLE, STALLED = 0, 1, 2, 3, 4


class Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("n_p", C.c_int), ("problem_type", C.c_int), ("n_bc", C.c_int),
                ("n_bca", C.c_int), ("f", RHS_FN), ("dfdu", JAC_FN), ("bc_times", TIMES_FN),
                ("bc", BC_FN), ("dbc", BCJ_FN), ("ctx", C.c_void_p), ("singular_term", dp),
                ("bc_uses_derivative", C.c_int)]


class Tableau(C.Structure):
    _fields_ = [("order", C.c_int), ("s", C.c_int), ("s_star", C.c_int),
                ("c", C.c_double * 5), ("v", C.c_double * 5), ("b", C.c_double * 5),
                ("x", (C.c_double * 5) * 5),
                ("c_star", C.c_double * 4), ("v_star", C.c_double * 4),
                ("x_star", (C.c_double * 9) * 4), ("tau_star", C.c_double)]


class Options(C.Structure):
    _fields_ = [("abstol", C.c_double), ("adaptive", C.c_int), ("defect_threshold", C.c_double),
                ("max_num_subintervals", C.c_int), ("maxiters", C.c_int),
                ("reinterp_inplace", C.c_int), ("max_outer", C.c_int), ("nlsolve", C.c_int),
                ("controller", C.c_int), ("ge_method", C.c_int), ("DE", C.c_double), ("GE", C.c_double)]


class Result(C.Structure):
    _fields_ = [("N", C.c_int), ("mesh", dp), ("y", dp), ("Kd", dp), ("Ki", dp),
                ("retcode", C.c_int), ("resid_norm", C.c_double), ("defect_norm", C.c_double),
                ("outer_iters", C.c_int), ("newton_iters", C.c_int), ("n_hist", C.c_int),
                ("hist_N", C.c_int * 64), ("hist_newton", C.c_int * 64),
                ("hist_defect", C.c_double * 64)]


def build(force: bool = False) -> str:
    """Compile the oracle with its committed Makefile (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("mirk_oracle.c", "mirk_problems.c", "mirk_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_tableau_get.argtypes = [C.c_int, C.POINTER(Tableau)]
        L.orc_set_interval_threads.argtypes = [C.c_int]
        L.orc_set_interval_threads.restype = None
        L.orc_interp_weights.argtypes = [C.c_int, C.c_double, dp, dp]
        L.orc_mesh_uniform.argtypes = [C.c_double, C.c_double, C.c_int, dp]
        L.orc_interval.argtypes = [dp, C.c_int, C.c_double]
        PP, TP = C.POINTER(Problem), C.POINTER(Tableau)
        L.orc_phi.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp]
        L.orc_interp_setup.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp]
        L.orc_eval_sol.argtypes = [PP, TP, C.c_int, dp, dp, dp, dp, C.c_double, C.c_int, C.c_int, dp]
        L.orc_loss.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp, dp]
        L.orc_jac_blocks.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp]
        L.orc_bc_jac.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp, ip, dp]
        L.orc_abd_solve.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, ip, dp, dp, dp, dp]
        L.orc_newton.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp, C.c_double, C.c_int, dp, ip]
        L.orc_nlsolve.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp, C.c_double, C.c_int, C.c_int, dp, ip]
        L.orc_defect.argtypes = [PP, TP, dp, C.c_int, dp, dp, dp, dp, dp]
        L.orc_defect.restype = C.c_double
        L.orc_mesh_select.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, C.c_int, ip, dp]
        L.orc_reinterp.argtypes = [PP, TP, C.c_int, dp, dp, dp, dp, C.c_int, dp, dp, C.c_int]
        L.orc_solve.argtypes = [PP, C.c_int, dp, C.c_int, dp, dp, C.POINTER(Options), C.POINTER(Result)]
        L.orc_result_free.argtypes = [C.POINTER(Result)]
        L.orc_default_options.argtypes = [C.POINTER(Options)]
        L.orc_builtin_problem.argtypes = [C.c_int, PP]
        L.orc_builtin_name.argtypes = [C.c_int]
        L.orc_builtin_name.restype = C.c_char_p
        L.orc_ensemble_solve.argtypes = [PP, C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, C.c_int,
                                         C.POINTER(Options), C.c_int, ip, ip, dp, ip]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


PROBLEM_IDS = {"pendulum": 0, "linear2": 1, "linear2_tp": 2, "swirling": 3, "lotka": 4,
               "torus": 5, "layer": 6, "chain8": 7, "chain16": 8, "bratu64": 9, "lane_emden": 10, "robin_sine": 11}


def builtin(name_or_id) -> Problem:
    pid = PROBLEM_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
    P = Problem()
    if lib().orc_builtin_problem(pid, C.byref(P)) != 0:
        raise KeyError(name_or_id)
    return P


def custom_problem(n, n_p, f, dfdu, bc_times, bc, dbc, problem_type=0, n_bc=None, n_bca=0):
    """Wrap python callables (numpy in / numpy out) as an oracle problem.  The returned object
    keeps the callbacks alive through `_keep`."""
    n_bc = n if n_bc is None else n_bc

    def _f(du, u, p, t, ctx):
        du_ = np.ctypeslib.as_array(du, (n,))
        du_[:] = f(np.ctypeslib.as_array(u, (n,)), np.ctypeslib.as_array(p, (max(n_p, 1),)), t)

    def _j(J, u, p, t, ctx):
        J_ = np.ctypeslib.as_array(J, (n, n))
        J_[:, :] = dfdu(np.ctypeslib.as_array(u, (n,)), np.ctypeslib.as_array(p, (max(n_p, 1),)), t)

    def _t(tm, p, t0, t1, ctx):
        ts = bc_times(np.ctypeslib.as_array(p, (max(n_p, 1),)), t0, t1)
        for k, v in enumerate(ts):
            tm[k] = v
        return len(ts)

    m_holder = {}

    def _m(p_):
        if "m" not in m_holder:
            m_holder["m"] = 2 if problem_type == 1 else len(bc_times(p_, 0.0, 1.0))
        return m_holder["m"]

    def _b(res, U, p, ctx):
        p_ = np.ctypeslib.as_array(p, (max(n_p, 1),))
        m = _m(p_)
        np.ctypeslib.as_array(res, (n_bc,))[:] = bc(np.ctypeslib.as_array(U, (m, n)), p_)

    def _bj(d, U, p, ctx):
        p_ = np.ctypeslib.as_array(p, (max(n_p, 1),))
        m = _m(p_)
        np.ctypeslib.as_array(d, (n_bc, m * n))[:, :] = dbc(np.ctypeslib.as_array(U, (m, n)), p_)

    cbs = (RHS_FN(_f), JAC_FN(_j), TIMES_FN(_t), BC_FN(_b), BCJ_FN(_bj))
    P = Problem(n, n_p, problem_type, n_bc, n_bca, *cbs, None)
    P._keep = cbs
    return P


MIRK6I = 7  # `order` code of the irrational 6th-order tableau (ORC_MIRK6I)


def set_interval_threads(nthreads: int) -> None:
    """Threads over mesh intervals in orc_phi / orc_jac_blocks (built-in problems only; bench.py's CPU baseline).  Results do not
    depend on it; default 1."""
    lib().orc_set_interval_threads(int(nthreads))


def tableau(order) -> Tableau:
    T = Tableau()
    if lib().orc_tableau_get(order, C.byref(T)) != 0:
        raise ValueError(f"unsupported order {order}")
    return T


def interp_weights(order, tau):
    ss = {2: 3, 3: 3, 4: 4, 5: 6, 6: 9, MIRK6I: 8}[order]
    w, wp = np.zeros(9), np.zeros(9)
    lib().orc_interp_weights(order, float(tau), _d(w), _d(wp))
    return w[:ss], wp[:ss]


def mesh_uniform(t0, t1, nint):
    m = np.zeros(nint + 1)
    lib().orc_mesh_uniform(float(t0), float(t1), int(nint), _d(m))
    return m


def interval(mesh, t):
    mesh = _arr(mesh)
    return lib().orc_interval(_d(mesh), len(mesh), float(t))


class Workspace:
    """Arrays of one evaluation on a fixed mesh (the oracle's MIRKCache)."""

    def __init__(self, P: Problem, order: int, p, mesh, y):
        self.P, self.T, self.order = P, tableau(order), order
        self.p = _arr(p) if len(np.atleast_1d(p)) else np.zeros(1)
        self.mesh = _arr(mesh)
        self.N = len(self.mesh)
        self.n = P.n
        self.y = _arr(y).reshape(self.N, self.n).copy()
        s, si = self.T.s, self.T.s_star - self.T.s
        self.Kd = np.zeros((self.N - 1, s, self.n))
        self.Ki = np.zeros((self.N - 1, si, self.n))

    def _a(self):
        return C.byref(self.P), C.byref(self.T), _d(self.p), self.N, _d(self.mesh), _d(self.y)

    def phi(self):
        out = np.zeros((self.N - 1, self.n))
        lib().orc_phi(*self._a(), _d(self.Kd), _d(out))
        return out

    def interp_setup(self):
        lib().orc_interp_setup(*self._a(), _d(self.Kd), _d(self.Ki))
        return self.Ki

    def loss(self):
        out = np.zeros(self.P.n_bc + (self.N - 1) * self.n)
        lib().orc_loss(*self._a(), _d(self.Kd), _d(self.Ki), _d(out))
        return out

    def jac_blocks(self):
        Lb = np.zeros((self.N - 1, self.n, self.n))
        Rb = np.zeros((self.N - 1, self.n, self.n))
        lib().orc_jac_blocks(*self._a(), _d(Lb), _d(Rb))
        return Lb, Rb

    def bc_jac(self):
        nodes = np.zeros(8, dtype=np.int32)
        B = np.zeros((8, self.P.n_bc, self.n))
        m = lib().orc_bc_jac(*self._a(), _d(self.Kd), _d(self.Ki), _i(nodes), _d(B))
        return nodes[:m].copy(), B[:m].copy()

    def dense_jacobian(self):
        """Global Jacobian in the reference's row order (for small-N checks)."""
        n, N, L = self.n, self.N, self.P.n_bc
        Lb, Rb = self.jac_blocks()
        nodes, B = self.bc_jac()
        J = np.zeros((L + (N - 1) * n, N * n))
        La = self.P.n_bca if self.P.problem_type == 1 else L
        off = La
        for i in range(N - 1):
            J[off + i * n: off + (i + 1) * n, i * n:(i + 1) * n] = Lb[i]
            J[off + i * n: off + (i + 1) * n, (i + 1) * n:(i + 2) * n] = Rb[i]
        for k, nd in enumerate(nodes):
            J[:La, nd * n:(nd + 1) * n] += B[k][:La]
            if La < L:
                J[La + (N - 1) * n:, nd * n:(nd + 1) * n] += B[k][La:]
        return J

    def dense_jacobian_reference_pattern(self):
        """The Jacobian the REFERENCE's default (sparse, coloured ForwardDiff) path assembles — quirk Q1 of SURVEY §8(c).

        The reference differentiates with a KNOWN sparsity pattern that is too narrow for n > 2
        (lib/BoundaryValueDiffEqMIRK/src/sparse_jacobians.jl:16-36: Standard problems band (1, 2n-1) over the collocation
        rows, two-point problems band (n+1, n+1) over all rows) and the greedy natural-order column colouring of a
        band of total width w = l + u + 1 is colour(c) = c mod w.  Coloured forward mode computes B = J_true * S (one
        column per colour) and decompression writes J[r, c] = B[r, colour(c)] for every (r, c) INSIDE the pattern: a true
        entry outside the band is not merely dropped, it is added to the in-band column of its colour in that row.
        For n <= 2 the pattern covers every true entry and this equals dense_jacobian().  (Restated from the published
        behaviour of SparseMatrixColorings / DifferentiationInterface, which are not under /root/reference: unpinned.)"""
        n, N, L = self.n, self.N, self.P.n_bc
        Jt = self.dense_jacobian()
        Jr = np.zeros_like(Jt)
        ncols = N * n
        if self.P.problem_type == 1:
            r0, r1, lo, up = 0, Jt.shape[0], n + 1, n + 1          # the whole matrix is one banded prototype
        else:
            r0, r1, lo, up = L, Jt.shape[0], 1, 2 * n - 1          # collocation rows only; the boundary rows are dense "fill"
            Jr[:L] = Jt[:L]
        w = lo + up + 1
        colour = np.arange(ncols) % w
        for R in range(r0, r1):
            r = R - r0
            B = np.bincount(colour, weights=Jt[R], minlength=w)    # compressed row: sum of the true entries per colour
            c0, c1 = max(0, r - lo), min(ncols - 1, r + up)
            cols = np.arange(c0, c1 + 1)
            Jr[R, cols] = B[colour[cols]]
        return Jr

    def newton_reference_pattern(self, abstol=1e-6, maxiters=1000, exact=False):
        """NewtonRaphson with dense solves of the reference-pattern Jacobian (exact=True: of the exact one, same code
        path) — for Newton-count comparisons at small N; O((nN)^3) per step."""
        it, ret = 0, MAXITERS
        nrm = float(np.max(np.abs(self.loss())))
        while it < maxiters:
            J = self.dense_jacobian() if exact else self.dense_jacobian_reference_pattern()
            try:
                d = np.linalg.solve(J, self.loss())
            except np.linalg.LinAlgError:
                ret = FAILURE
                break
            self.y -= d.reshape(self.y.shape)
            it += 1
            nrm = float(np.max(np.abs(self.loss())))
            if not np.isfinite(nrm):
                ret = UNSTABLE
                break
            if nrm <= abstol:
                ret = SUCCESS
                break
        return ret, it, nrm

    def eval_sol(self, t, deriv=0, bc_shortcut=False):
        out = np.zeros(self.n)
        lib().orc_eval_sol(C.byref(self.P), C.byref(self.T), self.N, _d(self.mesh), _d(self.y),
                           _d(self.Kd), _d(self.Ki), float(t), int(deriv), int(bc_shortcut), _d(out))
        return out

    def newton(self, abstol=1e-6, maxiters=1000):
        nrm, it = C.c_double(0), C.c_int(0)
        ret = lib().orc_newton(*self._a(), _d(self.Kd), _d(self.Ki), abstol, maxiters,
                               C.byref(nrm), C.byref(it))
        return ret, it.value, nrm.value

    def nlsolve(self, alg=0, abstol=1e-6, maxiters=1000):
        """alg 0: the reference's default polyalgorithm (NewtonRaphson -> + BackTracking -> TrustRegion), 1 / 2 / 3: one of them"""
        nrm, it = C.c_double(0), C.c_int(0)
        ret = lib().orc_nlsolve(*self._a(), _d(self.Kd), _d(self.Ki), abstol, maxiters, int(alg),
                                C.byref(nrm), C.byref(it))
        return ret, it.value, nrm.value

    def defect(self):
        err = np.zeros((self.N - 1, self.n))
        d = lib().orc_defect(*self._a(), _d(self.Kd), _d(self.Ki), _d(err))
        return d, err


def abd_solve(Lb, Rb, nodes, B, rhs_bc, rhs_phi):
    Lb, Rb, B = _arr(Lb), _arr(Rb), _arr(B)
    Nm1, n, _ = Lb.shape
    nodes = np.ascontiguousarray(nodes, dtype=np.int32)
    rhs_bc, rhs_phi = _arr(rhs_bc), _arr(rhs_phi)
    delta = np.zeros((Nm1 + 1, n))
    st = lib().orc_abd_solve(n, Nm1 + 1, len(rhs_bc), _d(Lb), _d(Rb), len(nodes), _i(nodes), _d(B),
                             _d(rhs_bc), _d(rhs_phi), _d(delta))
    return st, delta


def mesh_select(order, mesh, errors, abstol=1e-6, max_num_subintervals=3000):
    mesh, errors = _arr(mesh), _arr(errors)
    N, n = len(mesh), errors.shape[1]
    out = np.zeros(4 * (N - 1) + 1)
    Nn = C.c_int(0)
    info = lib().orc_mesh_select(order, n, N, _d(mesh), _d(errors), abstol, max_num_subintervals,
                                 C.byref(Nn), _d(out))
    return info, out[:Nn.value].copy()


def reinterp(ws: Workspace, mesh_new, inplace_quirk=True):
    mesh_new = _arr(mesh_new)
    y_new = np.zeros((len(mesh_new), ws.n))
    lib().orc_reinterp(C.byref(ws.P), C.byref(ws.T), ws.N, _d(ws.mesh), _d(ws.y), _d(ws.Kd), _d(ws.Ki),
                       len(mesh_new), _d(mesh_new), _d(y_new), int(inplace_quirk))
    return y_new


def default_options(**kw) -> Options:
    o = Options()
    lib().orc_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Solution:
    def __init__(self, P, order, R: Result):
        n = P.n
        T = tableau(order)
        self.P, self.order, self.T, self.n = P, order, T, n
        self.N = R.N
        self.t = np.ctypeslib.as_array(R.mesh, (R.N,)).copy()
        self.u = np.ctypeslib.as_array(R.y, (R.N, n)).copy()
        self.Kd = np.ctypeslib.as_array(R.Kd, (R.N - 1, T.s, n)).copy()
        self.Ki = np.ctypeslib.as_array(R.Ki, (R.N - 1, T.s_star - T.s, n)).copy()
        self.retcode, self.resid_norm, self.defect_norm = R.retcode, R.resid_norm, R.defect_norm
        self.outer_iters, self.newton_iters = R.outer_iters, R.newton_iters
        self.hist_N = list(R.hist_N[:R.n_hist])
        self.hist_newton = list(R.hist_newton[:R.n_hist])
        self.hist_defect = list(R.hist_defect[:R.n_hist])

    def __call__(self, t, deriv=0):
        out = np.zeros(self.n)
        lib().orc_eval_sol(C.byref(self.P), C.byref(self.T), self.N, _d(self.t), _d(self.u),
                           _d(self.Kd), _d(self.Ki), float(t), int(deriv), 0, _d(out))
        return out


def solve(P: Problem, order, p, mesh, y, **opts) -> Solution:
    mesh = _arr(mesh)
    N = len(mesh)
    y = _arr(y)
    if y.ndim == 1 and y.size == P.n:
        y = np.tile(y, (N, 1))
    y = np.ascontiguousarray(y.reshape(N, P.n))
    p = _arr(p) if len(np.atleast_1d(p)) else np.zeros(1)
    o = default_options(**opts)
    R = Result()
    st = lib().orc_solve(C.byref(P), order, _d(p), N, _d(mesh), _d(y), C.byref(o), C.byref(R))
    if st != 0:
        raise RuntimeError("orc_solve failed")
    sol = Solution(P, order, R)
    lib().orc_result_free(C.byref(R))
    return sol


def solve_dt(P, order, p, u0, tspan, dt, **opts) -> Solution:
    """solve(prob, alg; dt): Nig = cld(t1-t0, dt) (CORE/utils.jl:362), constant guess u0."""
    if dt <= 0:
        raise ValueError("dt must be positive")
    t0, t1 = tspan
    nint = int(np.ceil((t1 - t0) / dt))
    return solve(P, order, p, mesh_uniform(t0, t1, nint), np.asarray(u0, dtype=float), **opts)


def ensemble_solve(P, order, params, u0, tspan, nint, nthreads=1, **opts):
    params = _arr(params).reshape(-1, P.n_p)
    ntraj = params.shape[0]
    u0 = _arr(u0)
    ret = np.zeros(ntraj, dtype=np.int32)
    Nf = np.zeros(ntraj, dtype=np.int32)
    its = np.zeros(ntraj, dtype=np.int32)
    y0 = np.zeros((ntraj, P.n))
    o = default_options(**opts)
    lib().orc_ensemble_solve(C.byref(P), order, ntraj, _d(params), _d(u0), float(tspan[0]),
                             float(tspan[1]), int(nint), C.byref(o), int(nthreads), _i(ret), _i(Nf),
                             _d(y0), _i(its))
    return ret, Nf, y0, its
