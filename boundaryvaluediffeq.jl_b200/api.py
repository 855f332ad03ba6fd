"""Host-side mirror of the SciML interface for the MIRK path, over the C ABI.

The reference's toolchain (Julia) is absent from this image, so the host layer a Julia
`BoundaryValueDiffEqMIRK` backend would provide is written here in Python with the same names,
argument meaning and error behaviour, so tests read like the reference's own
(lib/BoundaryValueDiffEqMIRK/test/Core/mirk_basic_tests.jl).  The Julia glue that binds the same
C ABI with `ccall` is under julia/ and documented in INTEGRATION.md.

    prob = BVProblem(BVPDeviceFunction("pendulum"), [pi/2, pi/2], (0, pi/2), p=[9.81])
    sol  = solve(prob, MIRK4(), dt=0.05)         # MIRK/src/mirk.jl:49-53, :286-332
    sol.u, sol.t, sol(0.3), sol.retcode, sol.resid

The RHS / boundary conditions are *device functors* (csrc/problems.cuh) named in a registry —
arbitrary host closures cannot run inside a CUDA kernel — either built in or compiled from a
user-supplied CUDA functor with `compile_device_function`.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import tempfile
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib as B

__all__ = ["GlobalErrorControl", "SequentialErrorControl", "HybridErrorControl", "HOErrorControl", "REErrorControl", "NewtonRaphson", "BackTracking", "TrustRegion", "BVPDeviceFunction", "BVProblem", "TwoPointBVProblem", "MIRK2", "MIRK3", "MIRK4", "MIRK5", "MIRK6", "MIRK6I", "DefectControl",
           "BVPJacobianAlgorithm", "ReturnCode", "BVSolution", "MIRKCache", "init", "solve", "solve_b", "abd_solve",
           "EnsembleProblem", "EnsembleSolution", "EnsembleB200", "compile_device_function",
           "successful_retcode"]


def _d(a):
    return a.ctypes.data_as(B.dp)


def _i(a):
    return a.ctypes.data_as(B.ip)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class ReturnCode:
    """SciMLBase.ReturnCode values this path produces."""
    Success, Failure, MaxIters, Unstable, Stalled = 0, 1, 2, 3, 4
    names = B.RETCODE_NAMES


def successful_retcode(sol) -> bool:
    return sol.retcode == ReturnCode.Success


# ---- device functions ---------------------------------------------------------------------------
@dataclass(frozen=True)
class BVPDeviceFunction:
    """Stands where `BVPFunction(f!, bc!)` stands in the reference (MIRK/src/mirk.jl:71-116): the
    name of a device functor providing f, bc and the BC evaluation times."""
    name: str

    @property
    def problem_id(self) -> int:
        pid = C.c_int32(-1)
        B.check(B.lib().mirk_problem_lookup(self.name.encode(), C.byref(pid)))
        return pid.value

    @property
    def info(self) -> B.ProblemInfo:
        info = B.ProblemInfo()
        B.check(B.lib().mirk_problem_info_get(self.problem_id, C.byref(info)))
        return info


_PLUGIN_TEMPLATE = r"""
#include "ops.cuh"
namespace mirk { namespace problems {
%(source)s
} }
extern "C" const mirk::ProblemOps* mirk_plugin_ops(int order) {
    static const mirk::ProblemOps o4 = mirk::OpsImpl<mirk::problems::%(struct)s, 4>::make("%(name)s");
    static const mirk::ProblemOps o6 = mirk::OpsImpl<mirk::problems::%(struct)s, 6>::make("%(name)s");
    return order == 4 ? &o4 : order == 6 ? &o6 : nullptr;
}
"""


def compile_device_function(name: str, struct_name: str, source: str, workdir: Optional[str] = None) -> BVPDeviceFunction:
    """Compile a user functor (the contract at the top of csrc/problems.cuh) for sm_100a with nvcc
    and register it; one templated source serves residual (double) and Jacobian (Dual)."""
    workdir = workdir or tempfile.mkdtemp(prefix="mirk_plugin_")
    cu = os.path.join(workdir, f"{name}.cu")
    so = os.path.join(workdir, f"libmirk_plugin_{name}.so")
    with open(cu, "w") as fh:
        fh.write(_PLUGIN_TEMPLATE % {"source": source, "struct": struct_name, "name": name})
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-I", B.CSRC, "-I", B.INCLUDE, cu, "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for device function %s:\n%s" % (name, r.stderr[-3000:]))
    pid = C.c_int32(-1)
    B.check(B.lib().mirk_problem_register_plugin(name.encode(), so.encode(), C.byref(pid)))
    return BVPDeviceFunction(name)


# ---- problem / algorithm types ---------------------------------------------------------------------
@dataclass
class BVProblem:
    """BVProblem(f, u0, tspan, p).  `u0` is a state vector (constant guess, CORE/utils.jl:766-769),
    an (N, n) array / list of state vectors (guess on a uniform mesh of N nodes, CORE/utils.jl:339-348,
    694), a callable `u0(p, t)`, or an object carrying its own mesh — a previous solution (`BVSolution`)
    or anything with `.t` and `.u` like the reference's DiffEqArray / ODESolution guesses, whose nodes
    become the initial mesh (CORE/utils.jl:701-704).  `mesh=` overrides the node positions of an (N, n)
    guess.  The guess object is copied, never mutated (mirk_basic_tests.jl:717-719)."""
    f: BVPDeviceFunction
    u0: object
    tspan: Sequence[float]
    p: Sequence[float] = ()
    mesh: Optional[Sequence[float]] = None

    def __post_init__(self):
        if isinstance(self.f, str):
            self.f = BVPDeviceFunction(self.f)
        self.tspan = (float(self.tspan[0]), float(self.tspan[1]))
        self.p = _arr(self.p).ravel()
        if hasattr(self.u0, "t") and hasattr(self.u0, "u"):   # (a BVSolution is callable too: check this first)
            # guess AND mesh from a previous solution / DiffEqArray-like object
            if self.mesh is None:
                self.mesh = np.array(self.u0.t, dtype=np.float64)
            self.u0 = np.array([np.asarray(v, dtype=np.float64).ravel() for v in self.u0.u])
        if not callable(self.u0):
            self.u0 = _arr(self.u0)

    @property
    def problem_type(self) -> str:
        return "TwoPointBVProblem" if self.f.info.problem_type == 1 else "StandardBVProblem"

    def remake(self, **kw) -> "BVProblem":
        d = dict(f=self.f, u0=self.u0, tspan=self.tspan, p=self.p, mesh=self.mesh)
        d.update(kw)
        return BVProblem(**d)


def TwoPointBVProblem(f, u0, tspan, p=(), **kw) -> BVProblem:
    prob = BVProblem(f, u0, tspan, p, **kw)
    if prob.f.info.problem_type != 1:
        raise ValueError(f"device function {prob.f.name!r} is not a two-point problem")
    return prob


@dataclass(frozen=True)
class DefectControl:
    """CORE/src/calc_errors.jl:54-60"""
    defect_threshold: float = 0.1


@dataclass(frozen=True)
class HOErrorControl:
    """CORE/src/calc_errors.jl:126 — global error from the method of order + 2 on the same mesh"""


@dataclass(frozen=True)
class REErrorControl:
    """CORE/src/calc_errors.jl:139 — Richardson extrapolation: the same method on the halved mesh"""


@dataclass(frozen=True)
class GlobalErrorControl:
    """CORE/src/calc_errors.jl:67-73"""
    method: object = field(default_factory=HOErrorControl)


@dataclass(frozen=True)
class SequentialErrorControl:
    """CORE/src/calc_errors.jl:80-87: defect control first, global-error control once the defect passes"""
    defect: DefectControl = field(default_factory=DefectControl)
    global_error: GlobalErrorControl = field(default_factory=GlobalErrorControl)


@dataclass(frozen=True)
class HybridErrorControl:
    """CORE/src/calc_errors.jl:94-106: error = DE * defect + GE * global error"""
    DE: float = 1.0
    GE: float = 1.0
    defect: DefectControl = field(default_factory=DefectControl)
    global_error: GlobalErrorControl = field(default_factory=GlobalErrorControl)


def _controller_fields(c):
    """controller -> (code, ge_method, DE, GE, defect_threshold) of the C ABI"""
    gm = lambda g: 1 if isinstance(g.method, REErrorControl) else 0  # noqa: E731
    if isinstance(c, DefectControl):
        return 0, 0, 1.0, 1.0, float(c.defect_threshold)
    if isinstance(c, GlobalErrorControl):
        return 1, gm(c), 1.0, 1.0, 0.1
    if isinstance(c, SequentialErrorControl):
        return 2, gm(c.global_error), 1.0, 1.0, float(c.defect.defect_threshold)
    if isinstance(c, HybridErrorControl):
        return 3, gm(c.global_error), float(c.DE), float(c.GE), float(c.defect.defect_threshold)
    raise NotImplementedError(f"unknown error controller {type(c).__name__}")


@dataclass(frozen=True)
class BVPJacobianAlgorithm:
    """CORE/src/types.jl:13-128.  This backend always builds the exact block Jacobian by per-interval
    dual numbers; the fields are kept so reference call sites construct it unchanged."""
    bc_diffmode: object = None
    nonbc_diffmode: object = None
    diffmode: object = None


@dataclass(frozen=True)
class BackTracking:
    """LineSearch.BackTracking() — the line search of the polyalgorithm's second solver"""


@dataclass(frozen=True)
class NewtonRaphson:
    """NonlinearSolve.NewtonRaphson(; linesearch): `nlsolve = NewtonRaphson()` / `NewtonRaphson(linesearch = BackTracking())`"""
    linesearch: object = None


@dataclass(frozen=True)
class TrustRegion:
    """NonlinearSolve.TrustRegion() (Simple radius update, dogleg step)"""


def _nlsolve_code(nlsolve) -> int:
    """alg.nlsolve -> the C ABI's solver code.  None is the reference default: the polyalgorithm NewtonRaphson ->
    NewtonRaphson + BackTracking -> TrustRegion (CORE/src/default_internal_solve.jl:31-45)."""
    if nlsolve is None:
        return 0
    if isinstance(nlsolve, TrustRegion):
        return 3
    if isinstance(nlsolve, NewtonRaphson):
        if nlsolve.linesearch is None:
            return 1
        if isinstance(nlsolve.linesearch, BackTracking):
            return 2
    raise NotImplementedError("nlsolve must be None (the default polyalgorithm), NewtonRaphson(), "
                              "NewtonRaphson(linesearch=BackTracking()) or TrustRegion(): other solvers cannot run on the device")


@dataclass(frozen=True)
class _AbstractMIRK:
    """MIRK/src/algorithms.jl:55-61"""
    nlsolve: object = None
    optimize: object = None
    jac_alg: BVPJacobianAlgorithm = field(default_factory=BVPJacobianAlgorithm)
    defect_threshold: float = 0.1
    max_num_subintervals: int = 3000
    order = 0

    def __post_init__(self):
        if self.optimize is not None:
            raise NotImplementedError("`optimize` solvers are not supported by the B200 backend")
        _nlsolve_code(self.nlsolve)


@dataclass(frozen=True)
class MIRK2(_AbstractMIRK):
    order = 2


@dataclass(frozen=True)
class MIRK3(_AbstractMIRK):
    order = 3


@dataclass(frozen=True)
class MIRK4(_AbstractMIRK):
    order = 4


@dataclass(frozen=True)
class MIRK5(_AbstractMIRK):
    order = 5


@dataclass(frozen=True)
class MIRK6(_AbstractMIRK):
    order = 6


@dataclass(frozen=True)
class MIRK6I(_AbstractMIRK):
    """The 6th-order tableau with irrational abscissae (MIRK/src/mirk_tableaus.jl:154-194); `order` is the C ABI's
    tableau code 7, the convergence order is 6."""
    order = 7


# ---- cache / solution ----------------------------------------------------------------------------
class BVSolution:
    """What `solve` returns (MIRK/src/mirk.jl:324-331): `.u`, `.t`, `sol(t)`, `.retcode`, `.resid`,
    `.prob`, `.alg`, `.original` (iteration statistics)."""

    def __init__(self, cache: "MIRKCache", res: B.Result):
        n = cache.n
        N = res.n_mesh
        self.prob, self.alg = cache.prob, cache.alg
        self.t = np.zeros(N)
        self.u = np.zeros((N, n))
        B.check(B.lib().mirk_get_solution(cache._h, _d(self.t), _d(self.u)))
        L = cache.info.n_bc
        self.resid = np.zeros(L + (N - 1) * n)
        B.check(B.lib().mirk_get_residual(cache._h, _d(self.resid)))
        self.retcode = int(res.retcode)
        self.original = {
            "resid_norm": res.resid_norm, "defect_norm": res.defect_norm, "outer_iters": res.outer_iters,
            "newton_iters": res.newton_iters, "hist_n_mesh": list(res.hist_n_mesh[:res.n_hist]),
            "hist_newton": list(res.hist_newton[:res.n_hist]), "hist_defect": list(res.hist_defect[:res.n_hist])}
        self._cache = cache  # keeps the device stages alive for sol(t)

    @property
    def converged(self) -> bool:
        return self.retcode == ReturnCode.Success

    def __call__(self, t, deriv: int = 0, idxs=None):
        """sol(t) / sol(t, Val{1}) (MIRK/src/interpolation.jl:17-96)"""
        ts = _arr(np.atleast_1d(t))
        out = np.zeros((len(ts), self._cache.n))
        B.check(B.lib().mirk_interp(self._cache._h, _d(ts), len(ts), int(deriv), _d(out)))
        if idxs is not None:
            out = out[:, idxs]
        return out[0] if np.ndim(t) == 0 else out

    def stages(self):
        N, n, c = len(self.t), self._cache.n, self._cache
        Kd = np.zeros((N - 1, c.s, n))
        Ki = np.zeros((N - 1, c.s_star - c.s, n))
        B.check(B.lib().mirk_get_stages(c._h, _d(Kd), _d(Ki)))
        return Kd, Ki


class MIRKCache:
    """`init(prob, alg; dt, ...)` (MIRK/src/mirk.jl:49-265).  Owns one C-ABI handle."""

    def __init__(self, prob: BVProblem, alg: _AbstractMIRK, dt: float = 0.0, abstol: float = 1e-6,
                 adaptive: bool = True, controller: DefectControl = DefectControl(), nlsolve_kwargs=None,
                 optimize_kwargs=None, verbose=None, device: int = 0, chunk: int = 0,
                 reinterp_inplace: bool = False):
        ctrl, gem, DE, GE, thr = _controller_fields(controller)
        self.prob, self.alg, self.verbose = prob, alg, verbose
        self.nlsolve_kwargs = dict(nlsolve_kwargs or {})
        self.nlsolve_kwargs.setdefault("abstol", abstol)
        self.optimize_kwargs = dict(optimize_kwargs or {})
        self.info = prob.f.info
        self.n = self.info.n
        self.order = alg.order
        self.s, self.s_star = {2: (1, 3), 3: (2, 3), 4: (3, 4), 5: (4, 6), 6: (5, 9), 7: (5, 8)}[alg.order]
        p = prob.p
        self._p = p
        desc = B.Desc(prob.f.problem_id, alg.order, float(self.nlsolve_kwargs["abstol"]), int(bool(adaptive)),
                      thr, int(alg.max_num_subintervals),
                      int(self.nlsolve_kwargs.get("maxiters", 1000)), int(reinterp_inplace), int(chunk), int(device),
                      len(p), _d(p) if len(p) else None, _nlsolve_code(alg.nlsolve), ctrl, gem, DE, GE)
        self._h = B.Handle()
        B.check(B.lib().mirk_create(C.byref(desc), C.byref(self._h)))
        t0, t1 = prob.tspan
        u0 = prob.u0
        if callable(u0) or (isinstance(u0, np.ndarray) and u0.ndim == 1):
            if not (dt > 0):
                self.close()
                raise ValueError("dt must be positive")
        if callable(u0):
            nint = int(math.ceil((t1 - t0) / dt))
            mesh = mesh_uniform(t0, t1, nint)
            y = _arr(np.stack([_arr(u0(p, t)).ravel() for t in mesh]))
            if y.shape[1] != self.n:
                self.close()
                raise ValueError(f"u0(p, t) returns {y.shape[1]} states, the device function expects {self.n}")
            B.check(B.lib().mirk_set_mesh_guess(self._h, len(mesh), _d(mesh), _d(y)))
        elif u0.ndim == 1:
            if u0.size != self.n:
                self.close()
                raise ValueError(f"u0 has {u0.size} states, the device function expects {self.n}")
            B.check(B.lib().mirk_set_uniform_guess(self._h, t0, t1, float(dt), _d(u0)))
        else:
            # the C ABI copies n_mesh * n doubles from the host buffer: check the shapes before it does
            if u0.ndim != 2 or u0.shape[1] != self.n or u0.shape[0] < 2:
                self.close()
                raise ValueError(f"a guess on a mesh must be an (N >= 2, {self.n}) array, got shape {u0.shape}")
            mesh = _arr(prob.mesh).ravel() if prob.mesh is not None else mesh_uniform(t0, t1, u0.shape[0] - 1)
            if len(mesh) != u0.shape[0]:
                self.close()
                raise ValueError(f"the guess has {u0.shape[0]} nodes, its mesh {len(mesh)}")
            B.check(B.lib().mirk_set_mesh_guess(self._h, len(mesh), _d(mesh), _d(u0)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            B.lib().mirk_destroy(self._h)
            self._h = B.Handle()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- pieces of a Newton step, for parity tests and the bench ------------------------------------
    @property
    def n_mesh(self) -> int:
        N = C.c_int32(0)
        B.check(B.lib().mirk_get_mesh_size(self._h, C.byref(N)))
        return N.value

    def residual(self):
        N = self.n_mesh
        out = np.zeros(self.info.n_bc + (N - 1) * self.n)
        nrm = C.c_double(0)
        B.check(B.lib().mirk_residual(self._h, _d(out), C.byref(nrm)))
        return out, nrm.value

    def jacobian_blocks(self):
        N, n, L = self.n_mesh, self.n, self.info.n_bc
        Lb, Rb = np.zeros((N - 1, n, n)), np.zeros((N - 1, n, n))
        nodes = np.zeros(16, dtype=np.int32)
        Bc = np.zeros((self.info.max_bc_pts, L, n))
        m = C.c_int32(0)
        B.check(B.lib().mirk_jacobian_blocks(self._h, _d(Lb), _d(Rb), _i(nodes), _d(Bc), C.byref(m)))
        return Lb, Rb, nodes[:m.value].copy(), Bc[:m.value].copy()

    def linear_solve(self):
        delta = np.zeros((self.n_mesh, self.n))
        st = B.check(B.lib().mirk_linear_solve(self._h, _d(delta)))
        return st, delta

    def newton_step(self):
        nrm = C.c_double(0)
        st = B.check(B.lib().mirk_newton_step(self._h, C.byref(nrm)))
        return st, nrm.value

    def newton_solve(self):
        it, nrm = C.c_int32(0), C.c_double(0)
        ret = B.check(B.lib().mirk_newton_solve(self._h, C.byref(it), C.byref(nrm)))
        return ret, it.value, nrm.value

    def nlsolve_stats(self):
        """(steps, retcodes) of NewtonRaphson / + BackTracking / TrustRegion in the last nonlinear solve (-1: did not run)"""
        st, rc = np.zeros(3, dtype=np.int32), np.zeros(3, dtype=np.int32)
        B.check(B.lib().mirk_nlsolve_stats(self._h, _i(st), _i(rc)))
        return list(map(int, st)), list(map(int, rc))

    def defect(self):
        N = self.n_mesh
        err = np.zeros((N - 1, self.n))
        d = C.c_double(0)
        B.check(B.lib().mirk_defect(self._h, _d(err), C.byref(d)))
        return d.value, err

    def refine_mesh(self):
        Nn = C.c_int32(0)
        info = B.check(B.lib().mirk_refine_mesh(self._h, C.byref(Nn)))
        return info, Nn.value

    def solution(self):
        N = self.n_mesh
        t, u = np.zeros(N), np.zeros((N, self.n))
        B.check(B.lib().mirk_get_solution(self._h, _d(t), _d(u)))
        return t, u

    def stages(self):
        N = self.n_mesh
        Kd = np.zeros((N - 1, self.s, self.n))
        Ki = np.zeros((N - 1, self.s_star - self.s, self.n))
        B.check(B.lib().mirk_get_stages(self._h, _d(Kd), _d(Ki)))
        return Kd, Ki

    def bench_newton_steps(self, steps: int):
        tot = C.c_float(0)
        ph = (C.c_float * 8)()
        launches = C.c_int64(0)
        st = B.check(B.lib().mirk_bench_newton_steps(self._h, int(steps), C.byref(tot), ph, C.byref(launches)))
        return st, tot.value, list(ph), launches.value


def abd_solve(Lb, Rb, bc_nodes, Bc, rhs, two_point=False, La=0, device=0):
    """The almost-block-diagonal solver on its own (what FIRK's expanded form and MIRKN share with MIRK): delta with
    [boundary rows; blockbidiag(Lb_i, Rb_i)] delta = rhs.  Lb, Rb: (N-1, n, n); Bc: (m, n, n) on nodes bc_nodes."""
    Lb, Rb, Bc, rhs = _arr(Lb), _arr(Rb), _arr(Bc), _arr(rhs)
    nm1, n, _ = Lb.shape
    nodes = np.ascontiguousarray(bc_nodes, dtype=np.int32)
    delta = np.zeros((nm1 + 1, n))
    st = B.check(B.lib().mirk_abd_solve(n, nm1 + 1, int(bool(two_point)), int(La), _d(Lb), _d(Rb), len(nodes), _i(nodes), _d(Bc),
                                        _d(rhs), _d(delta), int(device)))
    return st, delta


def mesh_uniform(t0, t1, nint) -> np.ndarray:
    """collect(range(t0; stop=t1, length=nint+1)) — CORE/utils.jl:694"""
    m = np.zeros(nint + 1)
    B.check(B.lib().mirk_mesh_uniform(float(t0), float(t1), int(nint), _d(m)))
    return m


def init(prob: BVProblem, alg: _AbstractMIRK, **kw) -> MIRKCache:
    """SciMLBase.__init(prob, alg; dt, abstol, adaptive, controller, nlsolve_kwargs, ...)"""
    return MIRKCache(prob, alg, **kw)


def solve_b(cache: MIRKCache) -> BVSolution:
    """SciMLBase.solve!(cache) (MIRK/src/mirk.jl:286-332)"""
    res = B.Result()
    B.check(B.lib().mirk_solve(cache._h, C.byref(res)))
    return BVSolution(cache, res)


# ---- ensembles -------------------------------------------------------------------------------------
@dataclass
class EnsembleProblem:
    """EnsembleProblem(prob; prob_func) with the 2-argument prob_func(prob, i) of this SciMLBase major
    (MIRK/test/Core/ensemble_tests.jl:20,37).  `params=` is the packed fast path: an (ntraj, np) array
    used instead of calling prob_func per trajectory."""
    prob: BVProblem
    prob_func: Optional[Callable] = None
    params: Optional[np.ndarray] = None


@dataclass(frozen=True)
class EnsembleB200:
    """Ensemble algorithm of this backend: all trajectories of a rank are solved by one batched kernel
    launch sequence on its GPU; with torch.distributed initialised the trajectories are block-partitioned
    over ranks (no data-path collective)."""
    device: Optional[int] = None


class EnsembleSolution:
    def __init__(self, retcodes, n_mesh, newton_iters, y_first, u=None, t=None, first=0):
        self.retcodes = retcodes
        self.n_mesh = n_mesh
        self.newton_iters = newton_iters
        self.y_first = y_first
        self.u, self.t = u, t
        self.first = first  # global index of the first trajectory this rank holds

    @property
    def converged(self) -> bool:
        return bool(np.all(self.retcodes == ReturnCode.Success))

    def __len__(self):
        return len(self.retcodes)


def solve(prob, alg: _AbstractMIRK, ensemblealg=None, trajectories: Optional[int] = None, **kw):
    """solve(prob::BVProblem, alg; dt, ...) or solve(ens::EnsembleProblem, alg, EnsembleB200(); trajectories, dt, ...)"""
    if isinstance(prob, EnsembleProblem):
        from .ensemble import solve_ensemble
        return solve_ensemble(prob, alg, ensemblealg or EnsembleB200(), trajectories, **kw)
    cache = init(prob, alg, **kw)
    return solve_b(cache)
