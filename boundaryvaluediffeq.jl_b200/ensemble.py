"""EnsembleProblem over the batched MIRK kernel (one GPU thread per trajectory), sharded over ranks.

Mirrors `solve(EnsembleProblem(prob; prob_func), alg; trajectories, dt)` (SciMLBase driver; usage
lib/BoundaryValueDiffEqMIRK/test/Core/ensemble_tests.jl:20-38): the host calls prob_func per
trajectory to harvest parameters / initial states (or takes a packed `params=` array), then every
trajectory runs the complete adaptive solve on the device.  With torch.distributed initialised the
trajectories are block-partitioned over ranks — independent units, no data-path collective; only
the optional final gather of the per-trajectory outcomes communicates.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib as B
from .api import BVProblem, EnsembleProblem, EnsembleSolution, _AbstractMIRK, _arr, _d, _i, _nlsolve_code


def partition(ntraj: int, world: int) -> list:
    """Block partition of trajectories over ranks: [(first, count)] — sizes differ by at most one."""
    base, extra = divmod(int(ntraj), int(world))
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def harvest(ens: EnsembleProblem, trajectories: int) -> Tuple[np.ndarray, np.ndarray, bool]:
    """prob_func(prob, i) for i in 1..trajectories (1-based like the reference) -> packed (params, u0)."""
    prob = ens.prob
    n_p = len(prob.p)
    if ens.params is not None:
        params = _arr(ens.params).reshape(trajectories, -1)
        return params, _arr(prob.u0), False
    if ens.prob_func is None:
        return np.tile(_arr(prob.p), (trajectories, 1)).reshape(trajectories, n_p), _arr(prob.u0), False
    ps, u0s, per_traj = [], [], False
    for i in range(1, trajectories + 1):
        pi = ens.prob_func(prob, i)
        if not isinstance(pi, BVProblem):
            raise TypeError("prob_func must return a BVProblem (use prob.remake(p=...))")
        if pi.f.name != prob.f.name or tuple(pi.tspan) != tuple(prob.tspan):
            raise NotImplementedError("the batched ensemble needs one device function and one tspan for all trajectories")
        ps.append(_arr(pi.p))
        u0s.append(_arr(pi.u0))
        per_traj = per_traj or not np.array_equal(u0s[-1], _arr(prob.u0))
    params = np.stack(ps).reshape(trajectories, -1)
    u0 = np.stack(u0s) if per_traj else _arr(prob.u0)
    return params, u0, per_traj


def gather_outcomes(local: np.ndarray, counts: list, group=None) -> np.ndarray:
    """all_gather of a per-trajectory outcome array (first axis = this rank's trajectories)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(counts)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    pad = np.zeros((mx,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, counts)], axis=0)


class EnsembleHandle:
    """Owns one mirk_ensemble handle (device-resident inputs; used by bench.py for repeated runs)."""

    def __init__(self, prob: BVProblem, alg: _AbstractMIRK, ntraj: int, dt: float, abstol=1e-6, adaptive=True,
                 defect_threshold=0.1, maxiters=1000, device=0, node_cap=0, reinterp_inplace=False):
        if not (dt > 0):
            raise ValueError("dt must be positive")
        self.n = prob.f.info.n
        self.n_p = prob.f.info.n_params
        self.ntraj = int(ntraj)
        desc = B.EnsembleDesc(prob.f.problem_id, alg.order, float(abstol), int(bool(adaptive)), float(defect_threshold),
                              int(alg.max_num_subintervals), int(maxiters), int(reinterp_inplace), int(device),
                              int(node_cap), float(prob.tspan[0]), float(prob.tspan[1]), float(dt),
                              1 if _nlsolve_code(alg.nlsolve) == 1 else 0)
        self._h = B.Handle()
        B.check(B.lib().mirk_ensemble_create(C.byref(desc), self.ntraj, C.byref(self._h)))
        cap = C.c_int32(0)
        B.check(B.lib().mirk_ensemble_node_cap(self._h, C.byref(cap)))
        self.node_cap = int(cap.value)   # most nodes a trajectory's final mesh can have

    def set_inputs(self, params: np.ndarray, u0: np.ndarray, per_traj: bool = False):
        params = _arr(params).reshape(self.ntraj, -1)
        if params.shape[1] < self.n_p:
            raise ValueError("too few parameters per trajectory")
        params = np.ascontiguousarray(params[:, :max(self.n_p, 1)]) if self.n_p else params
        u0 = _arr(u0)
        B.check(B.lib().mirk_ensemble_set_inputs(self._h, _d(params) if self.n_p else None, _d(u0), int(per_traj)))

    def run(self) -> float:
        ms = C.c_float(0)
        B.check(B.lib().mirk_ensemble_run(self._h, C.byref(ms)))
        return ms.value

    def results(self):
        nt = self.ntraj
        ret, nm, its, outer = (np.zeros(nt, dtype=np.int32) for _ in range(4))
        rn, dn, yf = np.zeros(nt), np.zeros(nt), np.zeros((nt, self.n))
        B.check(B.lib().mirk_ensemble_get_results(self._h, _i(ret), _i(nm), _i(its), _i(outer), _d(rn), _d(dn), _d(yf)))
        return {"retcodes": ret, "n_mesh": nm, "newton_iters": its, "outer_iters": outer, "resid_norm": rn,
                "defect_norm": dn, "y_first": yf}

    def trajectory(self, i: int):
        N = C.c_int32(0)
        mesh, y = np.zeros(self.node_cap), np.zeros((self.node_cap, self.n))
        B.check(B.lib().mirk_ensemble_get_trajectory(self._h, int(i), C.byref(N), _d(mesh), _d(y)))
        return mesh[:N.value].copy(), y[:N.value].copy()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            B.lib().mirk_ensemble_destroy(self._h)
            self._h = B.Handle()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_ensemble(ens: EnsembleProblem, alg: _AbstractMIRK, ensemblealg, trajectories: Optional[int], dt: float = 0.0,
                   abstol: float = 1e-6, adaptive: bool = True, controller=None, nlsolve_kwargs=None, node_cap: int = 0,
                   gather: bool = False, keep_solutions: bool = False, **kw) -> EnsembleSolution:
    if trajectories is None:
        if ens.params is None:
            raise ValueError("trajectories must be given")
        trajectories = len(ens.params)
    if not (dt > 0):
        raise ValueError("dt must be positive")
    params, u0, per_traj = harvest(ens, trajectories)
    rank, world, device = 0, 1, ensemblealg.device
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    parts = partition(trajectories, world)
    first, count = parts[rank]
    if device is None:
        import os
        device = int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0
    nk = dict(nlsolve_kwargs or {})
    n = ens.prob.f.info.n
    u = t = None
    if count == 0:
        # more ranks than trajectories: this rank holds nothing and creates no handle, but still joins the gather
        res = {"retcodes": np.zeros(0, dtype=np.int32), "n_mesh": np.zeros(0, dtype=np.int32),
               "newton_iters": np.zeros(0, dtype=np.int32), "outer_iters": np.zeros(0, dtype=np.int32),
               "resid_norm": np.zeros(0), "defect_norm": np.zeros(0), "y_first": np.zeros((0, n))}
    else:
        h = EnsembleHandle(ens.prob, alg, count, dt, abstol=nk.get("abstol", abstol), adaptive=adaptive,
                           defect_threshold=(controller.defect_threshold if controller is not None else 0.1),
                           maxiters=nk.get("maxiters", 1000), device=device, node_cap=node_cap)
        sl = slice(first, first + count)
        h.set_inputs(params[sl], u0[sl] if per_traj else u0, per_traj)
        h.run()
        res = h.results()
        if keep_solutions:
            pairs = [h.trajectory(i) for i in range(count)]
            t, u = [a for a, _ in pairs], [b for _, b in pairs]
        h.close()
    if gather and world > 1:
        counts = [c for _, c in parts]
        res = {k: gather_outcomes(v, counts) for k, v in res.items()}
        first = 0
    return EnsembleSolution(res["retcodes"], res["n_mesh"], res["newton_iters"], res["y_first"], u=u, t=t, first=first)
