"""Mesh-partitioned single-problem mode: one BVP's mesh split into contiguous segments, one per GPU.

No counterpart in the reference (its hot path is one thread); SURVEY.md §8(e) defines it.  Residual,
Jacobian blocks and the block cyclic reduction are local to a segment; per Newton step the ranks
exchange one small reduced interface relation (peer-memory pushes over NVLink, or an NCCL all-gather inside
libmirkb200 on the solver's stream) and max-reduce three words.  torch.distributed is used here only to hand the
IPC handles / the NCCL unique id to every rank and to gather small host arrays for the caller.

`init_partitioned` gives the collective handle on a FIXED mesh; `solve_partitioned` runs the reference's adaptive outer
loop (DefectControl; MIRK/src/mirk.jl:286-388) around it: the defect estimate and the re-interpolation are local to a
segment, the mesh selector runs on the gathered per-interval estimates (mirk_mesh_select: the same kernel the
single-GPU driver uses), and every new mesh is re-partitioned into equal segments.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import List, Optional, Tuple

import numpy as np

from . import _lib as B
from .api import BVProblem, MIRKCache, _AbstractMIRK, _arr


def partition_mesh(n_nodes: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous node ranges [(lo, hi)] (inclusive) of `world` segments of a mesh with n_nodes nodes;
    neighbours share one boundary node, interval counts differ by at most one."""
    nint = int(n_nodes) - 1
    if world < 1 or nint < world:
        raise ValueError("need at least one interval per segment")
    base, extra = divmod(nint, world)
    out, lo = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((lo, lo + cnt))
        lo += cnt
    return out


def default_nccl_path() -> Optional[str]:
    """The NCCL torch ships (so one process never loads two NCCLs), else the system soname."""
    try:
        import torch
        hits = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
        if hits:
            return os.path.abspath(hits[0])
    except ImportError:
        pass
    return None


def _share_unique_id(group=None) -> bytes:
    import torch
    import torch.distributed as dist
    buf = (C.c_char * 128)()
    path = default_nccl_path()
    if dist.get_rank(group) == 0:
        B.check(B.lib().mirk_nccl_unique_id(C.cast(buf, C.c_void_p), path.encode() if path else None))
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def _gather_ipc_handles(mine: bytes, group=None) -> bytes:
    """all-gather of the 64-byte CUDA IPC handles, rank order"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).clone().to(dev)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t, group=group)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)


def init_partitioned(prob: BVProblem, alg: _AbstractMIRK, group=None, device: Optional[int] = None,
                     exchange: Optional[str] = None, **kw):
    """Collective.  `prob.u0` is the full (N, n) guess and `prob.mesh` the full mesh (every rank passes the
    same arrays); each rank keeps its segment on its GPU.  Returns (cache, (lo, hi)).

    exchange = "p2p" (default; MIRK_PART_XCHG overrides): the reduced interface relations travel by remote stores
    into peer memory from inside the pack kernel (CUDA IPC over NVLink, whole step graph-replayed);
    "nccl": one ncclAllGather + one ncclAllReduce per Newton step on the solver's stream."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    # (TwoPointBVProblems, and Standard problems whose boundary condition reads the two end points only: the library
    #  refuses anything else with MIRK_ERR_UNSUPPORTED)
    y, mesh = _arr(prob.u0), _arr(prob.mesh)
    lo, hi = partition_mesh(len(mesh), world)[rank]
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    local = BVProblem(prob.f, y[lo:hi + 1].copy(), (mesh[lo], mesh[hi]), p=prob.p, mesh=mesh[lo:hi + 1].copy())
    kw = dict(kw)
    kw["adaptive"] = False
    cache = MIRKCache(local, alg, device=device, **kw)
    exchange = exchange or os.environ.get("MIRK_PART_XCHG", "p2p")
    if exchange not in ("p2p", "nccl"):
        raise ValueError("exchange must be 'p2p' or 'nccl'")
    if exchange == "p2p":
        mine = (C.c_char * 64)()
        B.check(B.lib().mirk_partition_p2p_export(cache._h, rank, world, C.cast(mine, C.c_void_p)))
        allh = _gather_ipc_handles(bytes(mine), group)
        B.check(B.lib().mirk_partition_attach_p2p(cache._h, rank, world, C.cast(C.c_char_p(allh), C.c_void_p)))
    else:
        uid = _share_unique_id(group)
        path = default_nccl_path()
        B.check(B.lib().mirk_partition_attach(cache._h, rank, world, C.cast(C.c_char_p(uid), C.c_void_p),
                                              path.encode() if path else None))
    cache.exchange = exchange
    return cache, (lo, hi)


def gather_solution(cache: MIRKCache, n_nodes: int, group=None) -> np.ndarray:
    """All ranks' segments stitched back into the (N, n) solution (shared nodes taken from the left rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = partition_mesh(n_nodes, world)
    _, u = cache.solution()
    mx = max(hi - lo + 1 for lo, hi in parts)
    pad = np.zeros((mx, cache.n))
    pad[:u.shape[0]] = u
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    full = np.zeros((n_nodes, cache.n))
    for (lo, hi), o in reversed(list(zip(parts, outs))):
        full[lo:hi + 1] = o.cpu().numpy()[:hi - lo + 1]
    return full


class PartitionedSolution:
    """What solve_partitioned returns on every rank: sol.t, sol.u (N, n), sol.retcode, the per-outer-iteration
    histories mirk_result carries (hist_n_mesh, hist_newton, hist_defect) and the last norms."""

    def __init__(self, t, u, retcode, hist_n_mesh, hist_newton, hist_defect, resid_norm, defect_norm):
        self.t, self.u, self.retcode = t, u, retcode
        self.hist_n_mesh, self.hist_newton, self.hist_defect = hist_n_mesh, hist_newton, hist_defect
        self.resid_norm, self.defect_norm = resid_norm, defect_norm


def _allgather_obj(obj, group=None):
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def mesh_select(order: int, abstol: float, max_num_subintervals: int, mesh, est, device: int = 0):
    """mesh_selector! on a (global) mesh and its per-interval error estimates: (retcode, new mesh or None)."""
    mesh, est = _arr(mesh), _arr(est)
    out = np.zeros(max(len(mesh), int(max_num_subintervals) + 1))
    nn = C.c_int32(0)
    rc = B.check(B.lib().mirk_mesh_select(int(order), float(abstol), int(max_num_subintervals), len(mesh),
                                          mesh.ctypes.data_as(B.dp), est.ctypes.data_as(B.dp), C.byref(nn),
                                          out.ctypes.data_as(B.dp), int(device)))
    return rc, (out[:nn.value].copy() if rc == 0 else None)


def gather_estimates(defect_local: float, est_local, group=None):
    """(global defect norm, per-interval estimates of the whole mesh in rank order) from every rank's local ones; the
    maximum propagates NaN like the device reduction does."""
    parts = _allgather_obj((float(defect_local), np.asarray(est_local)), group)
    ds = [d for d, _ in parts]
    return (float("nan") if any(d != d for d in ds) else max(ds)), np.concatenate([e for _, e in parts])


def gather_rows(idx, vals, n_rows: int, group=None) -> np.ndarray:
    """Assemble an (n_rows, n) array from every rank's (row indices, rows): the new guess of the refined mesh."""
    vals = np.asarray(vals)
    out = np.zeros((n_rows, vals.shape[1]))
    for ii, vv in _allgather_obj((np.asarray(idx), vals), group):
        out[ii] = vv
    return out


def owned_nodes(mesh, lo: int, hi: int, rank: int, world: int, mesh_new) -> np.ndarray:
    """Which nodes of `mesh_new` the rank holding old nodes [lo, hi] re-interpolates: node t belongs to the rank whose
    segment contains the interval the reference's `interval(mesh, t)` = clamp(searchsortedfirst(mesh, t) - 1, 1, N - 1)
    picks (CORE/src/utils.jl:119-121), i.e. (mesh[lo], mesh[hi]]; t <= t0 goes to rank 0, anything beyond the last node
    to the last rank.  Every new node is owned by exactly one rank."""
    mesh, mesh_new = np.asarray(mesh), np.asarray(mesh_new)
    mine = (mesh_new > mesh[lo]) & (mesh_new <= mesh[hi])
    if rank == 0:
        mine |= mesh_new <= mesh[lo]
    if rank == world - 1:
        mine |= mesh_new > mesh[hi]
    return mine


def solve_partitioned(prob: BVProblem, alg: _AbstractMIRK, dt: float = 0.0, abstol: float = 1e-6, adaptive: bool = True,
                      defect_threshold: float = 0.1, group=None, device: Optional[int] = None,
                      exchange: Optional[str] = None, max_outer: int = 1000, **kw) -> PartitionedSolution:
    """Collective.  `solve(prob, alg; dt, abstol, adaptive)` of a BVProblem (two-point, or Standard with end-point boundary
    conditions) whose mesh is partitioned over the
    ranks of `group`: the outer loop of `solve!` with DefectControl (MIRK/src/mirk.jl:286-388; same order of
    operations as mirk_solve in csrc/mirk_b200.cu).  Every rank passes the same problem and gets the same result."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    t0, t1 = float(prob.tspan[0]), float(prob.tspan[1])
    if prob.mesh is not None:
        mesh, y = _arr(prob.mesh).copy(), _arr(prob.u0).copy()
    else:
        if not dt > 0.0:
            raise ValueError("dt must be positive")
        from .api import mesh_uniform
        mesh = mesh_uniform(t0, t1, int(np.ceil((t1 - t0) / dt)))
        y = np.tile(_arr(prob.u0).reshape(1, -1), (len(mesh), 1))
    max_sub = int(alg.max_num_subintervals)
    cache = None
    hist_n, hist_it, hist_d = [], [], []
    info, err_norm, resid_norm = 0, 2.0 * abstol, float("nan")
    try:
        while True:
            N = len(mesh)
            lo, hi = partition_mesh(N, world)[rank]
            if cache is None:
                full = BVProblem(prob.f, y, (t0, t1), p=prob.p, mesh=mesh)
                cache, _ = init_partitioned(full, alg, group=group, device=device, exchange=exchange, abstol=abstol, **kw)
            else:
                ms, ys = np.ascontiguousarray(mesh[lo:hi + 1]), np.ascontiguousarray(y[lo:hi + 1])
                B.check(B.lib().mirk_set_mesh_guess(cache._h, hi - lo + 1, ms.ctypes.data_as(B.dp), ys.ctypes.data_as(B.dp)))
            info, iters, resid_norm = cache.newton_solve()
            solved_mesh = mesh
            err_norm = 2.0 * abstol
            hist_n.append(N); hist_it.append(iters); hist_d.append(float("nan"))
            if not adaptive or len(hist_n) >= max_outer:
                break
            if info == 0:
                d_loc, errs = cache.defect()
                est_loc = np.max(np.abs(errs), axis=1) if errs.size else np.zeros(0)
                err_norm, est = gather_estimates(d_loc, est_loc, group)
                if not (err_norm <= defect_threshold):
                    info = 1
                hist_d[-1] = err_norm
                if info == 0 and err_norm > abstol:
                    rc, mesh_new = mesh_select(alg.order, abstol, max_sub, mesh, est, device)
                    if rc != 0:
                        info = rc
                        break
                    # new guess = old interpolant at the new nodes, each evaluated by the rank that owns its old interval
                    idx = np.nonzero(owned_nodes(mesh, lo, hi, rank, world, mesh_new))[0]
                    ts = np.ascontiguousarray(mesh_new[idx])
                    vals = np.zeros((len(ts), cache.n))
                    if len(ts):
                        B.check(B.lib().mirk_interp(cache._h, ts.ctypes.data_as(B.dp), len(ts), 0, vals.ctypes.data_as(B.dp)))
                    mesh, y = mesh_new, gather_rows(idx, vals, len(mesh_new), group)
                    continue
            if info != 0:
                if 2 * (N - 1) > max_sub:
                    info = 1
                    break
                half = np.empty(2 * N - 1)
                half[0::2] = mesh
                half[1::2] = (mesh[1:] + mesh[:-1]) / 2.0
                mesh, y = half, np.zeros((2 * N - 1, cache.n))   # halve-and-zero restart (quirk Q4)
                info = 0
            if not (info == 0 and err_norm > abstol):
                break
        if info == 0 and adaptive and err_norm > abstol:
            info = 2  # MaxIters (the safety net of the outer loop)
        u = gather_solution(cache, len(solved_mesh), group)
        return PartitionedSolution(solved_mesh, u, info, hist_n, hist_it, hist_d, resid_norm, err_norm)
    finally:
        if cache is not None:
            dist.barrier(group=group)  # nobody frees its exchange buffer while a peer may still push into it
            cache.close()
