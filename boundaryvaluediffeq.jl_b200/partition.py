"""Mesh-partitioned single-problem mode: one BVP's mesh split into contiguous segments, one per GPU.

No counterpart in the reference (its hot path is one thread); SURVEY.md §8(e) defines it.  Residual,
Jacobian blocks and the block cyclic reduction are local to a segment; per Newton step the ranks
exchange one small reduced interface relation (NCCL all-gather inside libmirkb200, on the solver's
stream) and an 8-byte all-reduce of |F|_inf.  torch.distributed is used here only to hand the NCCL
unique id to every rank and to gather solutions for the caller.
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
    if prob.f.info.problem_type != 1:
        raise NotImplementedError("mesh partitioning needs a TwoPointBVProblem")
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
