"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, argument
errors surface without a GPU, the ensemble sharding logic, and the world_size-2 gloo gather."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_in_the_header():
    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200 import _lib as B
    header = open(os.path.join(ROOT, "include", "mirk_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(mirk_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    lib = M.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mirk_b200.h but not exported"
    assert declared == set(B.SYMBOLS), (declared ^ set(B.SYMBOLS))


def test_no_device_is_an_error_not_a_fallback():
    import ctypes as C

    import mirk_b200 as M
    cnt = C.c_int32(0)
    if M.lib().mirk_device_count(C.byref(cnt)) == 0 and cnt.value > 0:
        pytest.skip("a CUDA device is present")
    prob = M.BVProblem("pendulum", [1.5, 1.5], (0.0, 1.5), p=[9.81])
    with pytest.raises(M.MirkError) as ei:
        M.solve(prob, M.MIRK4(), dt=0.05)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_registry_and_host_helpers():
    import mirk_b200 as M
    from oracle import oracle as O
    for name, pid in O.PROBLEM_IDS.items():
        f = M.BVPDeviceFunction(name)
        assert f.problem_id == pid
        P = O.builtin(name)
        assert (f.info.n, f.info.n_params, f.info.problem_type, f.info.n_bc, f.info.n_bca) == \
               (P.n, P.n_p, P.problem_type, P.n_bc, P.n_bca)
    with pytest.raises(M.MirkError):
        M.BVPDeviceFunction("nope").problem_id
    for t0, t1, nint in [(0.0, np.pi / 2, 32), (0.0, 0.5, 19999), (-1.0, 1.0, 7)]:
        assert np.array_equal(M.mesh_uniform(t0, t1, nint), O.mesh_uniform(t0, t1, nint))
    with pytest.raises(NotImplementedError):
        M.MIRK4(nlsolve="NewtonRaphson")
    # the sub-solvers of the reference's default polyalgorithm can be requested on their own
    assert M.MIRK4(nlsolve=M.NewtonRaphson()).nlsolve.linesearch is None
    M.MIRK4(nlsolve=M.NewtonRaphson(linesearch=M.BackTracking()))
    M.MIRK6(nlsolve=M.TrustRegion())


def test_partition_covers_every_trajectory_once():
    from boundaryvaluediffeq_jl_b200.ensemble import partition
    import mirk_b200  # noqa: F401
    for nt in (1, 7, 10, 262144, 262145):
        for world in (1, 2, 3, 8):
            parts = partition(nt, world)
            assert len(parts) == world and parts[0][0] == 0
            assert sum(c for _, c in parts) == nt
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_mesh_partition_segments_share_exactly_their_boundary_nodes():
    import mirk_b200  # noqa: F401
    from boundaryvaluediffeq_jl_b200.partition import partition_mesh
    for N in (9, 20000, 2000001):
        for world in (1, 2, 4, 8):
            parts = partition_mesh(N, world)
            assert parts[0][0] == 0 and parts[-1][1] == N - 1
            for (lo0, hi0), (lo1, hi1) in zip(parts, parts[1:]):
                assert hi0 == lo1
            sizes = [hi - lo for lo, hi in parts]
            assert sum(sizes) == N - 1 and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition_mesh(3, 4)


def test_harvest_calls_prob_func_one_based():
    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200.ensemble import harvest
    base = M.BVProblem("linear2", [0.0, 1.0], (0.0, 1.0), p=[1.0, 0.0, 1.0, 1.0, 0.0, 0, 0])
    seen = []

    def prob_func(prob, i):
        seen.append(i)
        p = prob.p.copy()
        p[0] = float(i)
        return prob.remake(p=p)

    params, u0, per = harvest(M.EnsembleProblem(base, prob_func=prob_func), 4)
    assert seen == [1, 2, 3, 4] and list(params[:, 0]) == [1.0, 2.0, 3.0, 4.0] and not per
    assert np.array_equal(base.p, [1.0, 0.0, 1.0, 1.0, 0.0, 0, 0])  # prob_func must not mutate the base problem


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
import mirk_b200  # noqa
from boundaryvaluediffeq_jl_b200.ensemble import partition, gather_outcomes
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
nt = 11
parts = partition(nt, 2)
first, count = parts[dist.get_rank()]
local = (np.arange(first, first + count, dtype=np.int32) * 3)          # stands for per-trajectory outcomes
local2 = np.stack([np.arange(first, first + count, dtype=np.float64)] * 2, axis=1)
g = gather_outcomes(local, [c for _, c in parts])
g2 = gather_outcomes(local2, [c for _, c in parts])
assert np.array_equal(g, np.arange(nt, dtype=np.int32) * 3), g
assert g2.shape == (nt, 2) and np.array_equal(g2[:, 0], np.arange(nt))
dist.barrier()
dist.destroy_process_group()
print("ok", dist_rank := int(sys.argv[1]))
"""


def test_two_rank_gloo_shard_and_gather(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def test_dmma_fragment_merge_layout_matches_sequential_elimination():
    """abd_mma.cuh's data movement (fragment layout, panel gather, published pivot rows, rank-4 DMMA updates)
    restated lane by lane in numpy (experiments/mma_merge_emul.py) reproduces a plain row-pivoted Gauss-Jordan
    elimination: same pivot rows, same surviving rows, same reciprocals."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mma_merge_emul", os.path.join(ROOT, "experiments", "mma_merge_emul.py"))
    emul = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emul)
    rng = np.random.default_rng(11)
    for _ in range(3):
        W = rng.standard_normal((32, 48))
        rhs = rng.standard_normal(32)
        W[:16, 32:] = 0.0   # carried rows have no B part
        W[16:, 16:32] = 0.0  # incoming rows have no A part
        Wa, ra, qa, ia = emul.merge_seq(W, rhs)
        Wb, rb, qb, ib = emul.merge_mma(W, rhs)
        assert (qa == qb).all()
        assert np.abs(Wa[:, 16:] - Wb[:, 16:]).max() < 1e-12
        assert np.abs(ra - rb).max() < 1e-12 and np.abs(ia - ib).max() < 1e-14


def test_dmma_fragment_merge_n32_four_warps_matches_sequential_elimination():
    """csrc/abd_mma32.cuh's data movement (four warps per merge, warp-local panels, block-wide pivot records and
    pivot rows) restated thread by thread in numpy (experiments/mma32_merge_emul.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mma32_merge_emul", os.path.join(ROOT, "experiments", "mma32_merge_emul.py"))
    emul = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emul)
    rng = np.random.default_rng(12)
    W = rng.standard_normal((64, 96))
    rhs = rng.standard_normal(64)
    W[:32, 64:] = 0.0
    W[32:, 32:64] = 0.0
    Wa, ra, qa, ia = emul.merge_seq(W, rhs)
    Wb, rb, qb, ib = emul.merge_mma32(W, rhs)
    assert (qa == qb).all()
    assert np.abs(Wa[:, 32:] - Wb[:, 32:]).max() < 1e-11
    assert np.abs(ra - rb).max() < 1e-11 and np.abs(ia - ib).max() < 1e-13


# ---- struct layouts across the C ABI: header (gcc) vs ctypes mirror vs the Julia glue's struct definitions ------
def _c_layout(tmp_path):
    exe = os.path.join(str(tmp_path), "c_abi_layout")
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_layout.c"),
                    "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return {k: int(v) for k, v in (ln.split() for ln in out.strip().splitlines())}


_JL_TYPES = {"Int32": (4, 4), "Float64": (8, 8), "Ptr{Float64}": (8, 8), "Ptr{Cvoid}": (8, 8), "Int64": (8, 8)}


def _julia_layout(struct_name):
    """(size, {field: offset}) of an immutable Julia struct of C-compatible fields, by the C layout rules Julia's
    `ccall` follows for isbits structs (natural alignment, NTuple{N,T} = T[N])."""
    src = open(os.path.join(ROOT, "julia", "BoundaryValueDiffEqMIRKB200", "src", "BoundaryValueDiffEqMIRKB200.jl")).read()
    m = re.search(r"^struct %s\n(.*?)^end" % struct_name, src, flags=re.M | re.S)
    assert m, f"struct {struct_name} not found in the Julia glue"
    off, offs, maxal = 0, {}, 1
    for ln in m.group(1).strip().splitlines():
        ln = ln.split("#")[0].strip()
        if not ln:
            continue
        name, typ = [x.strip() for x in ln.split("::")]
        nt = re.match(r"NTuple\{(\d+),\s*(\w+)\}", typ)
        if nt:
            sz, al = _JL_TYPES[nt.group(2)]
            sz *= int(nt.group(1))
        else:
            sz, al = _JL_TYPES[typ]
        off = (off + al - 1) // al * al
        offs[name] = off
        off += sz
        maxal = max(maxal, al)
    return (off + maxal - 1) // maxal * maxal, offs


def test_struct_layouts_agree_between_header_ctypes_and_julia(tmp_path):
    import mirk_b200  # noqa: F401
    from boundaryvaluediffeq_jl_b200 import _lib as B
    c = _c_layout(tmp_path)
    pairs = [("mirk_desc", B.Desc, "MirkDesc"), ("mirk_problem_info", B.ProblemInfo, "MirkProblemInfo"),
             ("mirk_result", B.Result, "MirkResult"), ("mirk_ensemble_desc", B.EnsembleDesc, "MirkEnsembleDesc")]
    import ctypes as C
    for cname, ct, jl in pairs:
        fields = {k.split(".", 1)[1]: v for k, v in c.items() if k.startswith(cname + ".")}
        assert fields, cname
        assert C.sizeof(ct) == c["sizeof." + cname], cname
        assert [f for f, _ in ct._fields_] == list(fields), (cname, "field order")
        for f, _ in ct._fields_:
            assert getattr(ct, f).offset == fields[f], (cname, f)
        jsize, joffs = _julia_layout(jl)
        assert jsize == c["sizeof." + cname], (jl, jsize)
        assert joffs == fields, (jl, joffs, fields)


def test_partitioned_reinterpolation_ownership_follows_the_reference_interval_rule():
    """partition.owned_nodes: every node of a refined mesh is re-interpolated by exactly one rank, the one whose segment
    holds interval(mesh, t) = clamp(searchsortedfirst(mesh, t) - 1, 1, N - 1) (CORE/src/utils.jl:119-121) — a new node that
    coincides with an old one belongs to the interval on its LEFT, hence to the left rank at a segment boundary."""
    from boundaryvaluediffeq_jl_b200 import partition
    rng = np.random.default_rng(3)
    mesh = np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0, 1, 37)]))
    new = np.sort(np.concatenate([mesh[::3], rng.uniform(0, 1, 80), [0.0, 1.0]]))
    for world in (1, 2, 3, 5):
        parts = partition.partition_mesh(len(mesh), world)
        owners = np.zeros(len(new), dtype=int)
        for r, (lo, hi) in enumerate(parts):
            m = partition.owned_nodes(mesh, lo, hi, r, world, new)
            owners += m
            # the reference's interval of every owned node lies inside this rank's segment
            iv = np.clip(np.searchsorted(mesh, new[m], side="left") - 1, 0, len(mesh) - 2)
            assert np.all((iv >= lo) & (iv + 1 <= hi))
        assert np.all(owners == 1)


_PART_WORKER = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
import mirk_b200  # noqa
from boundaryvaluediffeq_jl_b200 import partition
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)                                    # same numbers on both ranks
mesh = np.sort(np.concatenate([[0.0, 2.0], rng.uniform(0, 2, 40)]))
new = np.sort(np.concatenate([[0.0, 2.0], mesh[5::7], rng.uniform(0, 2, 61)]))
lo, hi = partition.partition_mesh(len(mesh), world)[rank]
# the adaptive step of solve_partitioned with a stand-in interpolant u(t) = (t, t^2): local estimates, global selector
# input, local re-interpolation of the owned nodes, assembled new guess
est_local = np.abs(np.sin(mesh[lo:hi]))                           # one entry per local interval
dn, est = partition.gather_estimates(float(est_local.max()), est_local)
assert est.shape == (len(mesh) - 1,) and np.array_equal(est, np.abs(np.sin(mesh[:-1]))) and dn == est.max()
dn_nan, _ = partition.gather_estimates(float("nan") if rank == 1 else 1.0, est_local)
assert dn_nan != dn_nan
idx = np.nonzero(partition.owned_nodes(mesh, lo, hi, rank, world, new))[0]
vals = np.stack([new[idx], new[idx] ** 2], axis=1)
y_new = partition.gather_rows(idx, vals, len(new))
assert np.array_equal(y_new[:, 0], new) and np.array_equal(y_new[:, 1], new ** 2)
dist.barrier()
dist.destroy_process_group()
print("ok")
"""


def test_two_rank_gloo_partitioned_adaptive_host_logic(tmp_path):
    """The host side of the mesh-partitioned adaptive loop (partition.gather_estimates / owned_nodes / gather_rows) on
    two gloo ranks: the gathered estimates are the whole mesh's in rank order, a NaN defect on one rank is a NaN
    everywhere, and the re-interpolated rows of both ranks assemble into the complete new guess."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "part_worker.py"
    script.write_text(_PART_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def test_bench_work_model_partitions_the_survey_byte_and_flop_counts():
    """B = 8 [2 n N + N + 2 * 2 n^2 (N - 1)] bytes and F = F_f + F_J + F_S flops per Newton step.  The per-kernel
    figures bench.py divides by the measured kernel times must SUM to those totals (C2: 169.1 MB, ~1.43 GF; C4 ~73 GF;
    C5 ~1.1 TF, 66 GB) — no kernel may be credited traffic the table does not list."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for n, N, order, B_expect, F_lo, F_hi in ((16, 20000, 6, 169.1e6, 1.35e9, 1.50e9), (128, 4000, 4, 2.10e9, 70e9, 76e9),
                                              (32, 2000000, 6, 66.6e9, 1.05e12, 1.15e12)):
        by, fl, B, F = bench.work_model(n, N, order)
        assert B == 8 * (2 * n * N + N + 2 * 2 * n * n * (N - 1))
        assert sum(by.values()) == B and abs(B - B_expect) / B_expect < 5e-3
        assert abs(sum(fl.values()) - F) / F < 1e-12 and F_lo < F < F_hi
        assert set(by) == set(bench.PHASES) == set(fl)
        # the dominant kernel (level-0 reduction) is credited the READ of the blocks only
        assert by["abd_reduce_level0"] == 8 * 2 * n * n * (N - 1)
