"""GPU parity of the batched ensemble kernel (one thread per trajectory) against the CPU oracle's
ensemble driver, trajectory by trajectory: same return codes, final mesh sizes, Newton counts, and
solution values within 1e-10 relative."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mirk_b200 as m
    return m


def _oracle_ensemble(O, name, order, params, u0, tspan, dt, **kw):
    nint = int(math.ceil((tspan[1] - tspan[0]) / dt))
    return O.ensemble_solve(O.builtin(name), order, params, u0, tspan, nint, nthreads=8, **kw)


@pytest.mark.parametrize("order", [4, 6])
def test_pendulum_sweep_matches_oracle(M, oracle, order):
    """BASELINE config C3 at a size the oracle finishes in seconds: g/L ~ U(8, 12), MIRK4/6, dt = 0.05"""
    from boundaryvaluediffeq_jl_b200 import configs
    params = configs.c3_ensemble_params(300)
    u0, tspan = [math.pi / 2, math.pi / 2], (0.0, math.pi / 2)
    ret, Nf, y0, its = _oracle_ensemble(oracle, "pendulum", order, params, u0, tspan, 0.05)
    alg = M.MIRK4() if order == 4 else M.MIRK6()
    ens = M.EnsembleProblem(M.BVProblem("pendulum", u0, tspan, p=[9.81]), params=params)
    sol = M.solve(ens, alg, M.EnsembleB200(), trajectories=300, dt=0.05, keep_solutions=True)
    assert sol.converged and np.all(ret == 0)
    assert np.array_equal(sol.retcodes, ret)
    assert np.array_equal(sol.n_mesh, Nf)
    assert np.array_equal(sol.newton_iters, its)
    assert np.max(np.abs(sol.y_first - y0)) < 1e-10 * np.max(np.abs(y0))
    # full solution of a few trajectories against single oracle solves
    for i in (0, 17, 299):
        ref = oracle.solve_dt(oracle.builtin("pendulum"), order, params[i], u0, tspan, 0.05)
        assert len(sol.t[i]) == ref.N
        assert np.max(np.abs(sol.t[i] - ref.t)) < 1e-10
        assert np.max(np.abs(sol.u[i] - ref.u)) < 1e-10 * np.max(np.abs(ref.u))


def test_reference_ensemble_test_with_prob_func(M, oracle):
    """lib/BoundaryValueDiffEqMIRK/test/Core/ensemble_tests.jl:7-40: u'' = -p u, bc u(0) = 1, u(1) = 0 sampled at
    interior time 1.0 of tspan (0, pi/2); prob_func swaps p; 10 trajectories must all converge."""
    tspan = (0.0, math.pi / 2)
    base = M.BVProblem("linear2", [0.0, 1.0], tspan, p=[1.0, 0.0, 1.0, 1.0, 0.0, 0, 0])
    ps = np.linspace(0.5, 5.0, 10)

    def prob_func(prob, i):
        p = prob.p.copy()
        p[0] = ps[i - 1]
        return prob.remake(p=p)

    for alg, order in ((M.MIRK4(), 4), (M.MIRK6(), 6)):
        sol = M.solve(M.EnsembleProblem(base, prob_func=prob_func), alg, trajectories=10, dt=0.1)
        assert sol.converged and len(sol) == 10
        params = np.stack([prob_func(base, i).p for i in range(1, 11)])
        ret, Nf, y0, its = _oracle_ensemble(oracle, "linear2", order, params, [0.0, 1.0], tspan, 0.1)
        assert np.array_equal(sol.retcodes, ret) and np.array_equal(sol.n_mesh, Nf)
        assert np.array_equal(sol.newton_iters, its)
        assert np.max(np.abs(sol.y_first - y0)) < 1e-10 * max(1.0, np.max(np.abs(y0)))


@pytest.mark.parametrize("name,order,p,u0,tspan,dt,kw", [
    ("layer", 4, [0.01], [0.0, 0.0], (-1.0, 1.0), 0.05, {"node_cap": 512}),      # redistribution, 3 outer iterations
    ("lotka", 4, [7.5, 4.0, 8.5, 5.0], [1.0, 2.0], (0.0, 10.0), 0.1, {"node_cap": 1024}),  # big defect
    ("torus", 4, [4.0, 1.0, 0.0, 0.0, 1.0, 2.0], [0.0, 0.0, 1.0, 2.0], (0.0, 1.0), 0.05, {}),  # n = 4
    ("swirling", 4, [0.01], [0.0] * 6, (0.0, 1.0), 0.01, {"abstol": 1e-4}),       # n = 6
    # singular term (y' = S y / t + f) through the warp kernel's stage-wise Jacobian; fixed mesh (see test_gpu_parity.py)
    ("lane_emden", 4, [], [1.0, 0.0], (0.0, 1.0), 0.01, {"adaptive": False}),
    ("lane_emden", 6, [], [1.0, 0.0], (0.0, 1.0), 0.02, {"adaptive": False}),
])
def test_single_trajectory_ensembles_follow_the_single_solve_path(M, oracle, name, order, p, u0, tspan, dt, kw):
    """Each thread runs the same adaptive loop as mirk_solve: compare a few-trajectory ensemble of identical
    problems with the oracle's single solve (mesh history end point, Newton count, values)."""
    kw = dict(kw)
    node_cap = kw.pop("node_cap", 0)
    ref = oracle.solve_dt(oracle.builtin(name), order, p, u0, tspan, dt, **kw)
    params = np.tile(np.asarray(p, dtype=float), (3, 1))
    ens = M.EnsembleProblem(M.BVProblem(name, u0, tspan, p=p), params=params)
    sol = M.solve(ens, M.MIRK4() if order == 4 else M.MIRK6(), trajectories=3, dt=dt, node_cap=node_cap,
                  keep_solutions=True, **kw)
    assert list(sol.retcodes) == [ref.retcode] * 3
    assert list(sol.n_mesh) == [ref.N] * 3
    assert list(sol.newton_iters) == [ref.newton_iters] * 3
    # the boundary-layer system (eps = 0.01) is ill-conditioned: one Newton step leaves |F| ~ 1e-9, and two
    # elimination orders then differ by cond * eps; everything else meets the 1e-10 north-star tolerance
    tol = 1e-8 if name in ("layer", "lotka") else 1e-10
    assert np.max(np.abs(sol.u[1] - ref.u)) < tol * max(1.0, np.max(np.abs(ref.u)))


def test_node_capacity_exhaustion_is_a_failure_not_a_crash(M):
    params = np.tile([0.01], (2, 1))
    ens = M.EnsembleProblem(M.BVProblem("layer", [0.0, 0.0], (-1.0, 1.0), p=[0.01]), params=params)
    sol = M.solve(ens, M.MIRK4(), trajectories=2, dt=0.05, node_cap=64)
    assert not sol.converged and list(sol.retcodes) == [M.ReturnCode.Failure] * 2


def test_full_size_sweep_properties(M):
    """BASELINE config C3 at full size (262 144 pendulum BVPs): every trajectory converges, the boundary
    conditions hold at the returned nodes, and a permutation of the parameters permutes the results."""
    from boundaryvaluediffeq_jl_b200 import configs, ensemble as E
    nt = 262144
    params = configs.c3_ensemble_params(nt)
    prob = M.BVProblem("pendulum", [math.pi / 2, math.pi / 2], (0.0, math.pi / 2), p=[9.81])
    h = E.EnsembleHandle(prob, M.MIRK4(), nt, 0.05)
    h.set_inputs(params, prob.u0)
    ms = h.run()
    r = h.results()
    assert np.all(r["retcodes"] == 0) and np.all(r["resid_norm"] <= 1e-6) and np.all(r["defect_norm"] <= 1e-6)
    mesh, y = h.trajectory(nt - 1)
    assert abs(y[-1, 0] - math.pi / 2) < 1e-6
    perm = np.random.default_rng(0).permutation(nt)
    h.set_inputs(params[perm], prob.u0)
    h.run()
    r2 = h.results()
    assert np.array_equal(r2["n_mesh"], r["n_mesh"][perm]) and np.array_equal(r2["y_first"], r["y_first"][perm])
    h.close()
    print(f"262144 pendulum BVPs in {ms:.1f} ms")


def test_one_shot_c_abi_entry_point(M, oracle):
    """mirk_ensemble_solve (create + set_inputs + run + get_results + destroy in one call), bound directly."""
    import ctypes as C
    from boundaryvaluediffeq_jl_b200 import _lib as B, configs
    nt = 64
    params = np.ascontiguousarray(configs.c3_ensemble_params(nt))
    u0 = np.array([math.pi / 2, math.pi / 2])
    desc = B.EnsembleDesc(M.BVPDeviceFunction("pendulum").problem_id, 4, 1e-6, 1, 0.1, 3000, 1000, 0, 0, 0,
                          0.0, math.pi / 2, 0.05)
    ret, nm, its = (np.zeros(nt, dtype=np.int32) for _ in range(3))
    yf = np.zeros((nt, 2))
    d = lambda a: a.ctypes.data_as(B.dp)   # noqa: E731
    i = lambda a: a.ctypes.data_as(B.ip)   # noqa: E731
    B.check(B.lib().mirk_ensemble_solve(C.byref(desc), nt, d(params), d(u0), 0, i(ret), i(nm), i(its), d(yf)))
    r_ret, r_nm, r_y0, r_its = oracle.ensemble_solve(oracle.builtin("pendulum"), 4, params, u0, (0.0, math.pi / 2), 32, nthreads=4)
    assert np.array_equal(ret, r_ret) and np.array_equal(nm, r_nm) and np.array_equal(its, r_its)
    assert np.max(np.abs(yf - r_y0)) < 1e-10 * np.max(np.abs(r_y0))
    # argument errors come back as status codes
    bad = B.EnsembleDesc(desc.problem_id, 4, 1e-6, 1, 0.1, 3000, 1000, 0, 0, 0, 0.0, 1.0, 0.0)
    with pytest.raises(ValueError, match="dt must be positive"):
        B.check(B.lib().mirk_ensemble_solve(C.byref(bad), nt, d(params), d(u0), 0, i(ret), i(nm), i(its), d(yf)))


@pytest.mark.parametrize("name,order,p,u0,tspan,dt", [
    ("lotka", 4, [7.5, 4.0, 8.5, 5.0], [5.0, 5.0], (0.0, 10.0), 0.1),                           # warp kernel (n = 2)
    ("torus", 4, [1.2, 1.0, 0.0, 0.0, 3.0, 5.0], [3.0, 0.0, 10.0, -20.0], (0.0, 1.0), 0.05),     # thread kernel (n = 4)
])
def test_trajectories_that_need_the_polyalgorithm_are_rerun_by_the_single_driver(M, oracle, name, order, p, u0, tspan, dt):
    """Plain Newton fails on these: under the default polyalgorithm the batched kernel hands such a trajectory to the
    single-problem driver (line-search / trust-region fallbacks), so the ensemble gives what the oracle's ensemble
    (every trajectory through the full polyalgorithm) gives."""
    ref = oracle.solve_dt(oracle.builtin(name), order, p, u0, tspan, dt)
    params = np.tile(np.asarray(p, dtype=float), (2, 1))
    ens = M.EnsembleProblem(M.BVProblem(name, u0, tspan, p=p), params=params)
    sol = M.solve(ens, M.MIRK4(), trajectories=2, dt=dt, keep_solutions=True)
    assert list(sol.retcodes) == [ref.retcode] * 2
    assert list(sol.n_mesh) == [ref.N] * 2
    # (total step counts include the diverging sub-solvers, whose counts are not reproducible across linear solvers)
    assert np.max(np.abs(sol.u[1] - ref.u)) < 1e-7 * max(1.0, np.max(np.abs(ref.u)))
