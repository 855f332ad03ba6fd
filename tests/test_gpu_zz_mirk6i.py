"""GPU parity for MIRK6I (`order` code 7 of the C ABI): the irrational 6th-order tableau through the
same kernels as MIRK6, against the CPU oracle.  Kept in its own file, last in collection order."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PENDULUM_T = (0.0, math.pi / 2)
LIN_P = [1.0, 0.0, 5.0, 5.0, 0.0, 0, 0]


@pytest.fixture(scope="module")
def M():
    import mirk_b200 as m
    return m


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


@pytest.mark.parametrize("name,p,tspan,nint", [
    ("pendulum", [9.81], PENDULUM_T, 32), ("linear2_tp", [1.0, 5.0, 0.0], (0.0, 5.0), 17),
    ("swirling", [0.01], (0.0, 1.0), 23), ("torus", [4.0, 1.0, 0.0, 0.0, 1.0, 2.0], (0.0, 1.0), 20),
])
def test_mirk6i_residual_jacobian_and_update_match_oracle(M, oracle, name, p, tspan, nint):
    O = oracle
    rng = np.random.default_rng(nint)
    P = O.builtin(name)
    mesh = np.asarray(O.mesh_uniform(tspan[0], tspan[1], nint))
    mesh[1:-1] += rng.uniform(-0.3, 0.3, nint - 1) * (tspan[1] - tspan[0]) / nint
    y = 0.3 * rng.standard_normal((nint + 1, P.n))
    ws = O.Workspace(P, O.MIRK6I, p, mesh, y)
    cache = M.init(M.BVProblem(name, y, tspan, p=p, mesh=mesh), M.MIRK6I(), adaptive=False)
    r, nrm = cache.residual()
    r_ref = ws.loss()
    assert _rel(r, r_ref) < 1e-13
    Lb, Rb, nodes, Bc = cache.jacobian_blocks()
    Lr, Rr = ws.jac_blocks()
    assert _rel(Lb, Lr) < 1e-12 and _rel(Rb, Rr) < 1e-12
    st, delta = cache.linear_solve()
    assert st == 0
    d_ref = np.linalg.solve(ws.dense_jacobian(), r_ref).reshape(delta.shape)
    assert _rel(delta, d_ref) < (1e-7 if name == "swirling" else 1e-9)
    cache.close()


@pytest.mark.parametrize("name,p,u0,tspan,dt,kw", [
    ("pendulum", [9.81], [math.pi / 2, math.pi / 2], PENDULUM_T, 0.05, {}),
    ("linear2", LIN_P, [5.0, -3.5], (0.0, 5.0), 0.2, {}),
    ("swirling", [0.01], [0.0] * 6, (0.0, 1.0), 0.01, {"abstol": 1e-4}),
])
def test_mirk6i_full_solve_matches_oracle(M, oracle, name, p, u0, tspan, dt, kw):
    O = oracle
    ref = O.solve_dt(O.builtin(name), O.MIRK6I, p, u0, tspan, dt, **kw)
    sol = M.solve(M.BVProblem(name, u0, tspan, p=p), M.MIRK6I(), dt=dt, **kw)
    assert sol.retcode == ref.retcode == 0
    assert sol.original["hist_n_mesh"] == ref.hist_N
    assert sol.original["hist_newton"] == ref.hist_newton
    assert _rel(sol.t, ref.t) < 1e-10 and _rel(sol.u, ref.u) < 1e-10
    ts = np.linspace(tspan[0], tspan[1], 29)
    for deriv in (0, 1):
        want = np.stack([ref(t, deriv) for t in ts])
        assert _rel(sol(ts, deriv=deriv), want) < 1e-9


def test_mirk6i_convergence_order_on_gpu(M):
    def exact(t):
        return 5.0 * (np.cos(t) - np.sin(t) / np.tan(5.0))

    errs = []
    for dt in (0.5, 0.25, 0.125):
        sol = M.solve(M.BVProblem("linear2", [5.0, -3.5], (0.0, 5.0), p=LIN_P), M.MIRK6I(), dt=dt, adaptive=False,
                      nlsolve_kwargs={"abstol": 1e-13})
        errs.append(np.max(np.abs(sol.u[:, 0] - exact(sol.t))))
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert abs(np.mean(rates) - 6) < 0.5


def test_previous_solution_as_initial_guess_carries_its_mesh(M, oracle):
    """CORE/utils.jl:701-704: a DiffEqArray / ODESolution guess brings its own mesh.  Restarting
    from a converged adaptive solution on its final mesh needs no refinement and at most one Newton step, and
    the guess object is left untouched."""
    O = oracle
    p = [9.81]
    u0 = [math.pi / 2, math.pi / 2]
    first = M.solve(M.BVProblem("pendulum", u0, PENDULUM_T, p=p), M.MIRK4(), dt=0.05)
    assert first.retcode == 0
    keep_t, keep_u = first.t.copy(), first.u.copy()
    again = M.solve(M.BVProblem("pendulum", first, PENDULUM_T, p=p), M.MIRK4())
    ref = O.solve(O.builtin("pendulum"), 4, p, keep_t, keep_u)
    assert again.retcode == ref.retcode == 0
    assert again.original["hist_n_mesh"] == ref.hist_N and again.original["hist_newton"] == ref.hist_newton
    assert len(again.original["hist_n_mesh"]) == 1 and again.original["hist_newton"][0] <= 1
    assert np.array_equal(again.t, keep_t) and _rel(again.u, ref.u) < 1e-10
    assert np.array_equal(first.t, keep_t) and np.array_equal(first.u, keep_u)

    class DiffEqArrayLike:   # the reference's other mesh-carrying guess type
        def __init__(self, t, u):
            self.t, self.u = list(t), [list(v) for v in u]
    third = M.solve(M.BVProblem("pendulum", DiffEqArrayLike(keep_t, keep_u), PENDULUM_T, p=p), M.MIRK4())
    assert third.retcode == 0 and np.array_equal(third.t, keep_t)
