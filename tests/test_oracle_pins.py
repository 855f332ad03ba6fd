"""Pins for the CPU oracle (tests/ may use oracle/; the product never does).

The reference holds no golden vectors and Julia cannot run here, so the
oracle is pinned against: exact-rational tableau identities, scipy.integrate.solve_bvp's own
collocation residual / global Jacobian (same Lobatto formula as MIRK4), finite differences,
dense LAPACK solves, and the analytic answers the reference's tests use
(/root/reference/lib/BoundaryValueDiffEqMIRK/test/Core/mirk_basic_tests.jl).
"""
from fractions import Fraction as F

import numpy as np
import pytest

PI = np.pi


# ---------------------------------------------------------------- tableaus (mirk_tableaus.jl:62-152)
def _exact_tableau(order):
    if order == 2:   # mirk_tableaus.jl:13-34
        c = [F(1, 2)]; v = [F(1, 2)]; b = [F(1)]; x = [[0]]
        cs = [F(0), F(1)]; vs = [F(0), F(1)]; xs = [[0, 0, 0], [0, 0, 0]]
    elif order == 3:   # :36-60
        c = [F(0), F(2, 3)]; v = [F(0), F(4, 9)]; b = [F(1, 4), F(3, 4)]; x = [[0, 0], [F(2, 9), 0]]
        cs = [F(1)]; vs = [F(1)]; xs = [[0, 0, 0]]
    elif order == 5:   # :89-118
        c = [F(0), F(1), F(3, 4), F(3, 10)]; v = [F(0), F(1), F(27, 32), F(837, 1250)]
        b = [F(5, 54), F(1, 14), F(32, 81), F(250, 567)]
        x = [[0] * 4, [0] * 4, [F(3, 64), F(-9, 64), 0, 0], [F(21, 1000), F(63, 5000), F(-252, 625), 0]]
        cs = [F(4, 5), F(13, 23)]; vs = cs
        xs = [[F(14, 1125), F(-74, 875), F(-128, 3375), F(104, 945), 0, 0],
              [F(1, 2), F(4508233, 1958887), F(48720832, 2518569), F(-27646420, 17629983), F(-11517095, 559682), 0]]
    elif order == 4:
        c = [F(0), F(1), F(1, 2)]; v = [F(0), F(1), F(1, 2)]
        b = [F(1, 6), F(1, 6), F(2, 3)]
        x = [[0, 0, 0], [0, 0, 0], [F(1, 8), F(-1, 8), 0]]
        cs = [F(3, 4)]; vs = [F(27, 32)]; xs = [[F(3, 64), F(-9, 64), 0, 0]]
    else:
        c = [F(0), F(1), F(1, 4), F(3, 4), F(1, 2)]
        v = [F(0), F(1), F(5, 32), F(27, 32), F(1, 2)]
        b = [F(7, 90), F(7, 90), F(16, 45), F(16, 45), F(2, 15)]
        x = [[0] * 5, [0] * 5, [F(9, 64), F(-3, 64), 0, 0, 0], [F(3, 64), F(-9, 64), 0, 0, 0],
             [F(-5, 24), F(5, 24), F(2, 3), F(-2, 3), 0]]
        cs = [F(7, 16), F(3, 8), F(9, 16), F(1, 8)]; vs = cs
        xs = [[F(1547, 32768), F(-1225, 32768), F(749, 4096), F(-287, 2048), F(-861, 16384), 0, 0, 0, 0],
              [F(83, 1536), F(-13, 384), F(283, 1536), F(-167, 1536), F(-49, 512), 0, 0, 0, 0],
              [F(1225, 32768), F(-1547, 32768), F(287, 2048), F(-749, 4096), F(861, 16384), 0, 0, 0, 0],
              [F(233, 3456), F(-19, 1152), 0, 0, 0, F(-5, 72), F(7, 72), F(-17, 216), 0]]
    return c, v, b, x, cs, vs, xs


@pytest.mark.parametrize("order", [2, 3, 4, 5, 6])
def test_tableau_matches_exact_rationals_and_identities(oracle, order):
    T = oracle.tableau(order)
    c, v, b, x, cs, vs, xs = _exact_tableau(order)
    s = len(c)
    assert T.s == s and T.s_star == s + len(cs)
    for r in range(s):
        assert T.c[r] == float(c[r]) and T.v[r] == float(v[r]) and T.b[r] == float(b[r])
        for j in range(s):
            assert T.x[r][j] == float(F(x[r][j]))
        # stage consistency c_r = v_r + sum_j x_rj
        assert c[r] == v[r] + sum(F(q) for q in x[r])
    for r in range(len(cs)):
        assert T.c_star[r] == float(cs[r]) and T.v_star[r] == float(vs[r])
        for j in range(T.s_star):
            assert T.x_star[r][j] == float(F(xs[r][j]))
        assert cs[r] == vs[r] + sum(F(q) for q in xs[r])
    # quadrature order conditions sum b c^k = 1/(k+1) up to the method order
    for k in range(order):
        assert sum(b[r] * c[r] ** k for r in range(s)) == F(1, k + 1)


@pytest.mark.parametrize("order", [4, 6])
def test_interp_weights_endpoint_identities(oracle, order):
    T = oracle.tableau(order)
    w0, wp0 = oracle.interp_weights(order, 0.0)
    w1, wp1 = oracle.interp_weights(order, 1.0)
    assert np.allclose(w0, 0, atol=1e-15)
    bfull = np.zeros(T.s_star); bfull[:T.s] = [T.b[r] for r in range(T.s)]
    assert np.allclose(w1, bfull, atol=5e-13)          # u(t_{i+1}) = y_i + h sum b_r K_r
    e1 = np.zeros(T.s_star); e1[0] = 1
    assert np.allclose(wp0, e1, atol=1e-14)            # u'(t_i) = K_1
    e2 = np.zeros(T.s_star); e2[1] = 1
    assert np.allclose(wp1, e2, atol=2e-11)            # u'(t_{i+1}) = K_2
    # w' is the derivative of w
    for tau in (0.226, 0.5, 0.7156):
        h = 1e-6
        wa, _ = oracle.interp_weights(order, tau - h)
        wb, _ = oracle.interp_weights(order, tau + h)
        _, wp = oracle.interp_weights(order, tau)
        assert np.allclose((wb - wa) / (2 * h), wp, atol=1e-7)


@pytest.mark.parametrize("order", [2, 3, 5])
def test_interp_weights_low_orders(oracle, order):
    """MIRK2/3/5 interpolants (interpolation.jl:463-527): w(0) = 0, w' is the derivative of w, the weights
    reproduce polynomials up to the interpolant's degree (sum_r w_r(tau) c~_r^k = tau^(k+1)/(k+1) with the
    abscissae of all s* stages), and u'(t_i) = f(y_i)."""
    T = oracle.tableau(order)
    w0, wp0 = oracle.interp_weights(order, 0.0)
    assert np.allclose(w0, 0, atol=1e-15)
    for tau in (0.25, 0.3, 0.5, 0.75):
        h = 1e-6
        wa, _ = oracle.interp_weights(order, tau - h)
        wb, _ = oracle.interp_weights(order, tau + h)
        w, wp = oracle.interp_weights(order, tau)
        assert np.allclose((wb - wa) / (2 * h), wp, atol=1e-7)
        call = [T.c[r] for r in range(T.s)] + [T.c_star[r] for r in range(T.s_star - T.s)]
        for k in range(2 if order < 5 else 4):   # quadrature conditions of the continuous extension
            assert abs(sum(w[r] * call[r] ** k for r in range(T.s_star)) - tau ** (k + 1) / (k + 1)) < 1e-12
    # the stage evaluated at t_i carries u'(t_i): K~_1 of MIRK2 (c* = 0), K_1 otherwise
    first = 1 if order == 2 else 0
    e = np.zeros(T.s_star); e[first] = 1
    assert np.allclose(wp0, e, atol=1e-14)


def test_interval_matches_searchsortedfirst(oracle):
    mesh = np.array([0.0, 0.5, 1.0, 2.0])
    # clamp(searchsortedfirst(mesh,t)-1, 1, N-1), 1-based -> 0-based here
    for t, want in [(-1.0, 0), (0.0, 0), (0.25, 0), (0.5, 0), (0.75, 1), (1.0, 1), (1.5, 2), (2.0, 2), (3.0, 2)]:
        assert oracle.interval(mesh, t) == want


def test_mesh_uniform_is_correctly_rounded(oracle):
    for (t0, t1, n) in [(0.0, PI / 2, 32), (0.0, 5.0, 25), (-1.0, 1.0, 200), (0.0, 0.5, 19999)]:
        m = oracle.mesh_uniform(t0, t1, n)
        want = [float(F(t0) + (F(t1) - F(t0)) * i / n) for i in range(n + 1)]
        assert m[0] == t0 and m[-1] == t1
        assert np.array_equal(m, np.array(want))


# ---------------------------------------------------------------- MIRK4 vs scipy.solve_bvp internals
def test_mirk4_residual_and_jacobian_match_scipy_collocation(oracle):
    from scipy.integrate import _bvp

    P = oracle.builtin("linear2_tp")   # only the ODE part is compared
    Pp = oracle.builtin("pendulum")
    rng = np.random.default_rng(3)
    mesh = np.sort(np.concatenate([[0.0, 1.3], rng.uniform(0, 1.3, 9)]))
    y = rng.normal(size=(len(mesh), 2))
    ws = oracle.Workspace(Pp, 4, [9.81], mesh, y)
    phi = ws.phi()
    Lb, Rb = ws.jac_blocks()

    def fun(x, yy, p=None):
        return np.vstack([yy[1], -9.81 * np.sin(yy[0])])

    def fun_jac(x, yy, p=None):
        J = np.zeros((2, 2, yy.shape[1]))
        J[0, 1] = 1.0
        J[1, 0] = -9.81 * np.cos(yy[0])
        return J, None

    h = np.diff(mesh)
    col_res, y_middle, f, f_middle = _bvp.collocation_fun(fun, y.T, np.zeros(0), mesh, h)
    # scipy: res = y_{i+1} - y_i - h/6 (f_i + f_{i+1} + 4 f_mid)  == Phi_i   (mirk_tableaus.jl:65-72)
    assert np.allclose(phi, col_res.T, rtol=0, atol=1e-13)
    # stages: K1 = f_i, K2 = f_{i+1}, K3 = f_mid
    assert np.allclose(ws.Kd[:, 0, :], f[:, :-1].T, atol=1e-14)
    assert np.allclose(ws.Kd[:, 1, :], f[:, 1:].T, atol=1e-14)
    assert np.allclose(ws.Kd[:, 2, :], f_middle.T, atol=1e-13)
    # blocks of the global Jacobian
    x_middle = mesh[:-1] + 0.5 * h
    df_dy, _ = fun_jac(mesh, y.T)
    df_dy_middle, _ = fun_jac(x_middle, y_middle)
    n, m = 2, len(mesh)
    i_jac, j_jac = _bvp.compute_jac_indices(n, m, 0)
    Jg = _bvp.construct_global_jac(n, m, 0, i_jac, j_jac, h, df_dy, df_dy_middle, None, None,
                                   np.zeros((2, 2)), np.zeros((2, 2)), None).toarray()
    for i in range(m - 1):
        assert np.allclose(Jg[i * n:(i + 1) * n, i * n:(i + 1) * n], Lb[i], atol=1e-12)
        assert np.allclose(Jg[i * n:(i + 1) * n, (i + 1) * n:(i + 2) * n], Rb[i], atol=1e-12)


# ---------------------------------------------------------------- analytic Jacobians vs finite differences
@pytest.mark.parametrize("name,order,p", [
    ("pendulum", 4, [9.81]), ("pendulum", 6, [9.81]),
    ("swirling", 4, [0.01]), ("swirling", 6, [0.3]),
    ("torus", 6, [3.0, 2.0, 0.5, -1.2, -0.5, 0.3]), ("lotka", 6, [7.5, 4.0, 8.0, 5.0]),
    ("layer", 4, [0.05]), ("chain8", 6, None), ("linear2", 6, [2.0, 0.3, 1.0, 0.9, 0.5, 0, 1]),
])
def test_block_jacobian_matches_central_differences(oracle, name, order, p):
    P = oracle.builtin(name)
    rng = np.random.default_rng(11)
    if p is None:
        p = np.concatenate([[9.81, 4.0], rng.uniform(-1, 1, 16)])
    N = 6
    mesh = np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0, 1, N - 2)]))
    y = rng.normal(size=(N, P.n)) * 0.5
    ws = oracle.Workspace(P, order, p, mesh, y)
    J = ws.dense_jacobian()
    Lrows = P.n_bca if P.problem_type == 1 else P.n_bc
    # the collocation rows are exact derivatives; check them by central differences
    f0 = ws.loss()
    Jfd = np.zeros_like(J)
    eps = 1e-6
    for k in range(N * P.n):
        yp = y.copy().reshape(-1); ym = yp.copy()
        yp[k] += eps; ym[k] -= eps
        fp = oracle.Workspace(P, order, p, mesh, yp).loss()
        fm = oracle.Workspace(P, order, p, mesh, ym).loss()
        Jfd[:, k] = (fp - fm) / (2 * eps)
    rows = slice(Lrows, Lrows + (N - 1) * P.n)
    scale = 1 + np.abs(Jfd[rows]).max()
    assert np.abs(J[rows] - Jfd[rows]).max() / scale < 2e-8
    # BC rows: exact when every BC point is an end point (no quirk Q2 in play)
    if name in ("swirling", "torus", "lotka", "layer", "chain8"):
        bc_rows = np.r_[0:Lrows, Lrows + (N - 1) * P.n:J.shape[0]]
        assert np.abs(J[bc_rows] - Jfd[bc_rows]).max() < 2e-8
    assert f0.shape[0] == J.shape[0]


def test_bc_jacobian_reference_pattern_Q2(oracle):
    """Interior BC time: derivative lands on the LEFT node of the containing interval only."""
    P = oracle.builtin("pendulum")
    mesh = oracle.mesh_uniform(0.0, PI / 2, 32)
    ws = oracle.Workspace(P, 4, [9.81], mesh, np.tile([PI / 2, PI / 2], (33, 1)))
    ws.loss()
    nodes, B = ws.bc_jac()
    # u(pi/4) is node 17 (1-based) -> searchsortedfirst-1 = interval 16 -> left node index 15 (0-based)
    assert list(nodes) == [15, 32]
    assert np.array_equal(B[0], [[1, 0], [0, 0]]) and np.array_equal(B[1], [[0, 0], [1, 0]])


# ---------------------------------------------------------------- ABD solve vs dense LAPACK
@pytest.mark.parametrize("n,N,interior", [(2, 9, True), (3, 17, False), (6, 12, True), (16, 40, False)])
def test_abd_solve_matches_dense_solve(oracle, n, N, interior):
    rng = np.random.default_rng(5)
    h = 0.1
    Lb = -np.eye(n)[None] + h * rng.normal(size=(N - 1, n, n))
    Rb = np.eye(n)[None] + h * rng.normal(size=(N - 1, n, n))
    nodes = [0, N - 1] + ([N // 2, N // 2] if interior else [])
    B = rng.normal(size=(len(nodes), n, n))
    rb, rp = rng.normal(size=n), rng.normal(size=(N - 1) * n)
    st, delta = oracle.abd_solve(Lb, Rb, nodes, B, rb, rp)
    assert st == 0
    J = np.zeros((N * n, N * n)); rhs = np.concatenate([rb, rp])
    for k, nd in enumerate(nodes):
        J[:n, nd * n:(nd + 1) * n] += B[k]
    for i in range(N - 1):
        J[n + i * n:n + (i + 1) * n, i * n:(i + 1) * n] = Lb[i]
        J[n + i * n:n + (i + 1) * n, (i + 1) * n:(i + 2) * n] = Rb[i]
    want = np.linalg.solve(J, rhs).reshape(N, n)
    assert np.abs(delta - want).max() / np.abs(want).max() < 1e-10 * max(1.0, np.linalg.cond(J) * 1e-4)


def test_abd_solve_is_stable_on_dichotomic_problem(oracle):
    """Growing and decaying modes (forward propagation would lose everything)."""
    n, N = 2, 400
    h = 0.05
    lam = 40.0
    A = np.array([[lam, 0.0], [0.0, -lam]])
    Q = np.array([[1.0, 1.0], [1.0, -1.0]]) / np.sqrt(2)
    A = Q @ A @ Q.T
    Lb = np.tile(-(np.eye(2) + h / 2 * A), (N - 1, 1, 1))
    Rb = np.tile(np.eye(2) - h / 2 * A, (N - 1, 1, 1))
    B = np.array([[[1.0, 0.0], [0.0, 0.0]], [[0.0, 0.0], [1.0, 0.0]]])
    x_true = np.random.default_rng(1).normal(size=(N, 2))
    rp = np.concatenate([Lb[i] @ x_true[i] + Rb[i] @ x_true[i + 1] for i in range(N - 1)])
    rb = B[0] @ x_true[0] + B[1] @ x_true[-1]
    st, delta = oracle.abd_solve(Lb, Rb, [0, N - 1], B, rb, rp)
    assert st == 0
    assert np.abs(delta - x_true).max() < 1e-9


# ---------------------------------------------------------------- end-to-end answers the reference tests use
def _exact_lin(t):  # mirk_basic_tests.jl:54-66
    return np.array([5 * (np.cos(t) - np.sin(t) / np.tan(5)), 5 * (-np.cos(t) / np.tan(5) - np.sin(t))])


@pytest.mark.parametrize("order", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("name,p", [("linear2", [1.0, 0.0, 5.0, 5.0, 0.0, 0, 0]), ("linear2_tp", [1.0, 5.0, 0.0])])
def test_convergence_order_on_linear_problem(oracle, order, name, p):
    """mirk_basic_tests.jl:122-139: estimated order within 0.4 of p, dts = 1/8, 1/4, 1/2."""
    P = oracle.builtin(name)
    errs = []
    for dt in (0.5, 0.25, 0.125):
        s = oracle.solve_dt(P, order, p, [5.0, -3.5], (0.0, 5.0), dt, adaptive=0, abstol=1e-8)
        assert s.retcode == oracle.SUCCESS
        errs.append(max(np.abs(s.u[i] - _exact_lin(s.t[i])).max() for i in range(s.N)))
    est = np.log2(errs[0] / errs[1]), np.log2(errs[1] / errs[2])
    assert abs(est[1] - order) < 0.4


@pytest.mark.parametrize("order", [4, 6])
@pytest.mark.parametrize("name,p", [("linear2", [0.0, 0.0, 5.0, 5.0, 0.0, 0, 0]), ("linear2_tp", [0.0, 5.0, 0.0])])
def test_affineness(oracle, order, name, p):
    """mirk_basic_tests.jl:103-118."""
    s = oracle.solve_dt(oracle.builtin(name), order, p, [5.0, -3.5], (0.0, 5.0), 0.2)
    assert np.abs(np.diff(s.u[:, 0]) + 0.2).max() + abs(s.u[0, 0] - 5) < 1e-2


@pytest.mark.parametrize("order", [4, 6])
def test_interpolation_two_point_analytic(oracle, order):
    """mirk_basic_tests.jl:183-222: u''=u/lambda, u(0)=1, u(1)=0, sol(0.001) and sol'(0.04) to 1e-6."""
    P = oracle.builtin("linear2")
    s = oracle.solve_dt(P, order, [-1.0, 0.0, 1.0, 1.0, 0.0, 0, 0], [1.0, 0.0], (0.0, 1.0), 0.001)

    def an(t):
        return np.array([(np.exp(-t) - np.exp(t - 2)) / (1 - np.exp(-2)),
                         (-np.exp(-t) - np.exp(t - 2)) / (1 - np.exp(-2))])

    def dan(t):
        return np.array([an(t)[1], an(t)[0]])

    assert s.retcode == oracle.SUCCESS
    assert np.allclose(s(0.001), an(0.001), atol=1e-6)
    assert np.allclose(s(0.04), an(0.04), atol=1e-6)
    assert np.allclose(s(0.04, deriv=1), dan(0.04), atol=1e-6)


@pytest.mark.parametrize("order", [4, 6])
@pytest.mark.parametrize("adaptive", [0, 1])
def test_interpolation_multipoint_analytic(oracle, order, adaptive):
    """mirk_basic_tests.jl:272-312: sol(pi/6)[1]=0.5, sol(pi/3)[2]=0.5 -> (sin t, cos t)."""
    P = oracle.builtin("linear2")
    s = oracle.solve_dt(P, order, [1.0, PI / 6, 0.5, PI / 3, 0.5, 0, 1], [0.0, 1.0], (0.0, PI / 2), 0.001,
                        adaptive=adaptive)
    assert s.retcode == oracle.SUCCESS
    for t in (PI / 6, PI / 3):
        assert np.allclose(s(t), [np.sin(t), np.cos(t)], atol=1e-6)


def test_maxiters_zero_returns_guess_untouched(oracle):
    """mirk_basic_tests.jl:678-720."""
    P = oracle.builtin("linear2_tp")
    mesh = oracle.mesh_uniform(0, 5, 10)
    guess = np.random.default_rng(0).normal(size=(11, 2))
    s = oracle.solve(P, 4, [1.0, 5.0, 0.0], mesh, guess, maxiters=0, adaptive=0)
    assert np.array_equal(s.u, guess)


def test_benchmark_pendulum_C1(oracle):
    """benchmark/simple_pendulum.jl:5-19,32,66-70 (BASELINE config 1)."""
    s = oracle.solve_dt(oracle.builtin("pendulum"), 4, [9.81], [PI / 2, PI / 2], (0.0, PI / 2), 0.05)
    assert s.retcode == oracle.SUCCESS and s.hist_N[0] == 33
    assert abs(s(PI / 4)[0] + PI / 2) < 1e-6 and abs(s.u[-1, 0] - PI / 2) < 1e-6
    assert s.defect_norm <= 1e-6 and s.resid_norm <= 1e-6


def test_swirling_flow_and_torus_and_lotka_and_layer_solve(oracle):
    """mirk_basic_tests.jl:315-344, 438-455, 726-756; test/misc/adaptivity_tests.jl:7-17."""
    s = oracle.solve_dt(oracle.builtin("swirling"), 4, [0.01], np.zeros(6), (0.0, 1.0), 0.01, abstol=1e-4)
    assert s.retcode == oracle.SUCCESS
    p = [3.0, 2.0, 0.5, -1.2, -0.5, 0.3]
    s = oracle.solve_dt(oracle.builtin("torus"), 4, p, [0.5, -1.2, 0, 0], (0.0, 1.0), 0.05)
    assert s.retcode == oracle.SUCCESS and s.N > 21
    assert np.allclose(s.u[0, :2], [0.5, -1.2], atol=1e-8) and np.allclose(s.u[-1, :2], [-0.5, 0.3], atol=1e-8)
    for order in (4, 6):
        s = oracle.solve_dt(oracle.builtin("lotka"), order, [7.5, 4.0, 8.0, 5.0], [1.0, 2.0], (0.0, 10.0), 0.01)
        assert s.retcode == oracle.SUCCESS
    s = oracle.solve_dt(oracle.builtin("layer"), 4, [0.001], [1.0, 0.0], (-1.0, 1.0), 0.01)
    assert s.retcode == oracle.SUCCESS
    # exact solution of the layer problem: cos(pi t) + erf(t/sqrt(2 eps))/erf(1/sqrt(2 eps))
    from scipy.special import erf
    ex = np.cos(PI * s.t) + erf(s.t / np.sqrt(0.002)) / erf(1 / np.sqrt(0.002))
    assert np.abs(s.u[:, 0] - ex).max() < 1e-4


def test_ensemble_of_ten_converges(oracle):
    """ensemble_tests.jl:7-40: p=[rand()], bc u(0)=1, u(1.0)=0 on tspan (0, pi/2), dt=0.1."""
    rng = np.random.default_rng(0)
    params = np.array([[k, 0.0, 1.0, 1.0, 0.0, 0, 0] for k in rng.uniform(size=10)])
    for order in (4, 6):
        ret, Nf, y0, its = oracle.ensemble_solve(oracle.builtin("linear2"), order, params, [0.0, 1.0],
                                                 (0.0, PI / 2), 16, nthreads=2)
        assert (ret == 0).all() and np.allclose(y0[:, 0], 1.0, atol=1e-6)


def test_mesh_selector_properties(oracle):
    """adaptivity.jl:23-75,250-304: halving when all s_hat equal; clamp to [N/2, 4(N-1)]; ends kept."""
    mesh = np.linspace(0, 1, 11)
    err = np.full((10, 2), 1e-4)
    info, m2 = oracle.mesh_select(4, mesh, err)
    assert info == 0 and len(m2) == 21 and np.allclose(m2[::2], mesh) and np.allclose(m2[1::2], mesh[:-1] + 0.05)
    err = np.full((10, 2), 1e-9); err[3] = 1e-3      # one bad interval -> redistribute toward it
    info, m2 = oracle.mesh_select(4, mesh, err)
    assert info == 0 and m2[0] == 0 and m2[-1] == 1 and np.all(np.diff(m2) > 0)
    inside = ((m2 > mesh[3]) & (m2 < mesh[4])).sum()
    assert inside >= len(m2) // 3
    err = np.full((10, 2), 10.0)                      # huge defect -> upper clamp 4*(N-1)
    err[0] *= 1.0001
    info, m2 = oracle.mesh_select(4, mesh, err)
    assert info == 0 and len(m2) == 41
    info, m2 = oracle.mesh_select(4, mesh, err, max_num_subintervals=30)
    assert info == 1


def test_reinterp_quirk_Q3_changes_only_the_newton_path(oracle):
    """mirk.jl:368-370 + adaptivity.jl:602: the reference rewrites y0 in place while reading it.
    That only perturbs the next initial guess: mesh sequence and converged values are unchanged."""
    P = oracle.builtin("pendulum")
    a = oracle.solve_dt(P, 4, [9.81], [PI / 2, PI / 2], (0.0, PI / 2), 0.05, reinterp_inplace=0)
    b = oracle.solve_dt(P, 4, [9.81], [PI / 2, PI / 2], (0.0, PI / 2), 0.05, reinterp_inplace=1)
    assert a.retcode == b.retcode == oracle.SUCCESS and a.hist_N == b.hist_N
    assert np.array_equal(a.t, b.t) and np.abs(a.u - b.u).max() < 5e-6  # both stop at |F|<=1e-6 on a linearly convergent path (Q2)
    assert sum(b.hist_newton) >= sum(a.hist_newton)


def _mirk6i_mp():
    """MIRK6I constants (mirk_tableaus.jl:154-194) in 60-digit arithmetic: the tableau is irrational, so the pin is
    'correctly rounded from a high-precision evaluation of the reference's expressions' instead of exact rationals."""
    import mpmath as mp
    mp.mp.dps = 60
    s21, s7, s3 = mp.sqrt(21), mp.sqrt(7), mp.sqrt(3)
    q = mp.mpf
    c = [q(0), q(1), q(1) / 2 - s21 / 14, q(1) / 2 + s21 / 14, q(1) / 2]
    v = [q(0), q(1), q(1) / 2 - 9 * s21 / 98, q(1) / 2 + 9 * s21 / 98, q(1) / 2]
    b = [q(1) / 20, q(1) / 20, q(49) / 180, q(49) / 180, q(16) / 45]
    x = [[q(0)] * 5 for _ in range(5)]
    x[2][0], x[2][1] = q(1) / 14 + s21 / 98, -q(1) / 14 + s21 / 98
    x[3][0], x[3][1] = q(1) / 14 - s21 / 98, -q(1) / 14 - s21 / 98
    x[4][0], x[4][1], x[4][2], x[4][3] = -q(5) / 128, q(5) / 128, 7 * s21 / 128, -7 * s21 / 128
    cs = [q(1) / 2, q(1) / 2 - s7 / 14, q(87) / 100]
    xs = [[q(0)] * 8 for _ in range(3)]
    xs[0][:4] = [q(1) / 64, -q(1) / 64, q(7) / 192 * s21, -q(7) / 192 * s21]
    xs[1][:6] = [q(3) / 112 + q(9) / 1960 * s7, -q(3) / 112 + q(9) / 1960 * s7,
                 q(11) / 840 * s7 + q(3) / 112 * s7 * s3, q(11) / 840 * s7 - q(3) / 112 * s7 * s3,
                 q(88) / 5145 * s7, -q(18) / 343 * s7]
    T12 = q(10) ** 12
    xs[2][:7] = [q(2707592511) / T12 - q(1006699707) / T12 * s7, -q(51527976591) / T12 - q(1006699707) / T12 * s7,
                 -q(610366393) / 75000000000 + q(7046897949) / T12 * s7 + q(14508670449) / T12 * s7 * s3,
                 -q(610366393) / 75000000000 + q(7046897949) / T12 * s7 - q(14508670449) / T12 * s7 * s3,
                 -q(12456457) / 1171875000 + q(1006699707) / 109375000000 * s7,
                 q(3020099121) / 437500000000 * s7 + q(47328957) / 625000000, -q(7046897949) / 250000000000 * s7]
    return c, v, b, x, cs, xs


def test_mirk6i_tableau_and_interpolant(oracle):
    """MIRK6I (`order` code 7).  Constants within 1 ulp of the reference's expressions, stage consistency
    c = v + sum x (also for the interpolation stages), quadrature conditions up to order 6, and for the continuous
    extension (interpolation.jl:582-710): w(0) = 0, w(1) = [b; 0], u'(t_i) = K_1, u'(t_{i+1}) = K_2, w' = dw/dtau and
    sum_r w_r(tau) c~_r^k = tau^(k+1)/(k+1) for k < 6 over all s* = 8 abscissae."""
    import mpmath as mp
    O = oracle
    T = O.tableau(O.MIRK6I)
    c, v, b, x, cs, xs = _mirk6i_mp()
    assert (T.s, T.s_star) == (5, 8) and T.tau_star == 0.4

    def close(a, exact):
        # a few ulps of the largest term: some entries are differences of two numbers 60x their size, and the
        # reference evaluates them in Float64 too
        return abs(a - float(exact)) <= 4 * np.spacing(max(abs(float(exact)), 0.1))

    for r in range(5):
        assert close(T.c[r], c[r]) and close(T.v[r], v[r]) and close(T.b[r], b[r])
        for j in range(5):
            assert close(T.x[r][j], x[r][j])
        assert abs(c[r] - v[r] - sum(x[r])) < mp.mpf(10) ** -50
    for r in range(3):
        assert close(T.c_star[r], cs[r]) and close(T.v_star[r], cs[r])
        for j in range(8):
            assert close(T.x_star[r][j], xs[r][j])
        assert abs(sum(xs[r])) < mp.mpf(10) ** -11   # c* = v*: the row sums vanish (to the reference's 12 printed digits)
    for k in range(6):
        assert abs(sum(b[r] * c[r] ** k for r in range(5)) - mp.mpf(1) / (k + 1)) < mp.mpf(10) ** -50
    w0, wp0 = O.interp_weights(O.MIRK6I, 0.0)
    w1, wp1 = O.interp_weights(O.MIRK6I, 1.0)
    bfull = np.zeros(8); bfull[:5] = [float(q) for q in b]
    assert np.allclose(w0, 0, atol=1e-15) and np.allclose(w1, bfull, atol=1e-13)
    e1 = np.zeros(8); e1[0] = 1
    e2 = np.zeros(8); e2[1] = 1
    assert np.allclose(wp0, e1, atol=1e-13) and np.allclose(wp1, e2, atol=1e-12)
    call = np.array([float(q) for q in c] + [float(q) for q in cs])
    for tau in (0.1, 0.4, 0.6, 0.93):
        h = 1e-6
        wa, _ = O.interp_weights(O.MIRK6I, tau - h)
        wb, _ = O.interp_weights(O.MIRK6I, tau + h)
        w, wp = O.interp_weights(O.MIRK6I, tau)
        assert np.allclose((wb - wa) / (2 * h), wp, atol=1e-7)
        for k in range(6):
            assert abs(w @ call ** k - tau ** (k + 1) / (k + 1)) < 1e-13


@pytest.mark.parametrize("name,p", [("linear2", [1.0, 0.0, 5.0, 5.0, 0.0, 0, 0]), ("linear2_tp", [1.0, 5.0, 0.0])])
def test_mirk6i_convergence_order_and_interpolation(oracle, name, p):
    """mirk_basic_tests.jl:122-139 for MIRK6I, plus the dense output against the analytic solution."""
    O = oracle
    P = O.builtin(name)
    errs = []
    for dt in (0.5, 0.25, 0.125):
        s = O.solve_dt(P, O.MIRK6I, p, [5.0, -3.5], (0.0, 5.0), dt, adaptive=0, abstol=1e-8)
        assert s.retcode == O.SUCCESS
        errs.append(max(np.abs(s.u[i] - _exact_lin(s.t[i])).max() for i in range(s.N)))
    assert abs(np.log2(errs[1] / errs[2]) - 6) < 0.4
    s = O.solve_dt(P, O.MIRK6I, p, [5.0, -3.5], (0.0, 5.0), 0.1)
    for t in (0.37, 2.51, 4.99):
        assert np.abs(s(t) - _exact_lin(t)).max() < 1e-6


def test_reference_pattern_jacobian_mode_quirk_Q1(oracle):
    """the Jacobian the reference's default sparse path assembles (pattern too narrow for n > 2, out-of-band
    entries aliased onto the in-band column of their colour).  n = 2: identical to the exact Jacobian.  n = 16: differs
    only in O(h) entries, exactly in the columns the aliasing rule predicts, and the Newton count on C2's problem is the
    same as with the exact Jacobian."""
    import math
    O = oracle
    mesh = O.mesh_uniform(0.0, math.pi / 2, 32)
    ws = O.Workspace(O.builtin("pendulum"), 4, [9.81], mesh, np.tile([math.pi / 2, math.pi / 2], (33, 1)))
    assert np.array_equal(ws.dense_jacobian(), ws.dense_jacobian_reference_pattern())
    # chain8 (two-point, n = 16): band (n + 1, n + 1), period 2n + 3 = 35
    n, nint = 16, 12
    rng = np.random.default_rng(3)
    p = np.concatenate([[9.81, 4.0], rng.uniform(-1, 1, 16)])
    mesh = O.mesh_uniform(0.0, 0.5, nint)
    y = 0.3 * rng.standard_normal((nint + 1, n))
    ws = O.Workspace(O.builtin("chain8"), 6, p, mesh, y)
    Jt, Jr = ws.dense_jacobian(), ws.dense_jacobian_reference_pattern()
    R, Cc = np.nonzero(Jt != Jr)
    assert len(R) > 0
    for r, c in zip(R, Cc):
        inside = -(n + 1) <= c - r <= n + 1
        if inside:   # an aliased value landed here: it is a true entry 35 columns away
            src = [c2 for c2 in (c - 35, c + 35) if 0 <= c2 < Jt.shape[1] and Jt[r, c2] != 0.0]
            assert src and Jr[r, c] == pytest.approx(Jt[r, c] + sum(Jt[r, c2] for c2 in src))
        else:        # a true entry outside the band: dropped from its own place
            assert Jr[r, c] == 0.0 and Jt[r, c] != 0.0
    a = O.Workspace(O.builtin("chain8"), 6, p, mesh, np.zeros((nint + 1, n)))
    b = O.Workspace(O.builtin("chain8"), 6, p, mesh, np.zeros((nint + 1, n)))
    ra, rb = a.newton_reference_pattern(exact=True, maxiters=50), b.newton_reference_pattern(maxiters=50)
    assert ra[0] == rb[0] == 0 and ra[1] == rb[1]


def test_polyalgorithm_order_and_stalled(oracle):
    """The default nonlinear solver tries NewtonRaphson, NewtonRaphson + BackTracking, TrustRegion in that order from the
    same start; the first success wins and the step count is the sum over the solvers that ran."""
    O = oracle
    mesh = O.mesh_uniform(0.0, 10.0, 100)
    y = np.tile([5.0, 5.0], (101, 1))
    runs = {}
    for alg in (0, 1, 2, 3):
        ws = O.Workspace(O.builtin("lotka"), 4, [7.5, 4.0, 8.5, 5.0], mesh, y)
        runs[alg] = ws.nlsolve(alg, maxiters=200)
    assert runs[1][0] != 0 and runs[2][0] == O.STALLED and runs[3][0] == 0
    assert runs[0][0] == 0 and runs[0][1] == runs[1][1] + runs[2][1] + runs[3][1]


def test_singular_term_lane_emden(oracle):
    """prob.singular_term (CORE/src/utils.jl:932-941): y' = S y / t + f(t, y) with S = [0 0; 0 -2], f = [y2, -y1] is the
    Lane-Emden equation of index 1, exact solution sin(t)/t (MIRK/test/Core/singular_bvp_tests.jl:15-21).  The term enters
    the discrete stages with t > 0 only — so the stage at t = 0 of the first interval misses the (finite) limit of S y / t,
    and the fixed-mesh solution converges to the exact one at SECOND order whatever the tableau (a property of the
    reference's rule, reproduced).  The analytic Jacobian (df/du + S/t) agrees with central differences."""
    O = oracle
    P = O.builtin("lane_emden")
    errs = []
    for nint in (50, 100):
        sol = O.solve_dt(P, 4, [], [1.0, 0.0], (0.0, 1.0), 1.0 / nint, adaptive=0)
        assert sol.retcode == 0
        t = sol.t[1:]
        errs.append(np.max(np.abs(sol.u[1:, 0] - np.sin(t) / t)))
    assert errs[1] < 1e-5 and 3.5 < errs[0] / errs[1] < 4.5
    mesh = O.mesh_uniform(0.0, 1.0, 10)
    rng = np.random.default_rng(0)
    y = rng.standard_normal((11, 2))
    ws = O.Workspace(P, 4, [], mesh, y)
    J = ws.dense_jacobian()
    eps = 1e-6
    Jfd = np.zeros_like(J)
    for k in range(y.size):
        yp, ym = y.copy().ravel(), y.copy().ravel()
        yp[k] += eps
        ym[k] -= eps
        Jfd[:, k] = (O.Workspace(P, 4, [], mesh, yp.reshape(y.shape)).loss() - O.Workspace(P, 4, [], mesh, ym.reshape(y.shape)).loss()) / (2 * eps)
    assert np.max(np.abs(J - Jfd)) < 1e-6 * max(1.0, np.max(np.abs(J)))


def test_derivative_boundary_condition_known_answer(oracle):
    """bc! reading sol(t, Val{1}) (MIRK/src/interpolation.jl:277-292): u'' = -u, u(0) = 0,
    u(pi/2) - 1 + alpha (u'(pi/4) - cos(pi/4)) = 0 has the solution (sin t, cos t) for every alpha.  The derivative is
    built from the Float64 stage buffers, so the boundary Jacobian ignores it and Newton converges linearly in that row
    (rate ~ alpha): more steps than the 1 a linear problem needs, and a final error set by abstol, not by the mesh."""
    import math
    O = oracle
    P = O.builtin("robin_sine")
    for order, dt in ((4, 0.05), (6, 0.1)):
        ref = O.solve_dt(P, order, [0.1, math.cos(math.pi / 4)], [0.0, 1.0], (0.0, math.pi / 2), dt)
        assert ref.retcode == 0 and ref.hist_newton[0] > 2
        t, u = np.asarray(ref.t), np.asarray(ref.u)
        assert np.max(np.abs(u[:, 0] - np.sin(t))) < 5e-6 and np.max(np.abs(u[:, 1] - np.cos(t))) < 5e-6
        # with alpha = 0 the condition is the plain Dirichlet one: one Newton step, discretisation-level error
        ref0 = O.solve_dt(P, order, [0.0, 0.0], [0.0, 1.0], (0.0, math.pi / 2), dt)
        assert ref0.retcode == 0 and ref0.hist_newton[0] == 1
        assert np.max(np.abs(np.asarray(ref0.u)[:, 0] - np.sin(np.asarray(ref0.t)))) < 1e-7


def test_interval_threads_do_not_change_the_oracle(oracle):
    """The CPU arm of bench.py runs the oracle's two interval-parallel loops (Phi, Jacobian blocks) on every host thread.
    Intervals are independent: residual, blocks and a Newton iterate must be bit-identical for any thread count."""
    O = oracle
    import mirk_b200  # noqa: F401  (package alias for configs)
    from boundaryvaluediffeq_jl_b200 import configs
    c = configs.c2_chain8(257)
    out = {}
    try:
        for nt in (1, 3, 8):
            O.set_interval_threads(nt)
            ws = O.Workspace(O.builtin(c.problem), c.order, c.p, c.mesh, c.y0)
            r = ws.loss()
            Lb, Rb = ws.jac_blocks()
            ws.newton(abstol=0.0, maxiters=2)
            out[nt] = (r, Lb, Rb, ws.y.copy())
    finally:
        O.set_interval_threads(1)
    for nt in (3, 8):
        for a, b in zip(out[1], out[nt]):
            assert np.array_equal(a, b)
