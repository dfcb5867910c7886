"""The numpy restatement of the reference NLPs: sizes, orderings, derivatives."""
import numpy as np
import pytest

from oracle.nlp import NLP, Params, default_scenario, predict_obstacles


def test_horizon_and_sizes():
    p = Params()
    assert p.N_p == 50  # arange(0, 5.1, 0.1) has 51 points (PKG/MPC_CBF_optimize_kin.py:32-33)
    sizes = {"kin_nocbf": (304, 204), "kin_cbf": (304, 303), "kin_cbf_pre": (304, 303), "dyn": (406, 455)}
    for kind, (nv, rows) in sizes.items():
        nlp = default_scenario(kind)
        assert nlp.nv == nv
        assert nlp.n_eq + nlp.n_ineq == rows
        assert len(nlp.g_ref(nlp.rollout_start())) == rows


def test_bounds_values():
    nlp = default_scenario("kin_cbf")
    N = nlp.N
    assert np.allclose(nlp.zL[:2], [-0.6108652381980153, -3.0])
    assert np.allclose(nlp.zU[:2], [0.6108652381980153, 3.0])
    xb = nlp.zL[2 * N: 2 * N + 4], nlp.zU[2 * N: 2 * N + 4]
    assert xb[0][0] == -np.inf and xb[0][1] == -1 and xb[0][3] == 0 and xb[1][1] == 5 and xb[1][3] == 40
    assert np.allclose(nlp.dL[:49], -0.008726646259971648) and np.allclose(nlp.dU[:49], 0.008726646259971648)
    assert np.all(nlp.dL[49:] == 0) and np.all(np.isinf(nlp.dU[49:]))
    # obstacle semi-axes: 2.4 + 2.4 + 1.0, 0.9 + 0.9 + 0.5
    assert np.allclose(nlp.osx, 5.8) and np.allclose(nlp.osy, 2.3)


def test_dyn_row_order_and_aligned_bounds():
    nlp = default_scenario("dyn")
    perm = nlp.g_perm()
    assert sorted(perm) == list(range(455))
    lo, hi = nlp.lbg_ubg_aligned()
    # [init(6), d0(6), d1(6), ddf1, dax1, d2(6), ...]
    assert np.all(lo[:18] == 0) and np.all(hi[:18] == 0)
    assert np.isclose(lo[18], -0.008726646259971648) and np.isclose(lo[19], -0.3) and np.isclose(hi[19], 0.15)
    assert np.all(lo[-51:] == 1) and np.all(np.isinf(hi[-51:]))


def test_objective_matches_definition():
    nlp = default_scenario("kin_cbf", N=5)
    rng = np.random.default_rng(1)
    z = rng.standard_normal(nlp.nv)
    U, X = nlp.split(z)
    Q, R, DR = np.diag(nlp.w.Q), np.diag(nlp.w.R), np.diag(nlp.w.DR)
    f = 0.0
    for i in range(5):
        e = X[i] - nlp.xs
        du = U[i] - (U[i - 1] if i > 0 else 0)
        f += e @ Q @ e + U[i] @ R @ U[i] + du @ DR @ du
    assert np.isclose(nlp.objective(z), f, rtol=1e-13)
    nlp0 = default_scenario("kin_nocbf", N=5)  # no i=0 rate cost
    f0 = 0.0
    Q, R, DR = np.diag(nlp0.w.Q), np.diag(nlp0.w.R), np.diag(nlp0.w.DR)
    for i in range(5):
        e = X[i] - nlp0.xs
        f0 += e @ Q @ e + U[i] @ R @ U[i] + (0 if i == 0 else (U[i] - U[i - 1]) @ DR @ (U[i] - U[i - 1]))
    assert np.isclose(nlp0.objective(z), f0, rtol=1e-13)


@pytest.mark.parametrize("kind,opts", [("kin_nocbf", {}), ("kin_cbf", {}), ("kin_cbf_pre", {}), ("dyn", {}),
                                       ("kin_cbf", {"cbf_gamma": 1.0}), ("kin_cbf_pre", {"cbf_gamma": 0.3, "xref": True})])
def test_derivatives_against_central_differences(kind, opts):
    N = 6
    rng = np.random.default_rng(0)
    opts = dict(opts)
    if opts.get("xref"):  # per-stage cost targets (the `aa` blend of PKG/MPC_CBF_optimize_kin.py:194-199)
        opts["xref"] = np.array([400, 3.5, 0, 30.0]) + rng.standard_normal((N, 4))
    nlp = default_scenario(kind, N=N, **opts)
    z = nlp.rollout_start(rng.uniform(-0.1, 0.1, (N, 2))) + 0.05 * rng.standard_normal(nlp.nv)
    eps = 1e-6

    def fd(fun):
        cols = []
        for i in range(nlp.nv):
            e = np.zeros(nlp.nv)
            e[i] = eps
            cols.append((np.atleast_1d(fun(z + e)) - np.atleast_1d(fun(z - e))) / (2 * eps))
        return np.array(cols).T

    g = nlp.grad(z)
    assert np.max(np.abs(fd(nlp.objective)[0] - g)) <= 1e-6 * np.max(np.abs(g))
    assert np.max(np.abs(fd(nlp.eq) - nlp.jac_eq(z))) <= 1e-7
    if nlp.n_ineq:
        assert np.max(np.abs(fd(nlp.ineq) - nlp.jac_ineq(z))) <= 1e-6
    le, li = rng.standard_normal(nlp.n_eq), rng.standard_normal(nlp.n_ineq)

    def gl(zz):
        return 0.7 * nlp.grad(zz) + nlp.jac_eq(zz).T @ le + (nlp.jac_ineq(zz).T @ li if nlp.n_ineq else 0)

    H = nlp.hess_lag(z, le, li, 0.7)
    assert np.max(np.abs(H - fd(gl))) <= 1e-8 * np.max(np.abs(H))
    assert np.max(np.abs(H - H.T)) == 0


def test_predict_obstacles_is_the_reference_recursion():
    obs = [np.array([[60, 10, np.pi / 4, 10, 3.6, 1.5]])]
    tr = predict_obstacles(obs, 0.1, 50)[0]
    x, y = 60.0, 10.0
    for k in range(51):
        assert tr[k, 0] == x and tr[k, 1] == y
        x = x + 10 * np.cos(np.pi / 4) * 0.1
        y = y + 10 * np.sin(np.pi / 4) * 0.1


def test_rk4_increment_model_derivatives_against_finite_differences():
    """oracle/nlp.py Rk4KinModel: forward-mode Jacobian and second-order adjoint of the Runge-Kutta increment function."""
    import numpy as np

    from oracle import nlp

    m = nlp.Rk4KinModel(nlp.Params(), 0.1)
    rng = np.random.default_rng(7)
    for _ in range(6):
        x = np.array([rng.uniform(0, 50), rng.uniform(0, 4), rng.uniform(-0.4, 0.4), rng.uniform(2, 30)])
        u = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-3, 3)])
        z = np.concatenate([x, u])
        lam = rng.normal(size=4)
        J, H = m.jac(x, u), m.hess(x, u, lam)
        Jn, Hn, h = np.zeros((4, 6)), np.zeros((6, 6)), 1e-6
        for i in range(6):
            e = np.zeros(6)
            e[i] = h
            Jn[:, i] = (m.f((z + e)[:4], (z + e)[4:]) - m.f((z - e)[:4], (z - e)[4:])) / (2 * h)
            Hn[:, i] = (lam @ m.jac((z + e)[:4], (z + e)[4:]) - lam @ m.jac((z - e)[:4], (z - e)[4:])) / (2 * h)
        assert np.abs(J - Jn).max() <= 1e-7 and np.abs(H - Hn).max() <= 1e-6 and np.abs(H - H.T).max() <= 1e-14
        # A keeps the sparsity of the Euler step, x and y never enter
        assert np.all(J[:, :2] == 0) and J[2, 2] == 0 and np.all(J[3, :5] == 0) and abs(J[3, 5] - 1.0) <= 1e-15
        # fourth-order accurate: against a fine explicit-Euler integration of the same right-hand side
        xe = x.copy()
        for _ in range(4000):
            xe = xe + 2.5e-5 * nlp.KinModel.f(m, xe, u)
        assert np.abs(x + 0.1 * m.f(x, u) - xe).max() <= 2e-4


def test_rk4_c_oracle_equals_dense_specification():
    import numpy as np

    from mpc_motion_planning_b200 import scenarios
    from oracle import c_oracle, ipm_dense, nlp

    x0, xs, obs = scenarios.kin_cbf_static(8)
    cfg = c_oracle.make_cfg("kin_cbf", integrator="rk4")
    n = 0
    for b in (1, 3):
        z, lam, info = c_oracle.solve(cfg, x0[b], xs[b], obs[b])
        P = nlp.NLP("kin_cbf", x0[b], xs[b], obs[b, :, 0, :], integrator="rk4")
        r = ipm_dense.solve(P, P.rollout_start())
        assert r.status == info.status == 0 and r.iters == info.iters
        assert abs(r.f - info.f) <= 1e-10 * abs(r.f) and np.abs(r.z - z).max() <= 1e-8
        n += 1
    assert n == 2
