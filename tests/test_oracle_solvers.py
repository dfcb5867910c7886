"""The C oracle (scalar Riccati) against the dense specification and the golden fixtures."""
import numpy as np
import pytest

from oracle import c_oracle, ipm_dense
from oracle.nlp import NLP, default_scenario


def _obs_for(nlp):
    N = nlp.N
    if nlp.kind == "kin_nocbf":
        return None
    ob = np.zeros((nlp.M, N + 1, 6))
    ob[:, :, 0:2] = nlp.oc
    ob[:, :, 4], ob[:, :, 5] = 4.8, 1.8
    return ob


@pytest.mark.parametrize("kind", ["kin_nocbf", "kin_cbf", "kin_cbf_pre", "dyn"])
def test_newton_step_riccati_equals_dense_solve(kind):
    """One regularised Newton step: stage-wise Riccati vs a dense KKT solve (<= 1e-9)."""
    N = 12
    nlp = default_scenario(kind, N=N)
    rng = np.random.default_rng(3)
    z = nlp.rollout_start(rng.uniform(-0.005, 0.005, (N, 2)))
    mu, dw, sigma = 0.5, 1e-2, 1e-3
    cfg = c_oracle.make_cfg(kind, N=N, init_mode=0)
    rc, dz, lamp = c_oracle.newton_step(cfg, nlp.x0, nlp.xs, _obs_for(nlp), z, mu, dw, sigma)
    assert rc == 0
    # dense replica of the same step: unit bound multipliers, zero constraint multipliers
    zL, zU = ipm_dense._relax(nlp.zL, nlp.zU, 1e-8)
    dL, dU = ipm_dense._relax(nlp.dL, nlp.dU, 1e-8) if nlp.n_ineq else (nlp.dL, nlp.dU)
    zp = ipm_dense._push(z, zL, zU, 1e-2, 1e-2)
    nv, ne, ni = nlp.nv, nlp.n_eq, nlp.n_ineq
    s = ipm_dense._push(nlp.ineq(zp), dL, dU, 1e-2, 1e-2) if ni else np.zeros(0)
    hl, hu = np.isfinite(zL), np.isfinite(zU)
    sl, su = np.where(hl, zp - zL, 1.0), np.where(hu, zU - zp, 1.0)
    Sig = np.where(hl, 1 / sl, 0) + np.where(hu, 1 / su, 0)
    g = sigma * nlp.grad(zp) - np.where(hl, mu / sl, 0) + np.where(hu, mu / su, 0)
    Jc = nlp.jac_eq(zp)
    W = nlp.hess_lag(zp, np.zeros(ne), np.zeros(ni), sigma) + np.diag(Sig + dw)
    rhs_z = -g.copy()
    if ni:
        Jd = nlp.jac_ineq(zp)
        hdl, hdu = np.isfinite(dL), np.isfinite(dU)
        gl, gu = np.where(hdl, s - dL, 1.0), np.where(hdu, dU - s, 1.0)
        D = np.where(hdl, 1 / gl, 0) + np.where(hdu, 1 / gu, 0) + dw
        gs = -np.where(hdl, mu / gl, 0) + np.where(hdu, mu / gu, 0) + 1e-5 * mu * (hdl & ~hdu)
        W = W + (Jd.T * D) @ Jd
        rhs_z -= Jd.T @ (D * (nlp.ineq(zp) - s) + gs)
    K = np.block([[W, Jc.T], [Jc, np.zeros((ne, ne))]])
    sol = np.linalg.solve(K, np.concatenate([rhs_z, -nlp.eq(zp)]))
    assert np.max(np.abs(sol[:nv] - dz)) <= 1e-9 * max(1.0, np.max(np.abs(dz)))
    assert np.max(np.abs(sol[nv:] - lamp)) <= 1e-9 * max(1.0, np.max(np.abs(lamp)))


@pytest.mark.parametrize("kind", ["kin_nocbf", "kin_cbf_pre", "dyn"])
def test_c_oracle_follows_the_dense_specification(kind):
    """Same iterates: identical iteration / regularisation counts, solutions equal to 1e-10."""
    N = 16
    nlp = default_scenario(kind, N=N)
    cfg = c_oracle.make_cfg(kind, N=N)
    z, lam, info = c_oracle.solve(cfg, nlp.x0, nlp.xs, _obs_for(nlp))
    r = ipm_dense.solve(nlp, nlp.rollout_start(), ipm_dense.IpmOptions())
    assert info.status == r.status == 0
    assert info.iters == r.iters and info.n_reg == r.n_reg
    assert np.max(np.abs(z - r.z)) <= 1e-10
    assert abs(info.f - r.f) <= 1e-12 * abs(r.f)
    assert np.max(np.abs(lam - r.lam_eq)) <= 1e-8 * max(1.0, np.max(np.abs(lam)))


@pytest.mark.parametrize("kind,gamma,xref", [("kin_cbf", 0.5, False), ("kin_cbf_pre", 1.0, False), ("kin_cbf_pre", 0.3, True),
                                             ("kin_cbf_pre", None, True)])
def test_c_oracle_follows_the_dense_specification_dcbf_and_stage_reference(kind, gamma, xref):
    """The options of SURVEY.md section 8f row N3: discrete-time CBF rows (coupling X_k and X_{k+1},
    folded into the stage by the dynamics in the Riccati form) and per-stage cost targets."""
    N = 16
    rng = np.random.default_rng(5)
    ref = np.array([400, 3.5, 0, 30.0]) + np.c_[np.zeros(N), 0.3 * rng.standard_normal(N), 0.01 * rng.standard_normal(N), rng.standard_normal(N)]
    nlp = default_scenario(kind, N=N, cbf_gamma=gamma, xref=ref if xref else None)
    cfg = c_oracle.make_cfg(kind, N=N, cbf_gamma=gamma, ref_trajectory=xref)
    z, lam, info = c_oracle.solve(cfg, nlp.x0, ref if xref else nlp.xs, _obs_for(nlp))
    r = ipm_dense.solve(nlp, nlp.rollout_start(), ipm_dense.IpmOptions())
    assert info.status == r.status == 0
    assert info.iters == r.iters and info.n_reg == r.n_reg
    assert np.max(np.abs(z - r.z)) <= 1e-10
    assert abs(info.f - r.f) <= 1e-12 * abs(r.f)
    assert np.max(np.abs(lam - r.lam_eq)) <= 1e-8 * max(1.0, np.max(np.abs(lam)))
    if gamma is not None:  # rows hold at the solution
        d = nlp.ineq(z)
        assert np.all(d >= nlp.dL - 1e-7) and np.all(d <= nlp.dU + 1e-7)


def test_c_oracle_reproduces_golden_fixtures(golden):
    """Fixtures come from the dense solver + SLSQP cross-check (tests/golden/make_golden.py)."""
    n_checked = 0
    for name, c in golden.items():
        kind = str(c["kind"])
        obs = c["obs"] if c["obs"].shape[0] else None
        cfg = c_oracle.make_cfg("kin_cbf_pre" if kind == "kin_cbf" else kind)
        z, lam, info = c_oracle.solve(cfg, c["x0"], c["xs"], obs)
        assert info.status == int(c["status"]), name
        if info.status == 0:
            assert info.iters == int(c["iters"]), name
            assert np.max(np.abs(z[:2] - c["z"][:2])) <= 1e-8, name
            assert abs(info.f - float(c["f"])) <= 1e-10 * abs(float(c["f"])), name
            n_checked += 1
    assert n_checked >= 15


def test_golden_fixtures_are_local_minima_confirmed_by_slsqp(golden):
    conv = [c for c in golden.values() if int(c["status"]) == 0]
    assert sum(int(c["local_min"]) for c in conv) >= 0.9 * len(conv)
    # anchor values recorded in SURVEY.md section 8c (bound_relax shifts them by ~2e-8 relative)
    assert abs(float(golden["kin_cbf_default"]["f"]) - 1.0947508480e8) <= 1e-6 * 1.09e8
    assert np.allclose(golden["kin_cbf_default"]["z"][:2], [0.03564617, 3.0], atol=1e-6)
    assert abs(float(golden["kin_cbf_pre_default"]["f"]) - 1.0859930886e8) <= 1e-6 * 1.09e8
    assert np.allclose(golden["kin_cbf_pre_default"]["z"][:2], [0.03581586, 3.0], atol=1e-6)


def test_infeasible_start_state_is_reported():
    """x0 outside the lane bound: the initial-condition rows cannot be met -> not converged."""
    nlp = default_scenario("kin_cbf")
    cfg = c_oracle.make_cfg("kin_cbf")
    x0 = nlp.x0.copy()
    x0[1] = 6.0  # Y_max = 5
    z, lam, info = c_oracle.solve(cfg, x0, nlp.xs, _obs_for(nlp))
    assert info.status != 0


def test_batch_threads_equal_serial():
    from mpc_motion_planning_b200 import scenarios

    x0, xs, obs = scenarios.kin_cbf_moving(24)
    cfg = c_oracle.make_cfg("kin_cbf_pre")
    a = c_oracle.solve_batch(cfg, x0, xs, obs, nthreads=1)
    b = c_oracle.solve_batch(cfg, x0, xs, obs, nthreads=4)
    for u, v in zip(a[:4], b[:4]):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("N", [8, 50])
def test_c_oracle_dyn_rows_as_shipped_follows_the_dense_specification(N):
    """dyn problem with the bound lists as shipped (oracle.nlp.ShippedDynNLP): controls tied over the horizon,
    relaxed x/y defects as per-stage inputs in the Riccati form."""
    from oracle.nlp import ShippedDynNLP

    nlp = ShippedDynNLP([0, 0, 0, 10, 0, 0.0], [600, 3.5, 0, 15, 0, 0.0], [100, -3.5], N=N)
    obs = np.zeros((1, N + 1, 6))
    obs[0, :, 0], obs[0, :, 1] = 100, -3.5
    z, lam, info = c_oracle.solve(c_oracle.make_cfg("dyn", N=N, rows_as_shipped=True), nlp.x0, nlp.xs, obs)
    r = ipm_dense.solve(nlp, nlp.rollout_start(), ipm_dense.IpmOptions())
    assert info.status == r.status == 0 and info.iters == r.iters and info.n_reg == r.n_reg
    assert np.max(np.abs(z - r.z)) <= 1e-10 and abs(info.f - r.f) <= 1e-12 * abs(r.f)
    U, X = nlp.split(z)
    assert np.abs(U - U[0]).max() == 0.0
    g = nlp.g_ref(z)
    assert np.all(g >= nlp.lbg_shipped - 1e-7) and np.all(g <= nlp.ubg_shipped + 1e-7)


def test_restoration_c_oracle_equals_dense_specification():
    """The restoration phase (elastic l1 problem on the inequality rows, dynamics kept as equalities, p/n eliminated on
    their central path) in the C oracle's Riccati form against the dense specification: same phases entered, same
    iteration counts, same answers on scenarios whose plain solve ends in a line-search failure."""
    from oracle import c_oracle, ipm_dense, nlp
    from mpc_motion_planning_b200 import scenarios

    x0, xs, obs = scenarios.kin_cbf_static(2000)
    cfg_plain = c_oracle.make_cfg("kin_cbf")
    cfg = c_oracle.make_cfg("kin_cbf", restoration=True, resto_max_calls=0)
    for b, n_resto in ((99, 1), (231, 1)):
        _, _, plain = c_oracle.solve(cfg_plain, x0[b], xs[b], obs[b])
        assert plain.status == c_oracle_status_infeasible()
        z, lam, info = c_oracle.solve(cfg, x0[b], xs[b], obs[b])
        P = nlp.NLP("kin_cbf", x0[b], xs[b], obs[b, :, 0, :])
        r = ipm_dense.solve(P, P.rollout_start(), ipm_dense.IpmOptions(restoration=True, resto_max_calls=0))
        assert (r.status, r.iters, r.n_resto, r.n_resto_iter, r.n_reg) == (info.status, info.iters, info.n_resto, info.n_resto_iter, info.n_reg)
        assert info.status == 0 and info.n_resto == n_resto
        assert abs(r.f - info.f) <= 1e-9 * abs(r.f) and np.abs(r.z - z).max() <= 1e-6


def c_oracle_status_infeasible():
    return 3


def test_psi_penalty_derivatives():
    """psi_mu(r): value / first / second derivative consistent, and the implied p, n solve the centred l1 split."""
    from oracle.ipm_dense import _psi

    rho = 1000.0
    for mu in (30.0, 0.1, 1e-5):
        r = np.array([-3.0, -0.2, -1e-4, 0.0, 1e-4, 0.2, 3.0])
        v, d1, d2 = _psi(r, mu, rho)
        h = 1e-4 * np.maximum(mu / rho, np.abs(r))  # well inside the smoothing width mu/rho of the kink at r = 0
        vp, d1p, _ = _psi(r + h, mu, rho)
        vm, d1m, _ = _psi(r - h, mu, rho)
        assert np.allclose((vp - vm) / (2 * h), d1, rtol=2e-4, atol=1e-6 * rho)
        assert np.allclose((d1p - d1m) / (2 * h), d2, rtol=2e-3)
        p_, n_ = mu / (rho - d1), mu / (rho + d1)  # central path: p z_p = mu with z_p = rho - lambda
        assert np.allclose(p_ - n_, r, atol=1e-9 * np.maximum(1, np.abs(r)))


def test_second_order_correction_is_never_accepted_on_these_nlps():
    """The specification carries IPOPT's second-order correction as an option (max_soc); the Riccati implementations do not.
    The reason is measured here: with IPOPT's default max_soc = 4 no corrected step is ever accepted, and the iterates are the
    same as without it - on converging scenarios and on one that ends in a line-search failure."""
    from mpc_motion_planning_b200 import scenarios
    from oracle import ipm_dense, nlp

    x0, xs, obs = scenarios.kin_cbf_static(40)
    for b in (0, 8):  # 8 ends with status 3
        P = nlp.NLP("kin_cbf", x0[b], xs[b], obs[b, :, 0, :])
        r0 = ipm_dense.solve(P, P.rollout_start(), ipm_dense.IpmOptions())
        r4 = ipm_dense.solve(P, P.rollout_start(), ipm_dense.IpmOptions(max_soc=4))
        assert r4.n_soc == 0 and (r0.status, r0.iters) == (r4.status, r4.iters) and np.array_equal(r0.z, r4.z)
    assert r0.status == 3
