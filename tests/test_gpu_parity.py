"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Tolerances are the
ones BASELINE.json states: first control 1e-4 absolute, optimal cost 1e-6 relative, same
converged / not-converged verdict."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

U0_ATOL = 1e-4
COST_RTOL = 1e-6


@pytest.fixture(scope="module")
def dev():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _gpu(solver, dev, x0, xs, obs, z_init=None, **kw):
    import torch

    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = solver.solve(t(x0), t(xs), t(obs) if obs is not None and obs.shape[1] else None, t(z_init), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _check(gpu, ref_u0, ref_cost, ref_st, min_conv, min_same_verdict=None):
    """BASELINE.json's three criteria against the oracle's answers.  Verdict: converged vs not.  Measured agreement on
    large samples is 99.0-99.6 % (tests/tools/resto_check.py, profiles/r02_restoration.txt): the scenarios that differ
    sit on the line-search failure boundary, where one ulp of libm/FMA difference changes the path; they are listed,
    counted and bounded - at most 1.5 % of the batch plus one scenario - not hidden."""
    st = gpu["status"]
    B = st.shape[0]
    both = (st <= 1) & (ref_st <= 1)  # 0 converged, 1 acceptable level: both count as success (IPOPT convention)
    assert both.mean() >= min_conv, both.mean()
    same = (st <= 1) == (ref_st <= 1)
    differ = np.where(~same)[0]
    if differ.size:
        print(f"verdict differs on {differ.size} of {B} scenarios: ids {differ.tolist()[:40]} "
              f"(gpu status {st[differ].tolist()[:40]}, oracle {ref_st[differ].tolist()[:40]})")
    allowed = int(np.ceil(0.015 * B)) + 1 if min_same_verdict is None else int(np.floor((1 - min_same_verdict) * B))
    assert differ.size <= allowed, (differ.size, allowed, differ.tolist()[:40])
    du = np.abs(gpu["u0"] - ref_u0).max(axis=1)
    dc = np.abs(gpu["cost"] - ref_cost) / np.abs(ref_cost)
    ok = (du <= U0_ATOL) & (dc <= COST_RTOL)
    # among commonly converged scenarios a different local minimum is possible in principle
    assert ok[both].mean() >= 0.995, (np.where(both & ~ok)[0], du[both].max(), dc[both].max())
    return both, same


@pytest.mark.parametrize("kind,gen,B", [("kin_nocbf", "kin_nocbf", 128), ("kin_cbf", "kin_cbf_static", 512),
                                         ("kin_cbf_pre", "kin_cbf_moving", 512), ("dyn", "dyn_static", 256)])
def test_batch_parity_with_oracle(dev, kind, gen, B):
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    x0, xs, obs = getattr(scenarios, gen)(B)
    s = BatchSolver(kind)
    g = _gpu(s, dev, x0, xs, obs)
    cfg = c_oracle.make_cfg(kind)
    u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs, obs if obs.shape[1] else None, nthreads=os.cpu_count())
    both, same = _check(g, u0, cost, st, 0.75 if kind not in ("kin_nocbf",) else 1.0)
    # same algorithm, same arithmetic up to libm ulps: iteration counts agree on almost every scenario
    assert (g["iters"][both] == it[both]).mean() >= 0.9
    assert s.launch_info()["launches"] >= 1


def test_golden_fixtures(dev, golden):
    from mpc_motion_planning_b200.solver import BatchSolver

    solvers = {}
    n = 0
    for name, c in golden.items():
        kind = str(c["kind"])
        if int(c["status"]) != 0:
            continue
        k = "kin_cbf_pre" if kind == "kin_cbf" else kind
        s = solvers.setdefault(k, BatchSolver(k))
        obs = c["obs"][None] if c["obs"].shape[0] else None
        g = _gpu(s, dev, c["x0"][None], c["xs"][None], obs, return_z=True)
        assert g["status"][0] == 0, name
        assert np.abs(g["u0"][0] - c["z"][:2]).max() <= U0_ATOL, name
        assert abs(g["cost"][0] - float(c["f"])) <= COST_RTOL * abs(float(c["f"])), name
        assert np.abs(g["z"][0] - c["z"]).max() <= 1e-5, name
        n += 1
    assert n >= 16


def test_dyn_default_scenario_and_shift(dev):
    """PKG/main_cbf_dyn_c_sim.py:44-51 first step; plant step + shift for the 6-state model."""
    import torch

    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    x0 = np.array([[0, 0, 0, 10, 0, 0.0]])
    xs = np.array([[600, 3.5, 0, 15, 0, 0.0]])
    obs = np.zeros((1, 1, 51, 6))
    obs[0, 0, :, 0], obs[0, 0, :, 1] = 100, -3.5
    s = BatchSolver("dyn")
    g = _gpu(s, dev, x0, xs, obs, return_z=True, return_lam=True)
    z, lam, info = c_oracle.solve(c_oracle.make_cfg("dyn"), x0[0], xs[0], obs[0])
    assert g["status"][0] == 0 == info.status
    assert abs(g["cost"][0] - info.f) <= COST_RTOL * info.f and np.abs(g["u0"][0] - z[:2]).max() <= U0_ATOL
    assert np.abs(g["z"][0] - z).max() <= 1e-5
    assert np.abs(g["lam"][0] - lam).max() <= 1e-5 * max(1.0, np.abs(lam).max())
    tx0, tz = torch.from_numpy(x0.copy()).to(dev), torch.from_numpy(g["z"].copy()).to(dev)
    s.shift(tx0, tz)
    torch.cuda.synchronize()
    from mpc_motion_planning_b200 import MPC_CBF_optimize_dyn

    f = MPC_CBF_optimize_dyn.MPC_optimize().f(x0[0], g["z"][0, :2]).full().ravel()
    assert np.allclose(tx0.cpu().numpy()[0], x0[0] + 0.1 * f, rtol=1e-13, atol=1e-13)
    U, X = g["z"][0, :100].reshape(50, 2), g["z"][0, 100:].reshape(51, 6)
    zs = np.concatenate([np.concatenate([U[1:], U[-1:]]).ravel(), np.concatenate([X[1:], X[-1:]]).ravel()])
    assert np.array_equal(tz.cpu().numpy()[0], zs)


def test_reference_default_scenarios_known_answers(dev):
    """SURVEY.md section 8c anchors (first MPC step of the reference mains)."""
    from mpc_motion_planning_b200.Obs_prediction import obs_prediction_batch
    from mpc_motion_planning_b200.solver import BatchSolver

    xs = np.array([[400, 3.5, 0, 30.0]])
    x0 = np.array([[0, 3, 0, 15.0]])
    obs = np.repeat(np.array([50, 3.5, 0, 8, 4.8, 1.8])[None, None, None, :], 51, axis=2)
    g = _gpu(BatchSolver("kin_cbf"), dev, x0, xs, obs)
    assert g["status"][0] == 0 and abs(g["cost"][0] - 1.0947508480e8) <= 1e-6 * 1.1e8
    assert np.allclose(g["u0"][0], [0.03564617, 3.0], atol=1e-6)
    obs = obs_prediction_batch(np.array([50, 3.5, 0, 10, 4.8, 1.8]), 0.1, 50)[None, None]
    g = _gpu(BatchSolver("kin_cbf_pre"), dev, x0, xs, obs)
    assert g["status"][0] == 0 and abs(g["cost"][0] - 1.0859930886e8) <= 1e-6 * 1.1e8
    assert np.allclose(g["u0"][0], [0.03581586, 3.0], atol=1e-6)


def test_size_independent_properties_at_full_batch(dev):
    """BASELINE configs[1] size (B = 10,000): every converged solution must be a feasible KKT
    point of the NLP as the reference states it, checked on the host from z alone."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    B, N = 10000, 50
    x0, xs, obs = scenarios.kin_cbf_static(B)
    s = BatchSolver("kin_cbf")
    g = _gpu(s, dev, x0, xs, obs, return_z=True)
    conv = g["status"] <= 1
    assert conv.mean() >= 0.75
    z = g["z"][conv]
    U = z[:, : 2 * N].reshape(-1, N, 2)
    X = z[:, 2 * N:].reshape(-1, N + 1, 4)
    # initial condition and Euler defects (PKG/MPC_CBF_optimize_kin.py:191,207-208)
    assert np.abs(X[:, 0] - x0[conv]).max() <= 1e-8
    f = np.stack([X[:, :-1, 3] * np.cos(X[:, :-1, 2]), X[:, :-1, 3] * np.sin(X[:, :-1, 2]),
                  X[:, :-1, 3] * np.tan(U[:, :, 0]) / 2.6, U[:, :, 1]], axis=2)
    assert np.abs(X[:, 1:] - (X[:, :-1] + 0.1 * f)).max() <= 1e-7
    # bounds, steering-rate rows, obstacle rows (with IPOPT's 1e-8 bound relaxation)
    eps = 2e-8
    assert U[:, :, 0].min() >= -0.6108652381980153 - eps and U[:, :, 0].max() <= 0.6108652381980153 + eps
    assert np.abs(U[:, :, 1]).max() <= 3 + 4e-8
    assert X[:, :, 1].min() >= -1 - eps and X[:, :, 1].max() <= 5 + 6e-8
    assert X[:, :, 3].min() >= -eps and X[:, :, 3].max() <= 40 + 5e-7
    assert np.abs(np.diff(U[:, :, 0], axis=1)).max() <= 0.008726646259971648 + eps
    o = obs[conv][:, 0, :N]
    h = (X[:, :N, 0] - o[:, :, 0]) ** 2 / 5.8**2 + (X[:, :N, 1] - o[:, :, 1]) ** 2 / 2.3**2 - 1
    assert h.min() >= -1e-7
    # reported cost is the objective of the returned point
    dX = X[:, :N] - xs[conv][:, None, :]
    Q, R, DR = np.array([1e1, 1e5, 3e5, 1e4]), np.array([1e4, 1e4]), np.array([1e5, 1e2])
    dU = np.diff(np.concatenate([np.zeros((len(U), 1, 2)), U], axis=1), axis=1)
    cost = (Q * dX**2).sum((1, 2)) + (R * U**2).sum((1, 2)) + (DR * dU**2).sum((1, 2))
    assert np.abs(cost - g["cost"][conv]).max() <= 1e-9 * cost.max()
    # idempotence: restarting from the solution stays there.  (The restart is re-centred with the initial
    # barrier parameter, which moves the iterate off the solution first; on about 1 % of these non-convex
    # scenarios it then slides into a neighbouring local minimum.)
    g2 = _gpu(s, dev, x0[conv][:512], xs[conv][:512], obs[conv][:512], z_init=z[:512])
    ok = g2["status"] <= 1
    assert ok.mean() >= 0.98
    assert (np.abs(g2["cost"] - g["cost"][conv][:512]) <= COST_RTOL * np.abs(g2["cost"]))[ok].mean() >= 0.97


def test_result_does_not_depend_on_batch_composition_or_entry_point(dev):
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    x0, xs, obs = scenarios.kin_cbf_moving(300)
    s = BatchSolver("kin_cbf_pre")
    full = _gpu(s, dev, x0, xs, obs)
    for i in (0, 17, 299):
        one = _gpu(s, dev, x0[i: i + 1], xs[i: i + 1], obs[i: i + 1])
        for k in ("u0", "cost", "status", "iters"):
            assert np.array_equal(one[k][0], full[k][i]), (i, k)
    host = s.solve(x0, xs, obs)  # numpy in -> host-pointer entry of the C ABI
    for k in ("u0", "cost", "status", "iters"):
        assert np.array_equal(host[k], full[k]), k


def test_edge_cases(dev):
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    s = BatchSolver("kin_cbf_pre")
    # empty batch
    e = s.solve(torch.zeros((0, 4), dtype=torch.float64, device=dev), torch.zeros((0, 4), dtype=torch.float64, device=dev),
                torch.zeros((0, 1, 51, 6), dtype=torch.float64, device=dev))
    assert e["u0"].shape == (0, 2)
    # start state outside the lane (Y_max = 5) or inside the obstacle ellipse: not converged
    x0, xs, obs = scenarios.kin_cbf_moving(4)
    x0[0, 1] = 6.0
    x0[1, 0:2] = obs[1, 0, 0, 0:2]
    g = _gpu(s, dev, x0, xs, obs)
    assert g["status"][0] != 0 and g["status"][1] != 0
    # horizons of the scaling sweep (N = 20, 100) and two obstacles
    for N in (20, 100):
        x0, xs, obs = scenarios.kin_cbf_moving(96, N=N)
        sN = BatchSolver("kin_cbf_pre", N=N)
        g = _gpu(sN, dev, x0, xs, obs)
        u0, cost, st, it, _ = c_oracle.solve_batch(c_oracle.make_cfg("kin_cbf_pre", N=N), x0, xs, obs, nthreads=os.cpu_count())
        _check(g, u0, cost, st, 0.7)
    x0, xs, oa = scenarios.kin_cbf_moving(96)
    _, _, ob = scenarios.kin_cbf_moving(96, seed=99)
    ob[:, :, :, 0] += 60.0
    obs2 = np.concatenate([oa, ob], axis=1)
    g = _gpu(BatchSolver("kin_cbf_pre", M=2), dev, x0, xs, obs2)
    u0, cost, st, it, _ = c_oracle.solve_batch(c_oracle.make_cfg("kin_cbf_pre", M=2), x0, xs, obs2, nthreads=os.cpu_count())
    _check(g, u0, cost, st, 0.6)
    # unsupported shapes fail loudly
    with pytest.raises(ValueError):
        s.solve(torch.zeros((2, 4), dtype=torch.float64, device=dev), torch.zeros((2, 4), dtype=torch.float64, device=dev),
                torch.zeros((2, 1, 50, 6), dtype=torch.float64, device=dev))


def test_as_given_start_and_warm_start_shift(dev):
    """The reference protocol: first solve from its guess, then shifted warm starts
    (PKG/main_cbf_kin_c_sim.py:16-26,92,120)."""
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B, N = 64, 50
    x0, xs, obs = scenarios.kin_cbf_static(B, seed=11)
    s = BatchSolver("kin_cbf", init="as_given")
    cfg = c_oracle.make_cfg("kin_cbf", init_mode=0)
    # as-given guess: a dynamically consistent roll-out with a small constant steer
    U = np.tile(np.array([0.0, 0.5]), (B, N, 1))
    X = np.zeros((B, N + 1, 4))
    X[:, 0] = x0
    for k in range(N):
        X[:, k + 1] = X[:, k] + 0.1 * np.stack([X[:, k, 3] * np.cos(X[:, k, 2]), X[:, k, 3] * np.sin(X[:, k, 2]),
                                                X[:, k, 3] * np.tan(U[:, k, 0]) / 2.6, U[:, k, 1]], axis=1)
    z0 = np.concatenate([U.reshape(B, -1), X.reshape(B, -1)], axis=1)
    g = _gpu(s, dev, x0, xs, obs, z_init=z0, return_z=True)
    u0, cost, st, it, zc = c_oracle.solve_batch(cfg, x0, xs, obs, z_init=z0, want_z=True, nthreads=os.cpu_count())
    _check(g, u0, cost, st, 0.7)
    # device-side plant step + shift equals the host formula
    tx0 = torch.from_numpy(x0).to(dev)
    tz = torch.from_numpy(g["z"]).to(dev)
    s.shift(tx0, tz)
    torch.cuda.synchronize()
    Ug, Xg = g["z"][:, : 2 * N].reshape(B, N, 2), g["z"][:, 2 * N:].reshape(B, N + 1, 4)
    x1 = x0 + 0.1 * np.stack([x0[:, 3] * np.cos(x0[:, 2]), x0[:, 3] * np.sin(x0[:, 2]), x0[:, 3] * np.tan(Ug[:, 0, 0]) / 2.6, Ug[:, 0, 1]], axis=1)
    zs = np.concatenate([np.concatenate([Ug[:, 1:], Ug[:, -1:]], axis=1).reshape(B, -1),
                         np.concatenate([Xg[:, 1:], Xg[:, -1:]], axis=1).reshape(B, -1)], axis=1)
    assert np.allclose(tx0.cpu().numpy(), x1, rtol=1e-14, atol=1e-14)
    assert np.array_equal(tz.cpu().numpy(), zs)
    # warm-started second step agrees with the oracle and needs fewer iterations than the cold one
    conv = g["status"] <= 1
    g2 = _gpu(s, dev, x1[conv], xs[conv], obs[conv], z_init=zs[conv])
    u2, c2, st2, it2, _ = c_oracle.solve_batch(cfg, x1[conv], xs[conv], obs[conv], z_init=zs[conv], nthreads=os.cpu_count())
    _check(g2, u2, c2, st2, 0.9, 0.95)
    assert g2["iters"][g2["status"] <= 1].mean() < g["iters"][conv].mean()


def test_reference_surface_call_protocol(dev, tmp_path, monkeypatch):
    """`optimize_problem(...)` + `solver(x0=,p=,lbx=,...)` exactly as PKG/main_cbf_kin_c_sim.py:87-123."""
    monkeypatch.chdir(tmp_path)
    from mpc_motion_planning_b200 import MPC_CBF_optimize_kin, RefPathGenerator
    from oracle import c_oracle

    mpc = MPC_CBF_optimize_kin.MPC_optimize()
    N_p, ns, nc = mpc.N_p, mpc.num_states, mpc.num_controls
    x0 = np.array([0, 3, 0, 15]).reshape(-1, 1).astype(float)
    xs = np.array([400, 3.5, 0, 30]).reshape(-1, 1).astype(float)
    next_states = np.zeros((N_p + 1, ns))
    u0 = np.zeros((N_p, nc))
    ref = RefPathGenerator.RefPathGenerator()
    ref.define_ref_path(x0, xs, mpc.T_S)
    obs = np.array([[50, 3.5, 0, 8, 4.8, 1.8]])
    lbg, ubg, lbx, ubx = mpc.initialize_constraints(obs)
    # first step: roll the zero controls out so that the guess is dynamically consistent
    for k in range(N_p):
        next_states[k + 1] = next_states[k] if k else x0.ravel()
        next_states[k + 1] = (next_states[k] if k else x0.ravel()) + mpc.T_S * mpc.f(next_states[k] if k else x0.ravel(), u0[k]).full().ravel()
    next_states[0] = x0.ravel()
    last_idx = 0
    costs = []
    cfg = c_oracle.make_cfg("kin_cbf", init_mode=0)
    for it in range(3):
        c_p = np.concatenate((x0, xs))
        init_control = np.concatenate((u0.reshape(-1, 1), next_states.reshape(-1, 1)))
        ref_traj, last_idx = ref.find_ref_traj(x0, xs, mpc.T_horizon, mpc.T_S, last_idx)
        solver = mpc.optimize_problem(ego_state=x0, ref_state=ref_traj, obstacle=obs)
        res = solver(x0=init_control, p=c_p, lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
        assert solver.stats()["success"] and solver.stats()["return_status"] == "Solve_Succeeded"
        sol = res["x"].full()
        assert sol.shape == (304, 1) and res["g"].full().shape == (303, 1)
        zo, _, info = c_oracle.solve(cfg, x0.ravel(), xs.ravel(), np.repeat(obs[:, None, :], 51, axis=1), init_control.ravel())
        assert info.status == 0 and abs(float(res["f"]) - info.f) <= COST_RTOL * info.f
        assert np.abs(sol.ravel()[:2] - zo[:2]).max() <= U0_ATOL
        g = res["g"].full().ravel()
        assert np.abs(g[:204]).max() <= 1e-7 and g[253:].min() >= -1e-7
        costs.append(float(res["f"]))
        u0 = sol[: N_p * nc].reshape(N_p, nc)
        x_m = sol[N_p * nc:].reshape(N_p + 1, ns)
        # shift_movement (PKG/main_cbf_kin_c_sim.py:16-26)
        x0 = x0 + mpc.T_S * mpc.f(x0, u0[0, :]).full()
        u0 = np.concatenate((u0[1:], u0[-1:]))
        next_states = np.concatenate((x_m[1:], x_m[-1:]), axis=0)
    assert abs(costs[0] - 1.0947508480e8) <= 1e-6 * 1.1e8


def test_batched_closed_loop_matches_oracle_loop(dev):
    """PKG/main_cbf_kin_c_sim_pre.py:86-126 for a batch: 15 MPC steps with a moving obstacle,
    warm starts and obstacle advance on the device, against the same loop around the CPU oracle."""
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.closed_loop import predict_obstacles, run_closed_loop
    from mpc_motion_planning_b200.Obs_prediction import obs_prediction_batch
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B, steps, N = 24, 15, 50
    x0, xs, obs = scenarios.kin_cbf_moving(B, seed=5)
    x0[0] = [0, 3, 0, 15]
    ob0 = obs[:, :, 0, :].copy()
    ob0[0, 0] = [50, 3.5, 0, 10, 4.8, 1.8]  # the reference's own scenario (main_cbf_kin_c_sim_pre.py:45-56)
    t = lambda a: torch.from_numpy(a).to(dev)
    assert np.array_equal(predict_obstacles(t(ob0), 0.1, N).cpu().numpy(), obs_prediction_batch(ob0, 0.1, N))
    s = BatchSolver("kin_cbf_pre")
    # the library's own batched obs_prediction (one kernel) and the mains' obstacle update: the reference's rows, bit for bit
    # up to the last ulp of device cos/sin against libm's
    st0 = t(ob0.copy())
    tr = s.predict_obstacles(st0, advance=True).cpu().numpy()
    host = obs_prediction_batch(ob0, 0.1, N)
    assert np.allclose(tr, host, rtol=0, atol=1e-12) and np.array_equal(tr[..., 2:], host[..., 2:])
    assert np.allclose(st0.cpu().numpy(), host[:, :, 1, :], rtol=0, atol=1e-12)  # advanced by one step = row 1 of the prediction
    out = run_closed_loop(s, t(x0), t(xs), t(ob0), steps)
    torch.cuda.synchronize()
    xg, ug, stg = out["x"].cpu().numpy(), out["u"].cpu().numpy(), out["status"].cpu().numpy()
    # oracle loop
    cfg = c_oracle.make_cfg("kin_cbf_pre")
    x, ob, z = x0.copy(), ob0.copy(), np.zeros((B, 304))
    alive = np.ones(B, bool)
    n_cmp = 0
    for k in range(steps):
        traj = obs_prediction_batch(ob, 0.1, N)
        u0, cost, st, it, zz = c_oracle.solve_batch(cfg, x, xs, traj, z_init=z, want_z=True, nthreads=os.cpu_count())
        alive &= (st <= 1) & (stg[k] <= 1)  # a failed solve makes the two loops diverge legitimately
        assert np.abs(ug[k][alive] - u0[alive]).max() <= U0_ATOL
        assert np.abs(xg[k][alive] - x[alive]).max() <= 1e-5
        n_cmp += alive.sum()
        U, X = zz[:, :100].reshape(B, 50, 2), zz[:, 100:].reshape(B, 51, 4)
        x = x + 0.1 * np.stack([x[:, 3] * np.cos(x[:, 2]), x[:, 3] * np.sin(x[:, 2]), x[:, 3] * np.tan(U[:, 0, 0]) / 2.6, U[:, 0, 1]], axis=1)
        z = np.concatenate([np.concatenate([U[:, 1:], U[:, -1:]], axis=1).reshape(B, -1),
                            np.concatenate([X[:, 1:], X[:, -1:]], axis=1).reshape(B, -1)], axis=1)
        ob = ob.copy()
        ob[..., 0] = ob[..., 0] + ob[..., 3] * np.cos(ob[..., 2]) * 0.1
        ob[..., 1] = ob[..., 1] + ob[..., 3] * np.sin(ob[..., 2]) * 0.1
    assert alive[0] and alive.mean() >= 0.6 and n_cmp >= 0.6 * B * steps
    # warm starts: later steps need far fewer iterations than the cold first one
    itg = out["iters"].cpu().numpy()
    assert itg[5:, alive].mean() < 0.7 * itg[0, alive].mean()


def test_reference_mains_run_closed_loop(dev):
    """mains/ = the reference's closed-loop mains on the drop-in modules (plots stripped)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, steps in (("main_cbf_kin_c_sim_pre", 80), ("main_cbf_dyn_c_sim", 100)):
        out = subprocess.run([sys.executable, f"{name}.py"], cwd=os.path.join(root, "mains"), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        text = " ".join(out.stdout.split())  # the 6-state summary wraps over two lines
        assert f"{name}: {steps} MPC steps, {steps} solved" in text, text[-400:]
    # the single-shot main (PKG/main_kin_s_sim.py): one solve from the all-zero guess, full result dict
    out = subprocess.run([sys.executable, "main_kin_s_sim.py"], cwd=os.path.join(root, "mains"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "main_kin_s_sim: 1 MPC steps, 2 solved" in " ".join(out.stdout.split()), out.stdout[-400:]


def test_in_kernel_obstacle_prediction_equals_host_trajectories(dev):
    """cfg.obs_input = MPCB_OBS_INITIAL: the library runs PKG/Obs_prediction.py's recursion itself."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    B = 200
    x0, xs, obs = scenarios.kin_cbf_moving(B, seed=21)
    a = _gpu(BatchSolver("kin_cbf_pre"), dev, x0, xs, obs)
    s2 = BatchSolver("kin_cbf_pre", obs_input="initial")
    import torch

    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)
    o = s2.solve(t(x0), t(xs), t(obs[:, :, 0, :]))
    torch.cuda.synchronize()
    b = {k: v.cpu().numpy() for k, v in o.items()}
    same = (a["status"] <= 1) & (b["status"] <= 1)
    assert ((a["status"] <= 1) == (b["status"] <= 1)).mean() >= 0.98 and same.mean() > 0.8
    # identical up to the last ulp of the device cos/sin in the obstacle increments
    assert np.abs(a["u0"] - b["u0"])[same].max() <= 1e-7
    assert (np.abs(a["cost"] - b["cost"])[same] <= 1e-9 * np.abs(a["cost"][same])).all()
    h = s2.solve(x0, xs, obs[:, :, 0, :])  # host entry with the compact layout
    assert np.array_equal(h["status"], b["status"]) and np.array_equal(h["cost"], b["cost"])


def _stage_reference(rng, xs, N):
    """per-stage cost targets around xs: what aa*ref_state[i+1] + (1-aa)*xs produces for a lane-change
    reference (PKG/MPC_CBF_optimize_kin.py:194-199)"""
    B = xs.shape[0]
    ref = np.repeat(xs[:, None, :], N, axis=1).copy()
    ramp = np.linspace(0.0, 1.0, N)[None, :]
    ref[:, :, 1] += rng.uniform(-1.0, 1.0, (B, 1)) * (1.0 - ramp)   # lateral offset that decays over the horizon
    ref[:, :, 3] += rng.uniform(-3.0, 3.0, (B, 1)) * (1.0 - ramp)
    return ref


@pytest.mark.parametrize("gamma,ref", [(1.0, "terminal"), (0.4, "terminal"), (None, "trajectory"), (0.6, "trajectory")])
def test_discrete_cbf_rows_and_stage_reference_parity(dev, gamma, ref):
    """the reference's commented `gamma*h_func + h_dot` row
    (PKG/MPC_CBF_optimize_kin_pre.py:250-254) and the `aa` reference blend (:194-199)."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B = 384
    x0, xs, obs = scenarios.kin_cbf_moving(B)
    rng = np.random.default_rng(77)
    xs_in = _stage_reference(rng, xs, 50) if ref == "trajectory" else xs
    s = BatchSolver("kin_cbf_pre", cbf_gamma=gamma, ref=ref)
    g = _gpu(s, dev, x0, xs_in, obs)
    cfg = c_oracle.make_cfg("kin_cbf_pre", cbf_gamma=gamma, ref_trajectory=(ref == "trajectory"))
    u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs_in, obs, nthreads=os.cpu_count())
    both, same = _check(g, u0, cost, st, 0.6)
    # same algorithm; device sin/cos/log differ from libm in the last ulp, which moves a filter or
    # regularisation decision on ~10% of the scenarios (same solutions, checked above)
    assert (g["iters"][both] == it[both]).mean() >= 0.8


def test_discrete_cbf_solution_satisfies_its_rows(dev):
    """Feasibility of the returned z against the numpy statement of the rows (oracle/nlp.py)."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle.nlp import NLP

    B = 64
    x0, xs, obs = scenarios.kin_cbf_moving(B)
    g = _gpu(BatchSolver("kin_cbf_pre", cbf_gamma=0.5), dev, x0, xs, obs, return_z=True)
    n = 0
    for b in np.where(g["status"] <= 1)[0][:24]:
        nlp = NLP("kin_cbf_pre", x0[b], xs[b], [obs[b, 0]], cbf_gamma=0.5)
        z = g["z"][b]
        assert np.abs(nlp.eq(z)).max() <= 1e-6
        d = nlp.ineq(z)
        assert np.all(d >= nlp.dL - 1e-6) and np.all(d <= nlp.dU + 1e-6)
        assert abs(nlp.objective(z) - g["cost"][b]) <= 1e-9 * abs(g["cost"][b])
        n += 1
    assert n >= 8


def test_reference_surface_with_the_switched_off_options(dev, tmp_path, monkeypatch):
    """The drop-in class with the reference's two dormant switches turned on: the commented
    `gamma*h_func + h_dot` row and `aa != 0` (PKG/MPC_CBF_optimize_kin_pre.py:194-199,250-254), with
    the reference's own `find_ref_traj` output as ref_state."""
    monkeypatch.chdir(tmp_path)
    from mpc_motion_planning_b200 import MPC_CBF_optimize_kin_pre, RefPathGenerator
    from mpc_motion_planning_b200.Obs_prediction import obs_prediction
    from oracle import c_oracle

    mpc = MPC_CBF_optimize_kin_pre.MPC_optimize()
    mpc.aa, mpc.gamma, mpc.cbf_rows, mpc.init = 0.25, 0.6, "dcbf", "rollout"
    N = mpc.N_p
    x0 = np.array([0, 3, 0, 15.0]).reshape(-1, 1)
    xs = np.array([400, 3.5, 0, 30.0]).reshape(-1, 1)
    rp = RefPathGenerator.RefPathGenerator()
    rp.define_ref_path(x0, xs, mpc.T_S)
    ref_traj, _ = rp.find_ref_traj(x0, xs, mpc.T_horizon, mpc.T_S, 0)
    obs = [np.array([[50, 3.5, 0, 10, 4.8, 1.8]])]
    tr = obs_prediction(obs, mpc.T_S, N)
    lbg, ubg, lbx, ubx = mpc.initialize_constraints(obs)
    solver = mpc.optimize_problem(ego_state=x0, ref_state=ref_traj, obs_trajectories=tr)
    res = solver(x0=np.zeros((2 * N + 4 * (N + 1), 1)), p=np.concatenate((x0, xs)), lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
    assert solver.stats()["success"]
    xref = mpc.aa * ref_traj[1: N + 1] + (1 - mpc.aa) * xs.ravel()[None, :]
    cfg = c_oracle.make_cfg("kin_cbf_pre", cbf_gamma=0.6, ref_trajectory=True)
    zo, _, info = c_oracle.solve(cfg, x0.ravel(), xref, np.array(tr))
    assert info.status == 0 and abs(float(res["f"]) - info.f) <= COST_RTOL * info.f
    assert np.abs(res["x"].full().ravel()[:2] - zo[:2]).max() <= U0_ATOL
    g = res["g"].full().ravel()
    assert g.shape == (303,) and np.abs(g[:204]).max() <= 1e-7 and g[253:].min() >= -1e-7


@pytest.mark.parametrize("kind,gen,N", [("kin_cbf_pre", "kin_cbf_moving", 2), ("kin_cbf_pre", "kin_cbf_moving", 128),
                                         ("dyn", "dyn_static", 3), ("dyn", "dyn_static", 128), ("kin_nocbf", "kin_nocbf", 128)])
def test_horizon_limits(dev, kind, gen, N):
    """Smallest and largest horizons the library accepts (2 <= N <= MPCB_NMAX = 128): the warps-per-block
    choice changes with the shared-memory footprint, the results must not."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B = 48
    g_ = getattr(scenarios, gen)
    x0, xs, obs = g_(B, N=N)
    s = BatchSolver(kind, N=N)
    g = _gpu(s, dev, x0, xs, obs)
    cfg = c_oracle.make_cfg(kind, N=N)
    u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs, obs if obs.shape[1] else None, nthreads=os.cpu_count())
    both = (g["status"] <= 1) & (st <= 1)
    differ = np.where((g["status"] <= 1) != (st <= 1))[0]
    assert differ.size <= 2, (differ.tolist(), g["status"][differ].tolist(), st[differ].tolist())  # 48 scenarios: 1.5 % + 1
    assert both.sum() >= 24
    du = np.abs(g["u0"] - u0).max(axis=1)
    dc = np.abs(g["cost"] - cost) / np.maximum(np.abs(cost), 1.0)
    assert np.all(du[both] <= U0_ATOL) and np.all(dc[both] <= COST_RTOL), (du[both].max(), dc[both].max())


def test_device_ref_traj_equals_the_host_generator(dev):
    """`mpcb_ref_traj_batch` against the host RefPathGenerator mirror (itself equal to the reference's
    output, tests/test_reference_vectors.py): index arithmetic bit for bit, both driving directions,
    `last_idx` carried over several calls like the mains' loops do."""
    import torch

    from mpc_motion_planning_b200 import RefPathGenerator
    from mpc_motion_planning_b200.solver import BatchSolver

    rng = np.random.default_rng(11)
    B, N, T_h, aa = 96, 50, 5, 0.3
    s = BatchSolver("kin_cbf_pre", ref="trajectory")
    start = rng.uniform(0, 20, B)
    xs = np.c_[np.where(np.arange(B) % 7 == 0, start - rng.uniform(50, 300, B), start + rng.uniform(100, 500, B)),
               rng.uniform(0, 4, B), rng.uniform(-0.05, 0.05, B), rng.uniform(10, 30, B)]
    gens = []
    for b in range(B):
        g = RefPathGenerator.RefPathGenerator()
        g.define_ref_path(np.array([start[b], 0, 0, 0.0]).reshape(-1, 1), xs[b].reshape(-1, 1), 0.1)
        gens.append(g)
    last = np.zeros(B, dtype=np.int32)
    t_last = torch.zeros(B, dtype=torch.int32, device=dev)
    t_xs, t_start = torch.from_numpy(xs).to(dev), torch.from_numpy(start).to(dev)
    x = np.c_[start, rng.uniform(0, 4, B), np.zeros(B), rng.uniform(5, 30, B)]
    for it in range(6):
        ref, stage = s.ref_traj(torch.from_numpy(x).to(dev), t_xs, t_start, t_last, T_h, aa)
        torch.cuda.synchronize()
        ref, stage = ref.cpu().numpy(), stage.cpu().numpy()
        for b in range(B):
            want, last[b] = gens[b].find_ref_traj(x[b].reshape(-1, 1), xs[b].reshape(-1, 1), T_h, 0.1, int(last[b]))
            assert np.array_equal(ref[b], want), (it, b)
            assert np.array_equal(stage[b], aa * want[1:] + (1 - aa) * xs[b][None, :]), (it, b)
        assert np.array_equal(t_last.cpu().numpy(), last)
        x[:, 0] += np.sign(xs[:, 0] - start) * rng.uniform(0.5, 6.0, B)  # drive along the path
        x[:, 1] += rng.normal(0, 0.1, B)


def test_closed_loop_with_stage_reference_on_device(dev):
    """closed loop with aa != 0: find_ref_traj + solve + plant step + shift, all on the device, against
    the same loop written with the host generator and the C oracle."""
    import torch

    from mpc_motion_planning_b200 import RefPathGenerator, scenarios
    from mpc_motion_planning_b200.closed_loop import run_closed_loop
    from mpc_motion_planning_b200.Obs_prediction import obs_prediction_batch
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B, steps, N, aa = 8, 6, 50, 0.3
    x0, xs, obs = scenarios.kin_cbf_moving(B, seed=5)
    obs0 = obs[:, :, 0, :].copy()
    s = BatchSolver("kin_cbf_pre", ref="trajectory", init="rollout")
    out = run_closed_loop(s, torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev), torch.from_numpy(obs0).to(dev), steps, aa=aa)
    torch.cuda.synchronize()
    X = out["x"].cpu().numpy()
    cfg = c_oracle.make_cfg("kin_cbf_pre", ref_trajectory=True)
    from mpc_motion_planning_b200 import MPC_CBF_optimize_kin_pre

    f = MPC_CBF_optimize_kin_pre.MPC_optimize().f
    ok = 0
    for b in range(B):
        g = RefPathGenerator.RefPathGenerator()
        g.define_ref_path(x0[b].reshape(-1, 1), xs[b].reshape(-1, 1), 0.1)
        x, o, z, last = x0[b].copy(), obs0[b, 0].copy(), None, 0
        good = True
        for k in range(steps):
            ref, last = g.find_ref_traj(x.reshape(-1, 1), xs[b].reshape(-1, 1), 5, 0.1, last)
            tgt = aa * ref[1:] + (1 - aa) * xs[b][None, :]
            zz, _, info = c_oracle.solve(cfg, x, tgt, obs_prediction_batch(o, 0.1, N)[None], z)
            if info.status > 1 or out["status"][k, b].item() > 1:
                good = False
                break
            x = x + 0.1 * f(x, zz[:2]).full().ravel()
            U, Xs = zz[: 2 * N].reshape(N, 2), zz[2 * N:].reshape(N + 1, 4)
            z = np.concatenate([np.concatenate([U[1:], U[-1:]]).ravel(), np.concatenate([Xs[1:], Xs[-1:]]).ravel()])
            o[0] += o[3] * np.cos(o[2]) * 0.1
            o[1] += o[3] * np.sin(o[2]) * 0.1
            assert np.abs(X[k + 1, b] - x).max() <= 1e-5, (b, k)
        ok += good
    assert ok >= B // 2


def test_results_do_not_depend_on_the_queue_order(dev):
    """`mpcb_set_order` is a scheduling hint only."""
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    B = 3000
    x0, xs, obs = scenarios.kin_cbf_moving(B)
    s = BatchSolver("kin_cbf_pre")
    a = _gpu(s, dev, x0, xs, obs)
    order = torch.argsort(torch.from_numpy(a["iters"]).to(dev), descending=True, stable=True).to(torch.int32)
    s.set_order(order)
    b = _gpu(s, dev, x0, xs, obs)
    s.set_order(None)
    for k in ("u0", "cost", "status", "iters"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_static_obstacle_rows_equal_repeated_trajectories(dev):
    """`obs_input="static"` ([B][M][6], what the static module's optimize_problem takes) gives bit-identical
    results to the same rows repeated over the horizon; the velocity column is ignored."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    B = 256
    x0, xs, obs = scenarios.kin_cbf_static(B)
    a = _gpu(BatchSolver("kin_cbf"), dev, x0, xs, obs)
    rows = obs[:, :, 0, :].copy()
    rows[:, :, 3] = 7.0  # a velocity must not move a static obstacle
    b = _gpu(BatchSolver("kin_cbf", obs_input="static"), dev, x0, xs, rows)
    for k in ("u0", "cost", "status", "iters"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_the_binding_stub_of_integration_md_runs(dev, tmp_path, monkeypatch):
    """INTEGRATION.md, option B: the ctypes stub a maintainer would add to the reference, executed as
    written (only the library path is made absolute) against the reference's default scenario."""
    import re

    monkeypatch.chdir(tmp_path)
    from mpc_motion_planning_b200 import MPC_CBF_optimize_kin, _lib

    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    code = re.search(r"```python\n# PKG/mpcb_binding.py.*?\n(.*?)```", text, re.S).group(1)
    code = code.replace('C.CDLL("libmpcb200.so")', f'C.CDLL("{_lib.SO_PATH}")')
    ns = {}
    exec(compile(code, "mpcb_binding.py", "exec"), ns)
    ns["lib"].mpcb_strerror.restype = __import__("ctypes").c_char_p
    mpc = MPC_CBF_optimize_kin.MPC_optimize()
    h = ns["make_handle"](mpc)
    N = mpc.N_p
    x0, xs = np.array([0, 3, 0, 15.0]), np.array([400, 3.5, 0, 30.0])
    obs = np.array([[50, 3.5, 0, 8, 4.8, 1.8]])
    z0 = np.zeros(2 * N + 4 * (N + 1))
    X = [x0]
    for _ in range(N):  # dynamically consistent first guess (the stub starts the solver exactly at x0=)
        X.append(X[-1] + mpc.T_S * mpc.f(X[-1], [0.0, 0.0]).full().ravel())
    z0[2 * N:] = np.array(X).ravel()
    z, f, status = ns["solve"](h, N, x0, xs, obs, z0)
    assert status == 0 and z.shape == (304, 1)
    assert abs(f - 1.0947508480e8) <= 1e-6 * 1.1e8 and np.allclose(z[:2, 0], [0.03564617, 3.0], atol=1e-6)


@pytest.mark.parametrize("tag", ["kin", "pre", "dyn", "nocbf"])
def test_cuda_reaches_the_kkt_points_verified_on_the_reference_expressions(dev, tag):
    """tests/golden/reference_nlp.npz holds, for the first step of each reference main, a point whose KKT
    residuals were measured on the reference's own NLP expressions (tests/test_reference_vectors.py)."""
    from mpc_motion_planning_b200.solver import BatchSolver
    from test_reference_vectors import KKT_CASES, kkt_case_obs

    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_nlp.npz"))
    kind, x0, xs = KKT_CASES[tag]
    obs = kkt_case_obs(ref, tag)
    g = _gpu(BatchSolver(kind), dev, np.array([x0]), np.array([xs]), obs[None] if obs is not None else np.zeros((1, 0, 51, 6)), return_z=True)
    assert g["status"][0] == 0
    assert np.abs(g["z"][0] - ref[f"{tag}_kkt_z"]).max() <= 1e-5
    assert np.abs(g["u0"][0] - ref[f"{tag}_kkt_z"][:2]).max() <= U0_ATOL
    assert abs(g["cost"][0] - float(ref[f"{tag}_kkt_f_ref"])) <= COST_RTOL * float(ref[f"{tag}_kkt_f_ref"])


@pytest.mark.parametrize("gamma,M", [(None, 3), (0.5, 3), (None, 4), (0.5, 4)])
def test_three_and_four_obstacles(dev, gamma, M):
    """M up to MPCB_MMAX = 4 (the mains carry a commented three-obstacle list, PKG/main_cbf_kin_c_sim.py:52-53)."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B = 192
    x0, xs, oa = scenarios.kin_cbf_moving(B)
    _, _, ob = scenarios.kin_cbf_moving(B, seed=99)
    _, _, oc = scenarios.kin_cbf_moving(B, seed=123)
    ob[:, :, :, 0] += 60.0
    oc[:, :, :, 0] += 120.0
    _, _, od = scenarios.kin_cbf_moving(B, seed=321)
    od[:, :, :, 0] += 180.0
    obs3 = np.concatenate([oa, ob, oc, od][:M], axis=1)
    g = _gpu(BatchSolver("kin_cbf_pre", M=M, cbf_gamma=gamma), dev, x0, xs, obs3)
    u0, cost, st, it, _ = c_oracle.solve_batch(c_oracle.make_cfg("kin_cbf_pre", M=M, cbf_gamma=gamma), x0, xs, obs3, nthreads=os.cpu_count())
    _check(g, u0, cost, st, 0.5)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_parity_under_randomised_problem_data(dev, seed):
    """Nothing in the kernels is specialised to the reference's constants: weights, bounds, margins, step
    and horizon are drawn at random (every mpcb_cfg field the reference hard-codes) and the CUDA path is
    compared with the oracle configured with the same numbers."""
    import ctypes as C

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    rng = np.random.default_rng(1000 + seed)
    kind = ["kin_cbf_pre", "kin_nocbf", "dyn", "kin_cbf_pre", "kin_cbf", "kin_cbf_pre"][seed - 1]
    gamma = 0.7 if seed == 4 else None
    N = int(rng.choice([24, 40, 64]))
    nx = 6 if kind == "dyn" else 4
    base = BatchSolver(kind, N=N).cfg
    sc = lambda a, lo=-0.7, hi=0.7: [float(v * 10 ** rng.uniform(lo, hi)) for v in a]
    ov = {"T": float(rng.choice([0.08, 0.1, 0.12])), "Q": sc(list(base.Q)[:nx]), "R": sc(list(base.R)), "DR": sc(list(base.DR)),
          "u_lo": [-rng.uniform(0.45, 0.7), -rng.uniform(2.0, 4.0)], "u_hi": [rng.uniform(0.45, 0.7), rng.uniform(2.0, 4.0)],
          "safe_l": float(rng.uniform(0.5, 1.5)), "safe_w": float(rng.uniform(0.3, 0.8)), "mu_init": float(rng.choice([10.0, 100.0, 1000.0]))}
    xl, xh = list(base.x_lo)[:nx], list(base.x_hi)[:nx]
    xl[1], xh[1], xh[3] = -rng.uniform(0.8, 1.5), rng.uniform(4.8, 6.0), rng.uniform(32.0, 45.0)
    ov["x_lo"], ov["x_hi"] = xl, xh
    if base.n_rate:
        k = rng.uniform(0.6, 2.0)
        ov["rate_lo"], ov["rate_hi"] = [v * k for v in list(base.rate_lo)], [v * k for v in list(base.rate_hi)]
    s = BatchSolver(kind, N=N, cbf_gamma=gamma, cfg_overrides=ov)
    ocfg = c_oracle.make_cfg(kind, N=N, cbf_gamma=gamma)
    for name, _ in c_oracle.OrcCfg._fields_:  # same numbers on both sides, field by field
        if name in ("obs_mode", "cbf_gamma", "ref_mode", "rows_as_shipped", "init_mode"):
            continue
        v = getattr(s.cfg, name)
        if hasattr(v, "__len__"):
            for i in range(len(v)):
                getattr(ocfg, name)[i] = v[i]
        else:
            setattr(ocfg, name, v)
    gen = {"kin_cbf_pre": scenarios.kin_cbf_moving, "kin_cbf": scenarios.kin_cbf_static, "kin_nocbf": scenarios.kin_nocbf, "dyn": scenarios.dyn_static}[kind]
    B = 160
    x0, xs, obs = gen(B, N=N, seed=500 + seed)
    g = _gpu(s, dev, x0, xs, obs)
    u0, cost, st, it, _ = c_oracle.solve_batch(ocfg, x0, xs, obs if obs.shape[1] else None, nthreads=os.cpu_count())
    both = (g["status"] <= 1) & (st <= 1)
    assert both.sum() >= 0.4 * B, both.sum()
    assert ((g["status"] <= 1) == (st <= 1)).mean() >= 0.9
    du = np.abs(g["u0"] - u0).max(axis=1)
    dc = np.abs(g["cost"] - cost) / np.maximum(np.abs(cost), 1.0)
    assert np.all(du[both] <= U0_ATOL) and np.all(dc[both] <= COST_RTOL), (du[both].max(), dc[both].max())


def test_dyn_rows_as_shipped_parity(dev):
    """dyn with the bound lists exactly as the reference ships them: tied controls,
    relaxed x/y defects.  CUDA vs the C oracle in the same mode, which follows the dense specification of
    oracle.nlp.ShippedDynNLP (tests/test_oracle_solvers.py)."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B = 256
    x0, xs, obs = scenarios.dyn_static(B)
    s = BatchSolver("dyn", dyn_bounds="as_shipped")
    g = _gpu(s, dev, x0, xs, obs, return_z=True)
    u0, cost, st, it, z = c_oracle.solve_batch(c_oracle.make_cfg("dyn", rows_as_shipped=True), x0, xs, obs, want_z=True, nthreads=os.cpu_count())
    both, same = _check(g, u0, cost, st, 0.8)
    assert (g["iters"][both] == it[both]).mean() >= 0.8
    U = g["z"][both][:, :100].reshape(-1, 50, 2)
    assert np.abs(U - U[:, :1, :]).max() <= 1e-9  # one control pair for the whole horizon
    # x positions are weakly determined here (weight 10 against 1e5 on y, and the relaxation slacks absorb
    # small shifts): compare the whole vector loosely, the criteria proper (u0, cost) are checked above
    assert np.abs(g["z"][both] - z[both]).max() <= 1e-3
    # the reference main's first step: constant controls (0.00804, 3.0), f = 1.7116e8 (aligned: (0.0939, 3.0), 1.7219e8)
    o = np.zeros((1, 1, 51, 6))
    o[0, 0, :, 0], o[0, 0, :, 1] = 100, -3.5
    d = _gpu(s, dev, np.array([[0, 0, 0, 10, 0, 0.0]]), np.array([[600, 3.5, 0, 15, 0, 0.0]]), o)
    assert d["status"][0] == 0 and abs(d["cost"][0] - 1.7116419462e8) <= 1e-6 * 1.7e8 and np.allclose(d["u0"][0], [0.00803901, 3.0], atol=1e-6)


def test_dyn_drop_in_class_solves_what_it_is_given(dev, tmp_path, monkeypatch):
    """The dyn drop-in class returns the reference's own lbg/ubg (as shipped) by default and the solver call
    recognises from lbg which pairing it was handed: as shipped -> constant controls, aligned -> the intended NLP."""
    monkeypatch.chdir(tmp_path)
    from mpc_motion_planning_b200 import MPC_CBF_optimize_dyn

    mpc = MPC_CBF_optimize_dyn.MPC_optimize()
    mpc.init = "rollout"
    x0 = np.array([0, 0, 0, 10, 0, 0.0]).reshape(-1, 1)
    xs = np.array([600, 3.5, 0, 15, 0, 0.0]).reshape(-1, 1)
    obs = np.array([100, -3.5])
    N = mpc.N_p
    z0 = np.zeros((2 * N + 6 * (N + 1), 1))
    res = {}
    for mode in ("as_shipped", "aligned"):
        mpc.dyn_bounds = mode
        lbg, ubg, lbx, ubx = mpc.initialize_constraints()
        solver = mpc.optimize_problem(ego_state=x0, ref_state=xs, obstacle=obs)
        r = solver(x0=z0, p=np.concatenate((x0, xs)), lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
        assert solver.stats()["success"]
        g = r["g"].full().ravel()
        assert np.all(g >= np.array(lbg) - 1e-6) and np.all(g <= np.array(ubg) + 1e-6)  # feasible for the lists it was given
        res[mode] = (float(r["f"]), r["x"].full().ravel())
    assert abs(res["as_shipped"][0] - 1.7116419462e8) <= 1e-6 * 1.7e8 and abs(res["aligned"][0] - 1.7218596389e8) <= 1e-6 * 1.7e8
    U = res["as_shipped"][1][: 2 * N].reshape(N, 2)
    assert np.abs(U - U[0]).max() <= 1e-9 and np.allclose(U[0], [0.00803901, 3.0], atol=1e-6)
    assert np.allclose(res["aligned"][1][:2], [0.09392595, 3.0], atol=1e-5)


def test_dyn_drop_in_zero_guess_switch(dev, tmp_path, monkeypatch):
    """The dyn class announces that it re-integrates the reference's all-zero state guess (PKG/main_cbf_dyn_c_sim.py:47-50)
    and `zero_guess = "as_given"` hands the caller's x0= to the solver untouched: vx = 0 puts the tire model's
    linearisation out of the range the stage-wise recursion can condense, so that call reports failure, loudly."""
    import warnings

    monkeypatch.chdir(tmp_path)
    from mpc_motion_planning_b200 import MPC_CBF_optimize_dyn

    mpc = MPC_CBF_optimize_dyn.MPC_optimize()
    x0 = np.array([0, 0, 0, 10, 0, 0.0]).reshape(-1, 1)
    xs = np.array([600, 3.5, 0, 15, 0, 0.0]).reshape(-1, 1)
    N = mpc.N_p
    z0 = np.zeros((2 * N + 6 * (N + 1), 1))
    lbg, ubg, lbx, ubx = mpc.initialize_constraints()
    solver = mpc.optimize_problem(ego_state=x0, ref_state=xs, obstacle=np.array([100, -3.5]))
    with pytest.warns(RuntimeWarning, match="all-zero state guess"):
        r = solver(x0=z0, p=np.concatenate((x0, xs)), lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
    assert solver.stats()["success"] and abs(float(r["f"]) - 1.7116419462e8) <= 1e-6 * 1.7e8
    mpc.zero_guess = "as_given"
    solver = mpc.optimize_problem(ego_state=x0, ref_state=xs, obstacle=np.array([100, -3.5]))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        solver(x0=z0, p=np.concatenate((x0, xs)), lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
    st = solver.stats()
    assert not st["success"] and st["return_status"] in ("Infeasible_Problem_Detected", "Restoration_Failed", "Maximum_Iterations_Exceeded", "Invalid_Number_Detected")


def test_plain_c_example_runs(dev, tmp_path):
    """examples/batch_solve.c: a C program against include/mpcb200.h, no Python in the loop."""
    import subprocess

    from mpc_motion_planning_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "batch_solve")
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "batch_solve.c"), "-L",
                    os.path.dirname(_lib.SO_PATH), "-lmpcb200", "-lm", "-o", exe], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(_lib.SO_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    run = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert run.returncode == 0 and run.stdout.strip().endswith("OK"), run.stdout + run.stderr


@pytest.mark.gpu
def test_pipelined_batches_equal_one_at_a_time():
    """PipelinedSolver (two handles, batches in flight on two streams; mpcb_submit_batch_host / mpcb_wait for host
    buffers) returns, for every batch, exactly what BatchSolver returns for it alone."""
    import torch
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.pipeline import PipelinedSolver
    from mpc_motion_planning_b200.solver import BatchSolver
    dev = torch.device("cuda:0")
    batches = [scenarios.kin_cbf_moving(B, seed=900 + i) for i, B in enumerate((700, 64, 2500, 1, 1300))]
    one = BatchSolver("kin_cbf_pre")
    want = []
    for x0, xs, obs in batches:
        o = one.solve(*(torch.from_numpy(v).to(dev) for v in (x0, xs, obs)))
        want.append({k: v.cpu() for k, v in o.items()})
    pipe = PipelinedSolver(2, "kin_cbf_pre")
    # device tensors: submit everything, then collect
    tickets, got = [], []
    for i, (x0, xs, obs) in enumerate(batches):
        if i >= pipe.lanes:                       # a lane holds one batch: collect before reusing it
            got.append({k: v.clone() for k, v in pipe.result(tickets[i - pipe.lanes]).items()})
        tickets.append(pipe.submit(*(torch.from_numpy(v).to(dev) for v in (x0, xs, obs))))
    for t in tickets[-pipe.lanes:]:
        got.append({k: v.clone() for k, v in pipe.result(t).items()})
    torch.cuda.synchronize()
    for w, g in zip(want, got):
        for k in ("u0", "cost", "status", "iters"):
            assert torch.equal(w[k], g[k].cpu()), k
    # page-locked host buffers through the C-ABI pair
    outs, tick = [], []
    for i, (x0, xs, obs) in enumerate(batches):
        B = x0.shape[0]
        h = [torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for v in (x0, xs, obs)]
        o = (torch.empty((B, 2), dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory(),
             torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
        tick.append(pipe.submit_host(B, h[0], h[1], h[2], None, *o))
        outs.append((h, o))
    pipe.wait()
    for w, (_, o) in zip(want, outs):
        assert torch.equal(w["u0"], o[0]) and torch.equal(w["cost"], o[1])
        assert torch.equal(w["status"], o[2]) and torch.equal(w["iters"], o[3])
    pipe.close()


@pytest.mark.gpu
def test_closed_loop_in_lanes_equals_the_single_handle_loop():
    """run_closed_loop_lanes (fleet split over two handles/streams, steps interleaved) = run_closed_loop, bit for bit."""
    import torch
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.closed_loop import run_closed_loop, run_closed_loop_lanes
    from mpc_motion_planning_b200.solver import BatchSolver
    dev = torch.device("cuda:0")
    x0, xs, obs = scenarios.kin_cbf_moving(301, seed=4242)
    tx0, txs = torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev)
    obs0 = torch.from_numpy(obs[:, :, 0, :].copy()).to(dev)
    one = run_closed_loop(BatchSolver("kin_cbf_pre"), tx0, txs, obs0, 6)
    two = run_closed_loop_lanes([BatchSolver("kin_cbf_pre") for _ in range(2)], tx0, txs, obs0, 6)
    torch.cuda.synchronize()
    for k in ("x", "u", "status", "iters"):
        assert one[k].shape == two[k].shape and torch.equal(one[k], two[k]), k


# --------------------------------------------------------------------------- ABI 0.2.0
@pytest.mark.parametrize("kind,gen", [("kin_nocbf", "kin_nocbf"), ("kin_cbf", "kin_cbf_static"), ("kin_cbf_pre", "kin_cbf_moving"),
                                      ("dyn", "dyn_static")])
def test_lam_g_and_lam_x_satisfy_stationarity_on_the_restated_nlp(dev, kind, gen):
    """res['lam_g'] (every row of g, the reference's order) and res['lam_x'] (CasADi's sign) of converged scenarios make
    grad f + J_g' lam_g + lam_x vanish on the numpy restatement of the NLP, and obey the sign rules of their bounds."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import nlp as onlp

    B = 24
    x0, xs, obs = getattr(scenarios, gen)(B)
    s = BatchSolver(kind)
    g = _gpu(s, dev, x0, xs, obs, return_z=True, return_lam=True, return_duals=True)
    assert g["lam_g"].shape == (B, s.n_g) and g["lam_x"].shape == (B, s.nv)
    n = 0
    for b in np.where(g["status"] == 0)[0][:8]:
        if kind == "kin_cbf":
            ob = obs[b, :, 0, :]
        elif kind == "kin_cbf_pre":
            ob = [obs[b, j] for j in range(obs.shape[1])]
        elif kind == "dyn":
            ob = obs[b, 0, 0, :2]
        else:
            ob = None
        P = onlp.NLP(kind, x0[b], xs[b], ob)
        z = g["z"][b]
        J = np.vstack([P.jac_eq(z), P.jac_ineq(z)] if P.n_ineq else [P.jac_eq(z)])[P.g_perm()]
        lam_g, lam_x = g["lam_g"][b], g["lam_x"][b]
        r = P.grad(z) + J.T @ lam_g + lam_x
        scale = max(1.0, np.abs(P.grad(z)).max())
        assert np.abs(r).max() <= 1e-6 * scale, (kind, b, np.abs(r).max(), scale)
        # head of lam_g = the equality multipliers already exported as lam_out
        if kind != "dyn":
            assert np.array_equal(lam_g[: P.n_eq], g["lam"][b])
        lo, hi = P.lbg_ubg_aligned()
        ineq = np.isinf(hi) & np.isfinite(lo)
        assert np.all(lam_g[ineq] <= 1e-9 * scale)  # g >= lb rows: multiplier <= 0 in CasADi's convention
        free = np.isinf(P.zL) & np.isinf(P.zU)
        assert np.all(lam_x[free] == 0.0)
        n += 1
    assert n >= 4


def test_two_streams_on_one_handle_are_serialised(dev):
    """A handle owns one work queue and one slab: a solve issued on a second stream queues behind the first on the device
    instead of racing it (include/mpcb200.h)."""
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    B = 3000
    x0, xs, obs = scenarios.kin_cbf_static(2 * B)
    s = BatchSolver("kin_cbf")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    a = [t(v[:B]) for v in (x0, xs, obs)]
    b = [t(v[B:]) for v in (x0, xs, obs)]
    ref_a = {k: v.clone() for k, v in s.solve(*a).items()}
    ref_b = {k: v.clone() for k, v in s.solve(*b).items()}
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(3):
        with torch.cuda.stream(s1):
            oa = s.solve(*a)
        with torch.cuda.stream(s2):
            ob = s.solve(*b)
        torch.cuda.synchronize()
        for k in ("u0", "cost", "status", "iters"):
            assert torch.equal(oa[k], ref_a[k]) and torch.equal(ob[k], ref_b[k]), k


def test_reserve_and_order_length(dev):
    import torch

    from mpc_motion_planning_b200 import _lib, scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    x0, xs, obs = scenarios.kin_cbf_static(64)
    s = BatchSolver("kin_cbf")
    s.reserve(64)
    free0 = torch.cuda.mem_get_info(dev)[0]
    h = s.solve(x0, xs, obs)  # numpy -> host entry, staged through the reserved buffers
    assert torch.cuda.mem_get_info(dev)[0] == free0, "the host entry allocated although the batch was reserved"
    g = _gpu(s, dev, x0, xs, obs)
    assert np.array_equal(h["status"], g["status"]) and np.array_equal(h["u0"], g["u0"])
    # an order is a permutation of exactly one batch size
    s.set_order(torch.arange(32, dtype=torch.int32, device=dev))
    with pytest.raises(_lib.MpcbError):
        _gpu(s, dev, x0, xs, obs)
    s.set_order(None)
    _gpu(s, dev, x0, xs, obs)


# --------------------------------------------------------------------------- restoration phase, verdict at scale
@pytest.mark.parametrize("kind,gen", [("kin_cbf", "kin_cbf_static"), ("kin_cbf_pre", "kin_cbf_moving")])
@pytest.mark.parametrize("calls", [1, 0])
def test_restoration_phase_parity(dev, kind, gen, calls):
    """cfg.restoration: a failed line search enters the restoration phase (sibling kernel, second pass) instead of ending
    with status 3.  Same algorithm as the C oracle / the dense specification: statuses, verdicts and answers agree, the
    success rate does not drop, and no scenario is left with the 'no restoration attempted' exit."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B = 3000
    x0, xs, obs = getattr(scenarios, gen)(B)
    plain = _gpu(BatchSolver(kind), dev, x0, xs, obs)
    s = BatchSolver(kind, restoration=True, resto_max_calls=calls)
    g = _gpu(s, dev, x0, xs, obs)
    u0, cost, st, it, _ = c_oracle.solve_batch(c_oracle.make_cfg(kind, restoration=True, resto_max_calls=calls), x0, xs, obs,
                                               nthreads=os.cpu_count())
    print(f"{kind} calls={calls}: gpu {np.bincount(g['status'], minlength=6)} oracle {np.bincount(st, minlength=6)} "
          f"without restoration {np.bincount(plain['status'], minlength=6)}")
    both, same = _check(g, u0, cost, st, 0.8)
    assert (g["iters"][both] == it[both]).mean() >= 0.9
    # restoration rescues scenarios, it never loses a success of the plain run (up to the chaotic handful)
    assert (g["status"] <= 1).sum() >= (plain["status"] <= 1).sum()
    lost = np.where((plain["status"] <= 1) & (g["status"] > 1))[0]
    assert lost.size <= 3, lost
    # scenarios the plain run solves are untouched by the second pass (they never enter it)
    keep = (plain["status"] <= 1) & (g["status"] <= 1)
    assert np.array_equal(plain["u0"][keep], g["u0"][keep]) and np.array_equal(plain["iters"][keep], g["iters"][keep])
    if calls == 0:  # no cap: every failure is a maximum-iterations exit, a detected local infeasibility or a failed restoration
        assert set(np.unique(g["status"])) <= {0, 1, 2, 3, 5}
        assert (g["status"] == 3).sum() <= (plain["status"] == 3).sum()
    assert s.launch_info()["launches"] >= 2  # main kernel + restoration sibling


def test_verdict_agreement_on_the_bench_batch(dev):
    """The driver-run benchmark's exact inputs (bench.py: kin_cbf_static(10000, seed BASE+2), obstacle rows as the static
    module takes them): nothing outside tolerance among the commonly converged scenarios, verdict equal on >= 99 %."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B = 10000
    x0, xs, obs_traj = scenarios.kin_cbf_static(B, seed=scenarios.BASE_SEED + 2)
    obs = np.ascontiguousarray(obs_traj[:, :, 0, :])
    g = _gpu(BatchSolver("kin_cbf", obs_input="static"), dev, x0, xs, obs)
    u0, cost, st, it, _ = c_oracle.solve_batch(c_oracle.make_cfg("kin_cbf"), x0, xs, obs_traj, nthreads=os.cpu_count())
    both, same = _check(g, u0, cost, st, 0.8, min_same_verdict=0.99)
    du = np.abs(g["u0"] - u0).max(axis=1)
    dc = np.abs(g["cost"] - cost) / np.abs(cost)
    assert np.all(du[both] <= U0_ATOL) and np.all(dc[both] <= COST_RTOL), (du[both].max(), dc[both].max())
    print(f"bench batch: verdict-equal {same.mean():.4f}, worst |du0| {du[both].max():.2e}, worst rel dcost {dc[both].max():.2e}")


# --------------------------------------------------------------------------- second engine: one scenario per lane
@pytest.mark.parametrize("kind,gen,M,B", [("kin_nocbf", "kin_nocbf", 1, 700), ("kin_cbf", "kin_cbf_static", 1, 1500),
                                           ("kin_cbf_pre", "kin_cbf_moving", 1, 1500), ("kin_cbf_pre", "kin_cbf_moving", 2, 300)])
def test_lane_engine_parity(dev, kind, gen, M, B):
    """cfg.engine = MPCB_ENGINE_LANE (csrc/mpcb_lane_kernel.cuh): every lane solves its own scenario, records
    structure-of-arrays in global memory.  Same algorithm: BASELINE's three criteria against the oracle, and against the
    warp-per-scenario engine on the same inputs (summation orders differ, so not bit for bit)."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    x0, xs, obs = getattr(scenarios, gen)(B)
    if M == 2:
        _, _, ob = scenarios.kin_cbf_moving(B, seed=99)
        ob[:, :, :, 0] += 60.0
        obs = np.concatenate([obs, ob], axis=1)
    lane = BatchSolver(kind, M=M, engine="lane")
    g = _gpu(lane, dev, x0, xs, obs, return_z=True, return_lam=True, return_duals=True)
    info = lane.launch_info()
    assert info["smem_bytes"] == 0 and info["block"] == 128  # the lane kernel ran, not the warp kernel
    w = _gpu(BatchSolver(kind, M=M, engine="warp"), dev, x0, xs, obs, return_z=True, return_lam=True, return_duals=True)
    u0, cost, st, it, _ = c_oracle.solve_batch(c_oracle.make_cfg(kind, M=M), x0, xs, obs if obs.shape[1] else None, nthreads=os.cpu_count())
    both, same = _check(g, u0, cost, st, 0.6 if M == 2 else 0.8)
    assert (g["iters"][both] == it[both]).mean() >= 0.9
    bw = (g["status"] <= 1) & (w["status"] <= 1)
    assert bw.mean() >= 0.6
    for key, tol in (("z", 1e-5), ("lam", None), ("lam_g", None), ("lam_x", None)):
        d = np.abs(g[key][bw] - w[key][bw]).max(axis=1)
        scale = np.maximum(1.0, np.abs(w[key][bw]).max(axis=1))
        assert np.quantile(d / scale, 0.995) <= (tol or 1e-5), (key, (d / scale).max())
    # not applicable configurations are refused, not silently served by the other engine
    from mpc_motion_planning_b200 import _lib
    with pytest.raises(_lib.MpcbError):
        BatchSolver("dyn", engine="lane")
    with pytest.raises(_lib.MpcbError):
        BatchSolver("kin_cbf", cbf_gamma=0.5, engine="lane")


def test_lane_engine_beyond_one_wave_and_auto_selection(dev):
    """More scenarios than resident lanes (the lanes refill from the queue, the grid is sized for equally full waves) and the
    automatic engine choice: the row-free family goes to the lane engine from 20,480 scenarios up, and every scenario gets the
    warp engine's answer either way."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    B = 70000
    x0, xs, obs = scenarios.kin_nocbf(B)
    auto = BatchSolver("kin_nocbf")  # engine="auto"
    g = _gpu(auto, dev, x0, xs, None)
    info = auto.launch_info()
    lanes = info["num_sms"] * info["blocks_per_sm"] * info["block"]
    assert info["smem_bytes"] == 0 and info["block"] == 128  # the lane kernel ran
    assert B > lanes and info["grid"] * info["block"] < lanes  # two waves, neither of them full
    assert info["grid"] == -(-(-(-B // 128)) // 2)
    w = _gpu(BatchSolver("kin_nocbf", engine="warp"), dev, x0, xs, None)
    assert np.array_equal(g["status"], w["status"]) and (g["status"] == 0).all()
    assert np.array_equal(g["iters"], w["iters"])
    assert np.abs(g["u0"] - w["u0"]).max() <= 1e-9 and np.abs(g["cost"] - w["cost"]).max() <= 1e-9 * np.abs(w["cost"]).max()
    small = _gpu(auto, dev, x0[:4096], xs[:4096], None)  # below the switch: the warp kernel
    assert auto.launch_info()["smem_bytes"] > 0
    assert np.array_equal(small["u0"], w["u0"][:4096])
    # warm start (init = as_given with the previous solution, the closed loop's protocol) through both engines
    n = 24000
    zw = _gpu(BatchSolver("kin_nocbf", engine="warp"), dev, x0[:n], xs[:n], None, return_z=True)["z"]
    x1 = x0[:n] + np.array([0.3, 0.02, 0.001, 0.1])
    res = {}
    for eng in ("warp", "auto"):
        sv = BatchSolver("kin_nocbf", engine=eng, init="as_given")
        res[eng] = _gpu(sv, dev, x1, xs[:n], None, z_init=zw, return_z=True)
        assert (sv.launch_info()["smem_bytes"] == 0) == (eng == "auto")
    assert np.array_equal(res["auto"]["status"], res["warp"]["status"]) and (res["warp"]["status"] == 0).all()
    assert np.abs(res["auto"]["z"] - res["warp"]["z"]).max() <= 1e-7 and res["warp"]["iters"].mean() < 14


def test_lane_engine_with_restoration_and_order(dev):
    """The lane engine hands failed line searches to the same restoration sibling, and honours a processing order."""
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    B = 2000
    x0, xs, obs = scenarios.kin_cbf_static(B)
    a = _gpu(BatchSolver("kin_cbf", engine="warp", restoration=True), dev, x0, xs, obs)
    s = BatchSolver("kin_cbf", engine="lane", restoration=True)
    s.set_order(torch.arange(B - 1, -1, -1, dtype=torch.int32, device=dev))
    b = _gpu(s, dev, x0, xs, obs)
    s.set_order(None)
    assert ((a["status"] <= 1) == (b["status"] <= 1)).mean() >= 0.985
    both = (a["status"] <= 1) & (b["status"] <= 1)
    assert np.abs(a["u0"][both] - b["u0"][both]).max() <= U0_ATOL
    assert (b["status"] == 3).sum() < 0.9 * (_gpu(BatchSolver("kin_cbf", engine="lane"), dev, x0, xs, obs)["status"] == 3).sum()


# --------------------------------------------------------------------------- Runge-Kutta shooting defects (cfg.integrator)
@pytest.mark.parametrize("kind,gen,M,B", [("kin_nocbf", "kin_nocbf", 1, 300), ("kin_cbf", "kin_cbf_static", 1, 1000),
                                           ("kin_cbf_pre", "kin_cbf_moving", 2, 300)])
def test_rk4_defects_parity_with_the_oracle_in_rk4_mode(dev, kind, gen, M, B):
    """cfg.integrator = MPCB_INTEGRATOR_RK4: X_{k+1} - Phi_RK4(X_k, U_k) with forward-mode Jacobians through the four
    stages and the exact second-order adjoint.  Checked against the oracle's own RK4 mode (oracle/nlp.py Rk4KinModel is
    the specification, its derivatives are checked against finite differences in the CPU suite); parity with the
    REFERENCE is Euler-only - the two integrators' optima differ by more than the tolerances, which is asserted too."""
    from mpc_motion_planning_b200 import _lib, scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    x0, xs, obs = getattr(scenarios, gen)(B)
    if M == 2:
        _, _, ob = scenarios.kin_cbf_moving(B, seed=99)
        ob[:, :, :, 0] += 60.0
        obs = np.concatenate([obs, ob], axis=1)
    s = BatchSolver(kind, M=M, integrator="rk4")
    g = _gpu(s, dev, x0, xs, obs, return_z=True)
    oobs = obs if obs.shape[1] else None
    u0, cost, st, it, z = c_oracle.solve_batch(c_oracle.make_cfg(kind, M=M, integrator="rk4"), x0, xs, oobs, want_z=True, nthreads=os.cpu_count())
    both, same = _check(g, u0, cost, st, 0.6 if M == 2 else 0.8)
    assert (g["iters"][both] == it[both]).mean() >= 0.9
    assert np.quantile(np.abs(g["z"][both] - z[both]).max(axis=1), 0.99) <= 1e-5
    # the returned trajectory satisfies the RK4 defects (fine Euler sub-stepping as an independent integrator)
    b = int(np.where(both)[0][0])
    N = 50
    U, X = g["z"][b, : 2 * N].reshape(N, 2), g["z"][b, 2 * N:].reshape(N + 1, 4)
    for k in (0, 17, 49):
        x = X[k].copy()
        for _ in range(2000):
            x = x + 5e-5 * np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), x[3] * np.tan(U[k, 0]) / 2.6, U[k, 1]])
        assert np.abs(x - X[k + 1]).max() <= 2e-4, (k, np.abs(x - X[k + 1]).max())
    # Euler and RK4 are different NLPs: their optima are further apart than BASELINE's tolerances
    e = _gpu(BatchSolver(kind, M=M, engine="lane"), dev, x0, xs, obs)
    be = both & (e["status"] <= 1)
    assert (np.abs(e["cost"][be] - g["cost"][be]) / np.abs(g["cost"][be])).max() > COST_RTOL
    # Runge-Kutta is not offered where it is not implemented
    for bad in (dict(kind="dyn"), dict(kind="kin_cbf", engine="warp"), dict(kind="kin_cbf", cbf_gamma=0.5), dict(kind="kin_cbf", restoration=True)):
        with pytest.raises(_lib.MpcbError):
            BatchSolver(bad.pop("kind"), integrator="rk4", **bad)


@pytest.mark.parametrize("tag", ["kin", "pre", "dyn", "nocbf"])
def test_cuda_reaches_64_kkt_points_per_nlp_verified_on_the_reference_expressions(dev, tag):
    """tests/golden/reference_nlp.npz `*_kkt64_*`: 64 seeded start states per module whose oracle solutions were checked,
    at generation time, to be KKT points of the NLP the REFERENCE builds (exact sympy derivatives of its own expressions:
    stationarity and complementarity <= 2e-8 in IPOPT's scaling, rows feasible to bound_relax_factor).  The CUDA path must
    land on those points."""
    from test_reference_vectors import KKT64_XS, KKT_CASES, kkt_case_obs

    from mpc_motion_planning_b200.solver import BatchSolver

    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_nlp.npz"), allow_pickle=True)
    kind = KKT_CASES[tag][0]
    x0 = ref[f"{tag}_kkt64_x0"]
    B = x0.shape[0]
    xs = np.tile(np.array(KKT64_XS[tag], float), (B, 1))
    obs = kkt_case_obs(ref, tag)
    ob = np.zeros((B, 0, 51, 6)) if obs is None else np.tile(obs[None], (B, 1, 1, 1))
    g = _gpu(BatchSolver(kind), dev, x0, xs, ob, return_z=True)
    ok = g["status"] == 0
    print(f"{tag}: {ok.sum()} of {B} converged; worst |dz| {np.abs(g['z'][ok] - ref[f'{tag}_kkt64_z'][ok]).max():.2e}")
    assert ok.sum() >= 62
    assert np.abs(g["z"][ok] - ref[f"{tag}_kkt64_z"][ok]).max() <= 1e-5
    assert np.abs(g["u0"][ok] - ref[f"{tag}_kkt64_z"][ok][:, :2]).max() <= U0_ATOL
    fr = ref[f"{tag}_kkt64_f_ref"][ok]
    assert (np.abs(g["cost"][ok] - fr) / np.abs(fr)).max() <= COST_RTOL


# --------------------------------------------------------------------------- sanitizer substitute
_SLOT_CHECK = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
t = lambda a: None if a is None or a.shape[1] == 0 else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
total = 0
for kind, gen, kw, B in (("kin_nocbf", "kin_nocbf", {}, 300), ("kin_cbf", "kin_cbf_static", {}, 3000), ("kin_cbf", "kin_cbf_static", {}, 200),
                         ("kin_cbf_pre", "kin_cbf_moving", dict(cbf_gamma=0.5), 600), ("kin_cbf_pre", "kin_cbf_moving", dict(N=100), 300),
                         ("kin_cbf", "kin_cbf_static", dict(restoration=True, resto_max_calls=0), 1500)):
    x0, xs, obs = getattr(scenarios, gen)(B, **({"N": kw["N"]} if "N" in kw else {}))
    s = BatchSolver(kind, engine="warp", **kw)
    out = s.solve(t(x0), t(xs), t(obs))
    torch.cuda.synchronize()
    n = s.debug_slot_errors()
    assert n >= 0, "not a -DMPCB_DEBUG_SLOTS build"
    print(kind, kw, B, "converged", int((out["status"] <= 1).sum()), "slot errors", n)
    total += n
print("TOTAL", total)
"""


def test_debug_build_sees_no_slot_ownership_violation(dev):
    """compute-sanitizer is closed on this pool; libmpcb200_debug.so (-DMPCB_DEBUG_SLOTS) tags every store into the aliased
    shared-memory slots (LAMP over CDEF; [HUU EE GU TK] -> gains -> slack steps / trial defects) and checks the tag at every
    load.  One solve per kernel family (persistent, small-batch all-shared, discrete-CBF, long-horizon layout, restoration
    sibling) must report zero violations - and the checker must trip when an expectation is made wrong on purpose."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "mpc_motion_planning_b200", "libmpcb200_debug.so")
    if not os.path.exists(so):
        pytest.skip("libmpcb200_debug.so not built (python -m mpc_motion_planning_b200.build --debug-slots)")
    env = dict(os.environ, MPCB200_LIB=so)
    out = subprocess.run([sys.executable, "-c", _SLOT_CHECK % root], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    print(out.stdout)
    assert "TOTAL 0" in out.stdout, out.stdout[-2000:]
    env["MPCB_DEBUG_SLOTS_SELFTEST"] = "1"
    out = subprocess.run([sys.executable, "-c", _SLOT_CHECK % root], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "TOTAL 0" not in out.stdout and "TOTAL" in out.stdout, out.stdout[-2000:]


# --------------------------------------------------------------------------- the reference's literal first guess
@pytest.mark.parametrize("kind,gen", [("kin_cbf", "kin_cbf_static"), ("kin_cbf_pre", "kin_cbf_moving"), ("kin_nocbf", "kin_nocbf")])
@pytest.mark.parametrize("mu_init", [30.0, 0.1])
def test_parity_from_the_all_zero_guess(dev, kind, gen, mu_init):
    """The reference's literal protocol: `x0=` all zeros, taken as given (PKG/main_cbf_kin_c_sim.py:47-50,92), with this
    library's mu_init 30 and with IPOPT's default 0.1.  From that guess most random scenarios do not converge within the
    reference's max_iter = 100 (`bench.py` carries the rates) - what is asserted is that the CUDA path and
    the oracle agree: same verdicts up to the chaotic handful, the same answers wherever both converge, the same iteration
    counts."""
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    from oracle import c_oracle

    B = 768
    x0, xs, obs = getattr(scenarios, gen)(B)
    s = BatchSolver(kind, init="as_given", mu_init=mu_init)
    g = _gpu(s, dev, x0, xs, obs)  # z_init = None: the all-zero guess
    cfg = c_oracle.make_cfg(kind, init_mode=0, mu_init=mu_init)
    u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs, obs if obs.shape[1] else None, nthreads=os.cpu_count())
    print(f"{kind} mu_init={mu_init}: gpu {np.bincount(g['status'], minlength=6)} oracle {np.bincount(st, minlength=6)} "
          f"mean iterations {g['iters'].mean():.1f} / {it.mean():.1f}")
    # from this guess the solves are long (50-85 iterations) and mostly end in a failed line search: more of them sit on the
    # chaotic boundary than from the roll-out start (measured 2.3 % verdict differences at mu_init 30 on kin-CBF static)
    both, same = _check(g, u0, cost, st, 0.0, min_same_verdict=0.95)
    if both.sum() >= 20:
        assert (g["iters"][both] == it[both]).mean() >= 0.85
    # the non-converged iterates (which the reference would consume unchecked) are NOT comparable between two
    # implementations: reported, not asserted
    print(f"   first controls within 1e-3 on {100 * (np.abs(g['u0'] - u0).max(axis=1) <= 1e-3).mean():.1f} % of ALL scenarios")
