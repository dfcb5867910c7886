"""The restated NLP, the bound lists and the host helpers against vectors produced by the
reference's OWN code (tests/golden/make_reference_vectors.py ran PKG/MPC_CBF_optimize_kin.py,
_kin_pre.py and _dyn.py unmodified - and the bytecode of the .pyc-only MPC_optimize_kin - on a
sympy-backed `casadi` stand-in and stored what they hand to `nlpsol`).  The IPOPT solve itself is not covered by these vectors."""
import json
import os

import numpy as np
import pytest

from oracle.nlp import NLP

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(HERE, "golden", "reference_nlp.npz"))


def _nlp(ref, tag, k):
    nx = 6 if tag == "dyn" else 4
    p = ref[f"{tag}_p"][k]
    if tag == "kin":
        return NLP("kin_cbf", p[:nx], p[nx:], ref["kin_obs"])
    if tag == "pre":
        return NLP("kin_cbf_pre", p[:nx], p[nx:], list(ref["pre_obs"]))
    if tag == "nocbf":
        return NLP("kin_nocbf", p[:nx], p[nx:])
    return NLP("dyn", p[:nx], p[nx:], ref["dyn_obs"])


@pytest.mark.parametrize("tag", ["kin", "pre", "dyn", "nocbf"])
def test_objective_and_constraints_equal_the_reference_expressions(ref, tag):
    """f(z,p) and g(z,p) in the reference's row order, two obstacles for the kinematic modules."""
    for k in range(ref[f"{tag}_z"].shape[0]):
        nlp, z = _nlp(ref, tag, k), ref[f"{tag}_z"][k]
        f, g = ref[f"{tag}_f"][k], ref[f"{tag}_g"][k]
        assert abs(nlp.objective(z) - f) <= 1e-14 * abs(f)
        assert g.shape == (nlp.n_eq + nlp.n_ineq,)
        assert np.max(np.abs(nlp.g_ref(z) - g)) <= 1e-11


@pytest.mark.parametrize("tag", ["kin", "pre"])
def test_bound_lists_equal_initialize_constraints(ref, tag):
    nlp = _nlp(ref, tag, 0)
    lo, hi = nlp.lbg_ubg_aligned()
    assert np.array_equal(lo, ref[f"{tag}_lbg"]) and np.array_equal(hi, ref[f"{tag}_ubg"])
    assert np.array_equal(nlp.zL, ref[f"{tag}_lbx"]) and np.array_equal(nlp.zU, ref[f"{tag}_ubx"])


def test_no_cbf_module_executed_from_its_bytecode(ref):
    """`MPC_optimize_kin` ships only as a CPython-3.7 .pyc; the generator ran its three methods through a
    3.7 bytecode interpreter (tests/golden/pyc37.py).  Scalar lbg/ubg, the lbx/ubx lists, the constructor
    attributes and the weights (through f) must match the restatement and the drop-in class."""
    from mpc_motion_planning_b200 import MPC_optimize_kin

    nlp = _nlp(ref, "nocbf", 0)
    assert float(ref["nocbf_lbg"]) == 0.0 and float(ref["nocbf_ubg"]) == 0.0 and nlp.n_ineq == 0
    assert np.array_equal(nlp.zL, ref["nocbf_lbx"]) and np.array_equal(nlp.zU, ref["nocbf_ubx"])
    m = MPC_optimize_kin.MPC_optimize()
    lbg, ubg, lbx, ubx = m.initialize_constraints()
    assert lbg == 0.0 and ubg == 0.0 and np.array_equal(lbx, ref["nocbf_lbx"]) and np.array_equal(ubx, ref["nocbf_ubx"])
    for k, v in json.loads(str(ref["nocbf_attrs"])).items():
        assert getattr(m, k) == v, k
    g = m._g_of(ref["nocbf_z"][0], ref["nocbf_p"][0], None)
    assert np.max(np.abs(g - ref["nocbf_g"][0])) <= 1e-11


def test_dyn_bound_lists_and_the_shipped_misalignment(ref):
    """the shipped lbg/ubg are a permutation of the aligned lists that pairs
    196 rows with the wrong bounds.  The aligned lists are what every implementation here solves."""
    nlp = _nlp(ref, "dyn", 0)
    lo, hi = nlp.lbg_ubg_aligned()
    assert np.array_equal(np.sort(lo), np.sort(ref["dyn_lbg"])) and np.array_equal(np.sort(hi), np.sort(ref["dyn_ubg"]))
    assert int((lo != ref["dyn_lbg"]).sum()) == 196 and int((hi != ref["dyn_ubg"]).sum()) == 196
    assert np.array_equal(nlp.zL, ref["dyn_lbx"]) and np.array_equal(nlp.zU, ref["dyn_ubx"])


def test_nlpsol_options_are_the_ones_the_solvers_assume(ref):
    for tag in ("kin", "pre", "dyn", "nocbf"):
        o = json.loads(str(ref[f"{tag}_opts"]))
        assert o == {"ipopt.max_iter": 100, "ipopt.print_level": 5, "print_time": 0, "ipopt.acceptable_tol": 1e-8,
                     "ipopt.acceptable_obj_change_tol": 1e-6}


def test_host_classes_reproduce_the_reference_attributes_lists_and_g(ref):
    """the drop-in `MPC_optimize` classes: constructor attributes, initialize_constraints, g, f"""
    from mpc_motion_planning_b200 import MPC_CBF_optimize_dyn, MPC_CBF_optimize_kin, MPC_CBF_optimize_kin_pre

    attrs = json.loads(str(ref["attrs"]))
    m = MPC_CBF_optimize_kin.MPC_optimize()
    for k, v in attrs.items():
        mine = getattr(m, k)
        if isinstance(v, list):
            assert np.array_equal(np.asarray(mine), np.asarray(v)), k
        else:
            assert mine == v, k
    lbg, ubg, lbx, ubx = m.initialize_constraints(ref["kin_obs"])
    assert np.array_equal(lbg, ref["kin_lbg"]) and np.array_equal(ubg, ref["kin_ubg"])
    assert np.array_equal(lbx, ref["kin_lbx"]) and np.array_equal(ubx, ref["kin_ubx"])
    g = m._g_of(ref["kin_z"][0], ref["kin_p"][0], m._obs_array(ref["kin_obs"]))
    assert np.max(np.abs(g - ref["kin_g"][0])) <= 1e-11

    mp = MPC_CBF_optimize_kin_pre.MPC_optimize()
    obs_list = [o[None, :] for o in ref["pre_obs0"]]
    lbg, ubg, lbx, ubx = mp.initialize_constraints(obs_list)
    assert np.array_equal(lbg, ref["pre_lbg"]) and np.array_equal(ubg, ref["pre_ubg"])
    g = mp._g_of(ref["pre_z"][1], ref["pre_p"][1], mp._obs_array(list(ref["pre_obs"])))
    assert np.max(np.abs(g - ref["pre_g"][1])) <= 1e-11

    md = MPC_CBF_optimize_dyn.MPC_optimize()
    lbg, ubg, lbx, ubx = md.initialize_constraints()
    assert np.array_equal(lbx, ref["dyn_lbx"]) and np.array_equal(ubx, ref["dyn_ubx"])
    assert np.array_equal(lbg, ref["dyn_lbg"]) and np.array_equal(ubg, ref["dyn_ubg"])  # exactly as shipped (default)
    md.dyn_bounds = "aligned"
    assert np.array_equal(np.sort(md.initialize_constraints()[0]), np.sort(ref["dyn_lbg"]))
    g = md._g_of(ref["dyn_z"][2], ref["dyn_p"][2], md._obs_array(ref["dyn_obs"]))
    assert np.max(np.abs(g - ref["dyn_g"][2])) <= 1e-11
    q = ref["dyn_rhs_in"]
    assert np.max(np.abs(md.f(q[:6], q[6:]).full().ravel() - ref["dyn_rhs_out"])) <= 1e-12


def test_generate_ref_path_equals_the_reference(ref):
    from mpc_motion_planning_b200 import MPC_CBF_optimize_kin

    m = MPC_CBF_optimize_kin.MPC_optimize()
    for q, want in zip(ref["genref_in"], ref["genref_out"]):
        got = m.generate_ref_path(q[:4].reshape(-1, 1), q[4:].reshape(-1, 1))
        assert got.shape == want.shape == (51, 4)
        assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) <= 1e-10


def test_ref_path_generator_and_obs_prediction_equal_the_reference(ref):
    from mpc_motion_planning_b200 import RefPathGenerator
    from mpc_motion_planning_b200.Obs_prediction import obs_prediction
    from oracle.nlp import predict_obstacles

    rp = RefPathGenerator.RefPathGenerator()
    x0, xs = np.array([0, 3, 0, 15.0]).reshape(-1, 1), np.array([400, 3.5, 0, 30.0]).reshape(-1, 1)
    assert np.array_equal(rp.define_ref_path(x0, xs, 0.1), ref["refpath_global"])
    tr, idx = rp.find_ref_traj(np.array([37.3, 2.9, 0.01, 17.5]).reshape(-1, 1), xs, 5, 0.1, 30)
    assert np.array_equal(tr, ref["refpath_traj"]) and idx == int(ref["refpath_idx"])
    obs_list = [o[None, :] for o in ref["pre_obs0"]]
    mine = obs_prediction(obs_list, 0.1, 50)
    assert np.array_equal(np.array(mine), ref["pre_obs"])
    assert np.array_equal(np.array(predict_obstacles(obs_list, 0.1, 50)), ref["pre_obs"])


def test_c_oracle_agrees_with_the_reference_objective_at_its_solution(ref):
    """the C oracle's reported cost is the reference's f at the returned z (same restated NLP)"""
    from oracle import c_oracle

    p = ref["pre_p"][0]
    obs = ref["pre_obs"][None, :1]  # one obstacle
    cfg = c_oracle.make_cfg("kin_cbf_pre")
    z, lam, info = c_oracle.solve(cfg, p[:4], p[4:], obs[0])
    nlp = NLP("kin_cbf_pre", p[:4], p[4:], [ref["pre_obs"][0]])
    assert info.status == 0 and abs(nlp.objective(z) - info.f) <= 1e-12 * abs(info.f)


KKT_CASES = {"kin": ("kin_cbf", [0, 3, 0, 15.0], [400, 3.5, 0, 30.0]), "pre": ("kin_cbf_pre", [0, 3, 0, 15.0], [400, 3.5, 0, 30.0]),
             "dyn": ("dyn", [0, 0, 0, 10, 0, 0.0], [600, 3.5, 0, 15, 0, 0.0]), "nocbf": ("kin_nocbf", [0, 0, 0, 20.0], [500, 3.5, 0, 30.0])}


def kkt_case_obs(ref, tag, N=50):
    """obstacle array (M,N+1,6) of the mains' default scenarios used for the KKT vectors"""
    if tag == "kin":
        return np.repeat(np.array([[50, 3.5, 0, 8, 4.8, 1.8]])[:, None, :], N + 1, axis=1)
    if tag == "pre":
        return ref["pre_obs"][:1]
    if tag == "dyn":
        o = np.zeros((1, N + 1, 6))
        o[0, :, 0], o[0, :, 1] = ref["dyn_obs"]
        return o
    return None


@pytest.mark.parametrize("tag", ["kin", "pre", "dyn", "nocbf"])
def test_oracle_solutions_are_kkt_points_of_the_reference_nlp(ref, tag):
    """At generation time the dense interior-point solution of each main's first step was checked against the
    REFERENCE'S expressions with exact (sympy) derivatives: stationarity below IPOPT's tol in IPOPT's scaling,
    rows and bounds feasible up to bound_relax_factor, complementarity below tol.  Here: those recorded
    residuals, and the C oracle reaching the same point."""
    from oracle import c_oracle

    k = lambda n: float(ref[f"{tag}_kkt_{n}"])
    assert k("stationarity_scaled") <= 1e-8 and k("complementarity_scaled") <= 1e-8
    assert k("g_violation") <= 1.01e-8 and k("x_violation") <= 4.01e-7  # 1e-8 * max(1, |bound|), |vx_max| = 40
    assert abs(k("f_ref") - k("f_oracle")) <= 1e-14 * abs(k("f_ref"))
    kind, x0, xs = KKT_CASES[tag]
    z, _, info = c_oracle.solve(c_oracle.make_cfg(kind), np.array(x0), np.array(xs), kkt_case_obs(ref, tag))
    assert info.status == 0
    assert np.max(np.abs(z - ref[f"{tag}_kkt_z"])) <= 1e-7 and abs(info.f - k("f_ref")) <= 1e-10 * k("f_ref")


KKT64_XS = {"kin": [400, 3.5, 0, 30.0], "pre": [400, 3.5, 0, 30.0], "dyn": [600, 3.5, 0, 15, 0, 0.0], "nocbf": [500, 3.5, 0, 30.0]}


@pytest.mark.parametrize("tag", ["kin", "pre", "dyn", "nocbf"])
def test_64_start_states_per_nlp_end_at_kkt_points_of_the_reference_nlp(ref, tag):
    """64 seeded start states per module (the parameter vector is symbolic in the recorded NLP, so one set of exact sympy
    derivatives of the REFERENCE'S expressions serves them all): recorded residuals of the oracle's points, and the C
    oracle reaching the same points."""
    from oracle import c_oracle

    k = lambda n: ref[f"{tag}_kkt64_{n}"]
    assert k("x0").shape[0] == 64
    assert k("stationarity_scaled").max() <= 2e-8 and k("complementarity_scaled").max() <= 2e-8  # IPOPT's tol 1e-8 on its scaled error
    assert k("g_violation").max() <= 1.01e-8
    kind = KKT_CASES[tag][0]
    cfg = c_oracle.make_cfg(kind)
    obs = kkt_case_obs(ref, tag)
    B = 64
    xs = np.tile(np.array(KKT64_XS[tag], float), (B, 1))
    ob = None if obs is None else np.tile(obs[None], (B, 1, 1, 1))
    u0, cost, st, it, z = c_oracle.solve_batch(cfg, k("x0"), xs, ob, want_z=True, nthreads=os.cpu_count())
    ok = st == 0
    assert ok.sum() >= 62  # the dense specification converged on all 64; the Riccati form may leave the chaotic one or two
    assert np.abs(z[ok] - k("z")[ok]).max() <= 1e-6
    assert (np.abs(cost[ok] - k("f_ref")[ok]) / np.abs(k("f_ref")[ok])).max() <= 1e-9


def test_dyn_problem_as_shipped_at_oracle_level(ref):
    """The dyn bound lists exactly as shipped define a different problem:
    every rate row becomes an equality (one control pair for the whole horizon) and the x/y defects of
    stages 2..N are relaxed.  `ShippedDynNLP` reproduces the shipped lists bit for bit and the dense
    interior point solves it; the CUDA path solves the aligned problem."""
    from oracle import ipm_dense
    from oracle.nlp import ShippedDynNLP

    nlp = ShippedDynNLP([0, 0, 0, 10, 0, 0.0], [600, 3.5, 0, 15, 0, 0.0], ref["dyn_obs"], N=20)
    full = ShippedDynNLP([0, 0, 0, 10, 0, 0.0], [600, 3.5, 0, 15, 0, 0.0], ref["dyn_obs"])
    assert np.array_equal(full.lbg_shipped, ref["dyn_lbg"]) and np.array_equal(full.ubg_shipped, ref["dyn_ubg"])
    assert full.n_eq == 306 and full.n_ineq == 98 + 51  # 49 rate pairs -> equalities, 49 x/y defect pairs -> ranges, 51 obstacle rows
    r = ipm_dense.solve(nlp, nlp.rollout_start(), ipm_dense.IpmOptions())
    U, X = nlp.split(r.z)
    assert r.status == 0 and np.abs(U - U[0]).max() <= 1e-10
    g = nlp.g_ref(r.z)
    assert np.all(g >= nlp.lbg_shipped - 1e-7) and np.all(g <= nlp.ubg_shipped + 1e-7)
