"""Golden vectors produced by the REFERENCE'S OWN CODE, run in the build container.

    python tests/golden/make_reference_vectors.py        (needs /root/reference; writes reference_nlp.npz)

CasADi/IPOPT cannot be installed here, so the reference's *solve* cannot run.  Its NLP *definition*
can: with tests/golden/casadi_stub.py registered as `casadi`, the unmodified
`MPC_optimize.__init__`, `initialize_constraints`, `optimize_problem` and `generate_ref_path` of
PKG/MPC_CBF_optimize_kin.py, _kin_pre.py and _dyn.py execute and hand their objective / constraint
expressions to `nlpsol`, where the stub records them.  This script evaluates those expressions at
seeded points and stores

    z, p, f(z,p), g(z,p), lbg, ubg, lbx, ubx, the nlpsol options, N_p, T_S

per module, plus outputs of `generate_ref_path`, `RefPathGenerator` and `obs_prediction`.
tests/test_reference_vectors.py checks the restated NLP (oracle/nlp.py), the drop-in host classes
and the bound lists against them.  `MPC_optimize_kin`, the no-CBF module, exists only as a
CPython-3.7 .pyc that this interpreter cannot import: its three methods are executed from the
bytecode by the small 3.7 interpreter in tests/golden/pyc37.py.
"""
from __future__ import annotations

import importlib
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = "/root/reference/CasaDi_MPC_Optimize_Multishoot"
K = 4  # evaluation points per module


def load_module(name, workdir):
    import casadi_stub

    sys.modules["casadi"] = casadi_stub
    casadi_stub.tools = sys.modules["casadi.tools"] = types.ModuleType("casadi.tools")  # imported, never used
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    os.chdir(workdir)  # the constructors read mpc_parameters.yaml from the CWD
    return importlib.import_module(name)


def evaluate(rec, pts_z, pts_p):
    """numeric f and g of the recorded NLP at the given points"""
    xs = [s for s in rec.prob["x"].a.reshape(-1, order="F")]
    ps = [s for s in rec.prob["p"].a.reshape(-1, order="F")]
    f = rec.prob["f"].a.reshape(-1)[0]
    g = list(rec.prob["g"].a.reshape(-1, order="F"))
    fun = sp.lambdify([xs, ps], [f] + g, "math", cse=True)
    F, G = [], []
    for z, p in zip(pts_z, pts_p):
        out = fun(list(z), list(p))
        F.append(float(out[0]))
        G.append([float(v) for v in out[1:]])
    return np.array(F), np.array(G)


def kkt_at_oracle_solution(rec, kind, x0, xs, obstacles, lbg, ubg, lbx, ubx):
    """Solve the restated NLP with the dense interior-point specification, then measure the KKT residuals of that
    point on the REFERENCE'S expressions (exact sympy derivatives of what the reference handed to nlpsol):
    stationarity  grad f + J_g' lam_g - z_L + z_U,  feasibility of g and x against the reference's own bound lists,
    complementarity.  Returns z*, f*, and the residual norms (objective-scaled like IPOPT's error)."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import ipm_dense
    from oracle.nlp import NLP

    nlp = NLP(kind, x0, xs, obstacles)
    r = ipm_dense.solve(nlp, nlp.rollout_start(), ipm_dense.IpmOptions())
    assert r.status == 0, (kind, r.status)
    xsym = list(rec.prob["x"].a.reshape(-1, order="F"))
    psym = list(rec.prob["p"].a.reshape(-1, order="F"))
    pos = {v: i for i, v in enumerate(xsym)}
    f = rec.prob["f"].a.reshape(-1)[0]
    g = list(rec.prob["g"].a.reshape(-1, order="F"))
    # sparse exact derivatives: every row depends on a handful of variables
    entries, exprs = [], [f] + g
    for row, e in enumerate(exprs):
        for v in sorted(sp.sympify(e).free_symbols & set(xsym), key=lambda q: pos[q]):
            entries.append((row, pos[v], sp.diff(e, v)))
    fun = sp.lambdify([xsym, psym], [e for _, _, e in entries] + exprs, "math", cse=True)
    pvec = np.concatenate([np.asarray(x0, float), np.asarray(xs, float)])
    vals = fun(list(r.z), list(pvec))
    nv, ng = len(xsym), len(g)
    J = np.zeros((1 + ng, nv))
    for (row, col, _), v in zip(entries, vals[:len(entries)]):
        J[row, col] = float(v)
    fval, gval = float(vals[len(entries)]), np.array([float(v) for v in vals[len(entries) + 1:]])
    lam_g = np.concatenate([r.lam_eq, r.lam_in])[nlp.g_perm()]
    stat = J[0] + J[1:].T @ lam_g - r.zl + r.zu
    lo, hi = (np.full(ng, float(lbg)), np.full(ng, float(ubg))) if np.isscalar(lbg) else nlp.lbg_ubg_aligned()
    if kind != "dyn" and not np.isscalar(lbg):  # the shipped lists are the aligned ones for the kinematic modules
        assert np.array_equal(lo, np.array(lbg, float)) and np.array_equal(hi, np.array(ubg, float))
    viol_g = float(max(np.max(lo - gval), np.max(gval - hi), 0.0))
    viol_x = float(max(np.max(np.array(lbx, float) - r.z), np.max(r.z - np.array(ubx, float)), 0.0))
    act = np.minimum(gval - lo, hi - gval)
    compl = float(np.max(np.abs(lam_g) * np.where(np.isfinite(act), act, 0.0) * (lo != hi)))
    return {"z": r.z, "f_ref": fval, "f_oracle": r.f, "stationarity_scaled": float(r.obj_scale * np.max(np.abs(stat))),
            "g_violation": viol_g, "x_violation": viol_x, "complementarity_scaled": float(r.obj_scale * compl), "iters": r.iters}


def _dense_job(args):
    """one dense interior-point solve of the restated NLP (worker of kkt_batch)"""
    kind, x0, xs, obstacles = args
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import ipm_dense
    from oracle.nlp import NLP

    from threadpoolctl import threadpool_limits

    nlp = NLP(kind, x0, xs, obstacles)
    with threadpool_limits(1):  # one BLAS thread per worker: the pool already uses every core
        r = ipm_dense.solve(nlp, nlp.rollout_start(), ipm_dense.IpmOptions())
    return r.status, r.z, r.f, np.concatenate([r.lam_eq, r.lam_in])[nlp.g_perm()], r.zl, r.zu, r.obj_scale


def kkt_batch(rec, kind, x0_draws, xs, obstacles, lbg, ubg, want=64):
    """`kkt_at_oracle_solution` for many start states: the parameter vector p = [x0; xs] is symbolic in the recorded NLP,
    so one set of exact derivatives serves every x0.  The first `want` draws whose dense solve converges are kept; for
    each, the KKT residuals of the oracle's point ON THE REFERENCE'S EXPRESSIONS are recorded next to the point."""
    from concurrent.futures import ProcessPoolExecutor
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.nlp import NLP

    xsym = list(rec.prob["x"].a.reshape(-1, order="F"))
    psym = list(rec.prob["p"].a.reshape(-1, order="F"))
    pos = {v: i for i, v in enumerate(xsym)}
    f = rec.prob["f"].a.reshape(-1)[0]
    g = list(rec.prob["g"].a.reshape(-1, order="F"))
    entries, exprs = [], [f] + g
    for row, e in enumerate(exprs):
        for v in sorted(sp.sympify(e).free_symbols & set(xsym), key=lambda q: pos[q]):
            entries.append((row, pos[v], sp.diff(e, v)))
    fun = sp.lambdify([xsym, psym], [e for _, _, e in entries] + exprs, "math", cse=True)
    nv, ng = len(xsym), len(g)
    with ProcessPoolExecutor(os.cpu_count()) as ex:
        sols = list(ex.map(_dense_job, [(kind, x0, xs, obstacles) for x0 in x0_draws]))
    nlp0 = NLP(kind, x0_draws[0], xs, obstacles)
    lo, hi = (np.full(ng, float(lbg)), np.full(ng, float(ubg))) if np.isscalar(lbg) else nlp0.lbg_ubg_aligned()
    keep = {k: [] for k in ("x0", "z", "f_ref", "stationarity_scaled", "g_violation", "complementarity_scaled")}
    for x0, (status, z, fo, lam_g, zl, zu, osc) in zip(x0_draws, sols):
        if status != 0 or len(keep["x0"]) >= want:
            continue
        pvec = np.concatenate([np.asarray(x0, float), np.asarray(xs, float)])
        vals = fun(list(z), list(pvec))
        J = np.zeros((1 + ng, nv))
        for (row, col, _), v in zip(entries, vals[:len(entries)]):
            J[row, col] = float(v)
        fval, gval = float(vals[len(entries)]), np.array([float(v) for v in vals[len(entries) + 1:]])
        stat = J[0] + J[1:].T @ lam_g - zl + zu
        act = np.minimum(gval - lo, hi - gval)
        keep["x0"].append(np.asarray(x0, float)); keep["z"].append(z); keep["f_ref"].append(fval)
        keep["stationarity_scaled"].append(float(osc * np.max(np.abs(stat))))
        keep["g_violation"].append(float(max(np.max(lo - gval), np.max(gval - hi), 0.0)))
        keep["complementarity_scaled"].append(float(osc * np.max(np.abs(lam_g) * np.where(np.isfinite(act), act, 0.0) * (lo != hi))))
        assert abs(fval - fo) <= 1e-12 * abs(fval)
    assert len(keep["x0"]) >= want, (kind, len(keep["x0"]))
    return {k: np.array(v) for k, v in keep.items()}


def rollout_points(mod, nx, N, x0, xs, rng):
    """z near an Euler roll-out (keeps the dyn sqrt rows real), p = [x0; xs] perturbed"""
    Z, P = [], []
    for _ in range(K):
        x = np.array(x0, float) + rng.normal(0, [1.0, 0.2, 0.01, 1.0, 0.05, 0.01][:nx])
        U = np.c_[rng.uniform(-0.02, 0.02, N), rng.uniform(-1, 1, N)]
        X = [x]
        for i in range(N):
            X.append(X[-1] + mod.T_S * mod.f(X[-1], U[i]).full().ravel())
        X = np.array(X) + rng.normal(0, 0.01, (N + 1, nx))
        Z.append(np.concatenate([U.reshape(-1), X.reshape(-1)]))
        P.append(np.concatenate([x, np.array(xs, float) + rng.normal(0, 0.1, nx)]))
    return np.array(Z), np.array(P)


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    sys.path.insert(0, PKG)
    from Obs_prediction import obs_prediction
    import RefPathGenerator

    out = {}
    rng = np.random.default_rng(20261018)

    # ---------------- kin-CBF static (PKG/main_cbf_kin_c_sim.py:45-55,77,99)
    m = load_module("MPC_CBF_optimize_kin", PKG)
    mpc = m.MPC_optimize()
    N = mpc.N_p
    obs = np.array([[50, 3.5, 0, 8, 4.8, 1.8], [90, 0.5, 0, 0, 4.2, 1.7]])
    x0, xs = [0, 3, 0, 15], [400, 3.5, 0, 30]
    lbg, ubg, lbx, ubx = mpc.initialize_constraints(obs)
    ref = np.tile(np.array(xs, float), (N + 1, 1))
    rec = mpc.optimize_problem(np.array(x0).reshape(-1, 1), ref, obs)
    Z, P = rollout_points(mpc, 4, N, x0, xs, rng)
    F, G = evaluate(rec, Z, P)
    out.update(kin_z=Z, kin_p=P, kin_f=F, kin_g=G, kin_lbg=np.array(lbg, float), kin_ubg=np.array(ubg, float),
               kin_lbx=np.array(lbx, float), kin_ubx=np.array(ubx, float), kin_obs=obs, kin_opts=json.dumps(rec.opts),
               kin_N=N, kin_T=mpc.T_S)
    one = np.array([[50, 3.5, 0, 8, 4.8, 1.8]])  # the main's own obstacle (PKG/main_cbf_kin_c_sim.py:55)
    rec1 = mpc.optimize_problem(np.array(x0).reshape(-1, 1), ref, one)
    l1 = mpc.initialize_constraints(one)
    out.update({f"kin_kkt_{k}": v for k, v in kkt_at_oracle_solution(rec1, "kin_cbf", x0, xs, one, *l1).items()})
    # 64 more start states per NLP (own generator: the draws above stay what they were)
    rng64 = np.random.default_rng(20261019)
    kin_draws = [np.array([rng64.uniform(0, 20), rng64.uniform(0, 4.5), rng64.uniform(-0.05, 0.05), rng64.uniform(10, 25)]) for _ in range(160)]
    out.update({f"kin_kkt64_{k}": v for k, v in kkt_batch(rec1, "kin_cbf", kin_draws, xs, one, l1[0], l1[1]).items()})
    # the quintic lane-change reference (:258-308), unused by the mains
    gr_in, gr_out = [], []
    for a, b in [([0, 3, 0, 15], [400, 3.5, 0, 30]), ([12.5, 0.2, 0.01, 22], [400, 3.5, 0, 25]), ([100, 4.0, 0, 8], [600, 0.0, 0, 12])]:
        a, b = np.array(a, float).reshape(-1, 1), np.array(b, float).reshape(-1, 1)
        gr_in.append(np.concatenate([a.ravel(), b.ravel()]))
        gr_out.append(mpc.generate_ref_path(a, b))
    out.update(genref_in=np.array(gr_in), genref_out=np.array(gr_out))
    out.update(attrs=json.dumps({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in vars(mpc).items()
                                 if isinstance(v, (int, float, str, bool, np.ndarray)) and k != "config"}))

    # ---------------- kin-CBF moving (PKG/main_cbf_kin_c_sim_pre.py:45-56,77,99)
    m = load_module("MPC_CBF_optimize_kin_pre", PKG)
    mpc = m.MPC_optimize()
    obs_list = [np.array([[50, 3.5, 0, 10, 4.8, 1.8]]), np.array([[80, 0.0, 0.03, 6, 3.9, 1.6]])]
    tr = obs_prediction(obs_list, mpc.T_S, N)
    lbg, ubg, lbx, ubx = mpc.initialize_constraints(obs_list)
    rec = mpc.optimize_problem(np.array(x0).reshape(-1, 1), ref, tr)
    Z, P = rollout_points(mpc, 4, N, x0, xs, rng)
    F, G = evaluate(rec, Z, P)
    out.update(pre_z=Z, pre_p=P, pre_f=F, pre_g=G, pre_lbg=np.array(lbg, float), pre_ubg=np.array(ubg, float),
               pre_lbx=np.array(lbx, float), pre_ubx=np.array(ubx, float), pre_obs=np.array(tr), pre_obs0=np.array(obs_list).reshape(-1, 6),
               pre_opts=json.dumps(rec.opts))

    tr1 = obs_prediction(obs_list[:1], mpc.T_S, N)
    rec1 = mpc.optimize_problem(np.array(x0).reshape(-1, 1), ref, tr1)
    l1 = mpc.initialize_constraints(obs_list[:1])
    out.update({f"pre_kkt_{k}": v for k, v in kkt_at_oracle_solution(rec1, "kin_cbf_pre", x0, xs, tr1, *l1).items()})
    out.update({f"pre_kkt64_{k}": v for k, v in kkt_batch(rec1, "kin_cbf_pre", kin_draws, xs, tr1, l1[0], l1[1]).items()})

    # ---------------- dyn (PKG/main_cbf_dyn_c_sim.py:44-51,77,89): the module reads `Veh_w`, which the
    # shipped YAML spells `Veh_W`: run it from a copy with the key added
    tmp = tempfile.mkdtemp()
    with open(os.path.join(PKG, "mpc_parameters.yaml")) as fh:
        text = fh.read()
    with open(os.path.join(tmp, "mpc_parameters.yaml"), "w") as fh:
        fh.write(text.replace("  Veh_W: 1.8", "  Veh_W: 1.8\n  Veh_w: 1.8"))
    m = load_module("MPC_CBF_optimize_dyn", tmp)
    mpc = m.MPC_optimize()
    x0d, xsd, obsd = [0, 0, 0, 10, 0, 0], [600, 3.5, 0, 15, 0, 0], np.array([100, -3.5])
    lbg, ubg, lbx, ubx = mpc.initialize_constraints()
    rec = mpc.optimize_problem(np.array(x0d).reshape(-1, 1), np.tile(np.array(xsd, float), (N + 1, 1)), obsd)
    Z, P = rollout_points(mpc, 6, N, x0d, xsd, rng)
    F, G = evaluate(rec, Z, P)
    out.update(dyn_z=Z, dyn_p=P, dyn_f=F, dyn_g=G, dyn_lbg=np.array(lbg, float), dyn_ubg=np.array(ubg, float),
               dyn_lbx=np.array(lbx, float), dyn_ubx=np.array(ubx, float), dyn_obs=obsd, dyn_opts=json.dumps(rec.opts))
    out.update({f"dyn_kkt_{k}": v for k, v in kkt_at_oracle_solution(rec, "dyn", x0d, xsd, obsd, lbg, ubg, lbx, ubx).items()})
    dyn_draws = [np.array([rng64.uniform(0, 20), rng64.uniform(-0.5, 4.5), rng64.uniform(-0.05, 0.05), rng64.uniform(8, 20), 0.0, 0.0]) for _ in range(110)]
    out.update({f"dyn_kkt64_{k}": v for k, v in kkt_batch(rec, "dyn", dyn_draws, xsd, obsd, lbg, ubg).items()})
    xq, uq = np.array([1.0, 0.5, 0.02, 12.0, 0.3, 0.05]), np.array([0.03, 1.2])
    out.update(dyn_rhs_in=np.concatenate([xq, uq]), dyn_rhs_out=mpc.f(xq, uq).full().ravel())
    shutil.rmtree(tmp)
    os.chdir(HERE)

    # ---------------- no-CBF kin module: only PKG/__pycache__/MPC_optimize_kin.cpython-37.pyc exists.  Its three
    # methods are executed from the bytecode (tests/golden/pyc37.py) on the same casadi stand-in
    # (PKG/main_kin_c_sim.py:42-46,68,83 for the call protocol).  Its __init__ reads `Veh_w` like the dyn module.
    import math
    import yaml
    import pyc37
    import casadi_stub
    from helpers import load_config

    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "mpc_parameters.yaml"), "w") as fh:
        fh.write(text.replace("  Veh_W: 1.8", "  Veh_W: 1.8\n  Veh_w: 1.8"))
    os.chdir(tmp)
    mod = pyc37.load_pyc(os.path.join(PKG, "__pycache__", "MPC_optimize_kin.cpython-37.pyc"))
    glb = {"ca": casadi_stub, "np": np, "math": math, "yaml": yaml, "load_config": load_config, "PARAMS_FILE": "mpc_parameters.yaml"}
    me = types.SimpleNamespace()
    pyc37.run(pyc37.find_code(mod, "MPC_optimize", "__init__"), glb, me)
    lbg, ubg, lbx, ubx = pyc37.run(pyc37.find_code(mod, "MPC_optimize", "initialize_constraints"), glb, me)
    x0n, xsn = [0, 0, 0, 20], [500, 3.5, 0, 30]
    rec = pyc37.run(pyc37.find_code(mod, "MPC_optimize", "optimize_problem"), glb, me, np.array(x0n).reshape(-1, 1),
                    np.tile(np.array(xsn, float), (N + 1, 1)))
    Z, P = rollout_points(me, 4, N, x0n, xsn, rng)
    F, G = evaluate(rec, Z, P)
    out.update(nocbf_z=Z, nocbf_p=P, nocbf_f=F, nocbf_g=G, nocbf_lbg=np.array(lbg, float), nocbf_ubg=np.array(ubg, float),
               nocbf_lbx=np.array(lbx, float), nocbf_ubx=np.array(ubx, float), nocbf_opts=json.dumps(rec.opts),
               nocbf_attrs=json.dumps({k: v for k, v in vars(me).items() if isinstance(v, (int, float, str, bool))}))
    out.update({f"nocbf_kkt_{k}": v for k, v in kkt_at_oracle_solution(rec, "kin_nocbf", x0n, xsn, None, lbg, ubg, lbx, ubx).items()})
    out.update({f"nocbf_kkt64_{k}": v for k, v in kkt_batch(rec, "kin_nocbf", kin_draws[:64], xsn, None, lbg, ubg).items()})
    shutil.rmtree(tmp)
    os.chdir(HERE)

    # ---------------- RefPathGenerator / obs_prediction (numpy only, imported as they are)
    rp = RefPathGenerator.RefPathGenerator()
    x0c, xsc = np.array(x0, float).reshape(-1, 1), np.array(xs, float).reshape(-1, 1)
    glob = rp.define_ref_path(x0c, xsc, 0.1)
    rt, idx = rp.find_ref_traj(np.array([37.3, 2.9, 0.01, 17.5]).reshape(-1, 1), xsc, 5, 0.1, 30)
    out.update(refpath_global=glob, refpath_traj=rt, refpath_idx=idx)

    np.savez_compressed(os.path.join(HERE, "reference_nlp.npz"), **out)
    print("wrote reference_nlp.npz:", {k: np.asarray(v).shape for k, v in out.items()})
