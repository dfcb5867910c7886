"""Generate tests/golden/solutions.npz.

No reference golden vectors exist (the reference has no tests and CasADi/IPOPT are not
installable here), so these fixtures are produced by TWO independent CPU solvers on the restated
NLPs (oracle/nlp.py):
  (a) the dense interior-point specification oracle/ipm_dense.py (scipy LDL^T, no Riccati);
  (b) scipy.optimize SLSQP (an active-set SQP, different algorithm and code base), started
      (b1) from the same roll-out point and (b2) from (a)'s solution.
A scenario is stored with `agree=1` when (b1) reaches the same optimum as (a) (cost 1e-6 rel, u0
1e-4), and `local_min=1` when (b2) stays at (a)'s solution (it is a KKT point / local minimiser).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys
import time

import numpy as np
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpc_motion_planning_b200 import scenarios  # noqa: E402
from oracle import ipm_dense  # noqa: E402
from oracle.nlp import NLP, default_scenario  # noqa: E402


def slsqp(nlp, z0):
    cons = [{"type": "eq", "fun": nlp.eq, "jac": nlp.jac_eq}]
    if nlp.n_ineq:
        fin = np.isfinite(nlp.dU)

        def f_in(z):
            d = nlp.ineq(z)
            return np.concatenate([d - nlp.dL, (nlp.dU - d)[fin]])

        def j_in(z):
            J = nlp.jac_ineq(z)
            return np.vstack([J, -J[fin]])

        cons.append({"type": "ineq", "fun": f_in, "jac": j_in})
    sc = 1e-4
    with np.errstate(all="ignore"):
        r = minimize(lambda z: sc * nlp.objective(z), z0, jac=lambda z: sc * nlp.grad(z), method="SLSQP",
                     bounds=list(zip(nlp.zL, nlp.zU)), constraints=cons, options={"maxiter": 600, "ftol": 1e-14})
    return r.x, r.fun / sc


def obs_of(nlp, N):
    if nlp.kind == "kin_nocbf":
        return np.zeros((0, N + 1, 6))
    ob = np.zeros((nlp.M, N + 1, 6))
    ob[:, :, 0:2] = nlp.oc
    if nlp.kind != "dyn":
        ob[:, :, 4] = 2 * (nlp.osx - nlp.p.Veh_L / 2 - 1.0)
        ob[:, :, 5] = 2 * (nlp.osy - nlp.p.Veh_W / 2 - 0.5)
    return ob


def main():
    cases = []
    for kind in ("kin_nocbf", "kin_cbf", "kin_cbf_pre", "dyn"):
        cases.append((kind + "_default", default_scenario(kind)))
    x0, xs, obs = scenarios.kin_cbf_static(6)
    for i in range(6):
        cases.append((f"kin_cbf_rand{i}", NLP("kin_cbf_pre", x0[i], xs[i], [obs[i, 0]])))
    x0, xs, obs = scenarios.kin_cbf_moving(6)
    for i in range(6):
        cases.append((f"kin_cbf_pre_rand{i}", NLP("kin_cbf_pre", x0[i], xs[i], [obs[i, 0]])))
    x0, xs, obs = scenarios.kin_nocbf(3)
    for i in range(3):
        cases.append((f"kin_nocbf_rand{i}", NLP("kin_nocbf", x0[i], xs[i], None)))
    x0, xs, obs = scenarios.dyn_static(4)
    for i in range(4):
        cases.append((f"dyn_rand{i}", NLP("dyn", x0[i], xs[i], obs[i, 0, 0, 0:2])))
    out = {}
    names = []
    for name, nlp in cases:
        t = time.time()
        z0 = nlp.rollout_start()
        r = ipm_dense.solve(nlp, z0, ipm_dense.IpmOptions())
        agree = local_min = 0
        f1 = f2 = np.nan
        if r.status == 0:
            z1, f1 = slsqp(nlp, z0)
            agree = int(abs(f1 - r.f) <= 1e-6 * abs(r.f) and np.abs(z1[:2] - r.z[:2]).max() <= 1e-4)
            z2, f2 = slsqp(nlp, r.z)
            local_min = int(abs(f2 - r.f) <= 1e-6 * abs(r.f) and np.abs(z2[:2] - r.z[:2]).max() <= 1e-4)
        kind = nlp.kind if nlp.kind != "kin_cbf" else "kin_cbf"
        names.append(name)
        out[name + "/kind"] = np.array(kind)
        out[name + "/x0"] = nlp.x0
        out[name + "/xs"] = nlp.xs
        out[name + "/obs"] = obs_of(nlp, nlp.N)
        out[name + "/z"] = r.z
        out[name + "/f"] = np.array(r.f)
        out[name + "/status"] = np.array(r.status)
        out[name + "/iters"] = np.array(r.iters)
        out[name + "/lam_eq"] = r.lam_eq
        out[name + "/agree"] = np.array(agree)
        out[name + "/local_min"] = np.array(local_min)
        out[name + "/f_slsqp_rollout"] = np.array(f1)
        out[name + "/f_slsqp_from_ipm"] = np.array(f2)
        print(f"{name:22s} st {r.status} it {r.iters:3d} f {r.f:.10e} u0 {r.z[:2]} slsqp(rollout) {f1:.10e} agree {agree} "
              f"slsqp(from ipm) {f2:.10e} local_min {local_min}  {time.time() - t:.1f}s", flush=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "solutions.npz"), **out)


if __name__ == "__main__":
    main()
