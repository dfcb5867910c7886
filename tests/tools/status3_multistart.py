"""Which of the scenarios that the interior-point solve gives up on (status 3, restoration off) are locally infeasible
for the start point the benchmark uses, and which are solver weakness?  An independent solver - scipy SLSQP, an active-set
SQP with its own globalisation - is run on the restated NLP (oracle/nlp.py) of a seeded sample of the status-3 set of
BASELINE configs[1] from three start points:
   (a) the benchmark's own start (zero controls, Euler roll-out),
   (b) / (c) a steering pulse to the left / right that puts the roll-out on either side of the obstacle.
Prints one line per scenario and a summary; the output is committed as profiles/r02_status3_multistart.txt.

  python tests/tools/status3_multistart.py [sample=150] [workers=6]
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpc_motion_planning_b200 import scenarios  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle.nlp import NLP  # noqa: E402

B = 10000
X0, XS, OBS = scenarios.kin_cbf_static(B, seed=scenarios.BASE_SEED + 2)


def slsqp(nlp, z0, maxiter=400):
    fin = np.isfinite(nlp.dU)
    cons = [{"type": "eq", "fun": nlp.eq, "jac": nlp.jac_eq},
            {"type": "ineq", "fun": lambda z: np.concatenate([nlp.ineq(z) - nlp.dL, (nlp.dU - nlp.ineq(z))[fin]]),
             "jac": lambda z: np.vstack([nlp.jac_ineq(z), -nlp.jac_ineq(z)[fin]])}]
    sc = 1e-4
    with np.errstate(all="ignore"):
        r = minimize(lambda z: sc * nlp.objective(z), z0, jac=lambda z: sc * nlp.grad(z), method="SLSQP",
                     bounds=list(zip(np.where(np.isfinite(nlp.zL), nlp.zL, None), np.where(np.isfinite(nlp.zU), nlp.zU, None))),
                     constraints=cons, options={"maxiter": maxiter, "ftol": 1e-12})
    z = r.x
    viol = max(np.abs(nlp.eq(z)).max(), np.maximum(nlp.dL - nlp.ineq(z), 0).max(), np.maximum(nlp.ineq(z) - nlp.dU, 0).max(),
               np.maximum(nlp.zL - z, 0).max(), np.maximum(z - nlp.zU, 0).max())
    return bool(r.success and viol <= 1e-6), float(nlp.objective(z)), float(viol), int(r.nit)


def starts(nlp):
    N = nlp.N
    out = [("rollout", nlp.rollout_start())]
    for name, sgn in (("left", 1.0), ("right", -1.0)):
        U = np.zeros((N, 2))
        k = np.arange(N)
        # triangular steering pulse, 0 -> 0.064 rad -> 0 over 16 steps (0.008 rad per step: inside the rate limit of 8.7e-3)
        U[:, 0] = sgn * 0.008 * np.clip(np.minimum(k, 16 - k), 0, None)
        out.append((name, nlp.rollout_start(U)))
    return out


def work(b):
    nlp = NLP("kin_cbf", X0[b], XS[b], OBS[b, :, 0, :])
    res = []
    for name, z0 in starts(nlp):
        t = time.time()
        ok, f, viol, nit = slsqp(nlp, z0)
        res.append((name, ok, f, viol, nit, time.time() - t))
    return b, res


if __name__ == "__main__":
    sample = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    cfg = c_oracle.make_cfg("kin_cbf")
    u0, cost, st, it, _ = c_oracle.solve_batch(cfg, X0, XS, OBS, nthreads=os.cpu_count())
    cfg_r = c_oracle.make_cfg("kin_cbf", restoration=True, resto_max_calls=0)
    _, cost_r, st_r, it_r, _ = c_oracle.solve_batch(cfg_r, X0, XS, OBS, nthreads=os.cpu_count())
    fail = np.where(st == 3)[0]
    rng = np.random.default_rng(20261019)
    ids = np.sort(rng.choice(fail, size=min(sample, fail.size), replace=False))
    print(f"# configs[1]: {fail.size} of {B} scenarios end with status 3 (restoration off); sample of {ids.size}, seed 20261019")
    print("# id | dy0 = y0 - obstacle y | SLSQP from roll-out / left pulse / right pulse: ok(f) or viol | interior point with restoration (no cap): status, f")
    n_roll = n_any = n_resto = n_roll_feas = n_any_feas = 0
    with ProcessPoolExecutor(workers) as ex:
        for b, res in ex.map(work, ids):
            cells = []
            for name, ok, f, viol, nit, dt in res:
                cells.append(f"{name}: " + (f"ok f={f:.6e}" if ok else f"fail viol={viol:.2e}") + f" ({nit} it, {dt:.0f} s)")
            roll_ok = res[0][1]
            any_ok = any(r[1] for r in res)
            n_roll += roll_ok
            n_any += any_ok
            n_roll_feas += res[0][1] or res[0][3] <= 1e-6
            n_any_feas += any(r[1] or r[3] <= 1e-6 for r in res)
            n_resto += st_r[b] <= 1
            print(f"{b:5d} | {X0[b, 1] - OBS[b, 0, 0, 1]:+.2f} | " + " | ".join(cells) + f" | resto: status {st_r[b]} f={cost_r[b]:.6e}", flush=True)
    print(f"# summary: of {ids.size} sampled status-3 scenarios SLSQP reaches a feasible KKT point from the benchmark's own start on {n_roll} "
          f"({100 * n_roll / ids.size:.1f} %), from at least one of the three starts on {n_any} ({100 * n_any / ids.size:.1f} %); "
          f"the interior point with restoration succeeds on {n_resto} ({100 * n_resto / ids.size:.1f} %)")
    print(f"# SLSQP ends at a FEASIBLE point (violation <= 1e-6, converged or not) from the benchmark's start on {n_roll_feas}, from at least one "
          f"start on {n_any_feas} of {ids.size}: the scenarios are globally feasible, what fails is reaching a KKT point from this start")
