"""Large-sample parity of the CUDA path against the C oracle (same seeded scenarios, same start
points): counts, worst differences and verdict agreement per configuration.  Tolerances of
BASELINE.json: first control 1e-4 absolute, cost 1e-6 relative, same converged / not verdict."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
from oracle import c_oracle

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000


def stage_ref(xs, N, seed=77):
    rng = np.random.default_rng(seed)
    ramp = 1.0 - np.linspace(0.0, 1.0, N)[None, :]
    ref = np.repeat(xs[:, None, :], N, axis=1).copy()
    ref[:, :, 1] += rng.uniform(-1.0, 1.0, (xs.shape[0], 1)) * ramp
    ref[:, :, 3] += rng.uniform(-3.0, 3.0, (xs.shape[0], 1)) * ramp
    return ref


def run(name, kind, gen, N=50, M=1, gamma=None, ref=False, seed=None, B=B, shipped=False):
    kw = {} if seed is None else {"seed": seed}
    x0, xs, obs = gen(B, N=N, **kw)
    if M == 2:
        _, _, ob = gen(B, N=N, seed=4242)
        ob[:, :, :, 0] += 60.0
        obs = np.concatenate([obs, ob], axis=1)
    xs_in = stage_ref(xs, N) if ref else xs
    s = BatchSolver(kind, N=N, M=M, cbf_gamma=gamma, ref="trajectory" if ref else "terminal", dyn_bounds="as_shipped" if shipped else "aligned")
    out = s.solve(t(x0), t(xs_in), t(obs) if obs.shape[1] else None)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in out.items()}
    cfg = c_oracle.make_cfg(kind, N=N, M=max(M, 1), cbf_gamma=gamma, ref_trajectory=ref, rows_as_shipped=shipped)
    t0 = time.time()
    u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs_in, obs if obs.shape[1] else None, nthreads=os.cpu_count())
    tc = time.time() - t0
    both = (g["status"] <= 1) & (st <= 1)
    same = (g["status"] <= 1) == (st <= 1)
    du = np.abs(g["u0"] - u0).max(axis=1)
    dc = np.abs(g["cost"] - cost) / np.maximum(1.0, np.abs(cost))
    ok = both & (du <= 1e-4) & (dc <= 1e-6)
    print(f"{name:44s} B={B:6d} both-ok {both.sum():6d}  within-tol {ok.sum():6d}  outside {int((both & ~ok).sum()):3d}  "
          f"max|du0| {du[both].max():.1e}  max dcost {dc[both].max():.1e}  verdict-equal {same.mean():.4f}  iters-equal {(g['iters'] == it).mean():.3f}  "
          f"(oracle {tc:.1f}s)", flush=True)


print(torch.cuda.get_device_name(0), "oracle threads", os.cpu_count(), flush=True)
run("kin no-CBF (MPC_optimize_kin)", "kin_nocbf", scenarios.kin_nocbf)
run("kin-CBF static (configs[1] generator)", "kin_cbf", scenarios.kin_cbf_static)
run("kin-CBF moving (configs[2] generator)", "kin_cbf_pre", scenarios.kin_cbf_moving)
run("dyn (configs[3] generator)", "dyn", scenarios.dyn_static)
run("dyn, bound lists as shipped", "dyn", scenarios.dyn_static, shipped=True)
run("kin-CBF moving N=20", "kin_cbf_pre", scenarios.kin_cbf_moving, N=20)
run("kin-CBF moving N=100", "kin_cbf_pre", scenarios.kin_cbf_moving, N=100, B=B // 2)
run("kin-CBF moving, two obstacles", "kin_cbf_pre", scenarios.kin_cbf_moving, M=2)
for gm in (1.0, 0.5, 0.2):
    run(f"discrete-time CBF rows gamma={gm}", "kin_cbf_pre", scenarios.kin_cbf_moving, gamma=gm)
run("per-stage cost targets", "kin_cbf_pre", scenarios.kin_cbf_moving, ref=True)
run("discrete-time CBF gamma=0.5 + stage targets, M=2", "kin_cbf_pre", scenarios.kin_cbf_moving, M=2, gamma=0.5, ref=True)
