"""Throughput of every BASELINE.json config at its stated size on one GPU (device-timed,
inputs resident in HBM).  Writes JSON lines to gpurun_out/configs.jsonl."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver

dev = torch.device("cuda:0")
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "configs.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
rows = []

def run(name, kind, gen, B, N=50, chunk=None, **solver_kw):
    x0, xs, obs = gen(B, N=N) if kind != "kin_nocbf_" else gen(B)
    if solver_kw.get("ref") == "trajectory":  # per-stage targets: a lateral/speed offset that decays over the horizon
        rng = np.random.default_rng(77)
        ramp = 1.0 - np.linspace(0.0, 1.0, N)[None, :]
        ref = np.repeat(xs[:, None, :], N, axis=1)
        ref[:, :, 1] += rng.uniform(-1.0, 1.0, (B, 1)) * ramp
        ref[:, :, 3] += rng.uniform(-3.0, 3.0, (B, 1)) * ramp
        xs = ref
    s = BatchSolver(kind, N=N, **solver_kw)
    a, b = torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev)
    c = torch.from_numpy(obs).to(dev) if obs.shape[1] else None
    s.solve(a[:2048], b[:2048], c[:2048] if c is not None else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = s.solve(a, b, c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st, it = o["status"].cpu().numpy(), o["iters"].cpu().numpy()
    row = {"config": name, "kind": kind, "N": N, "B": B, "ms": ms, "solves_per_s": B / ms * 1e3, "success_frac": float((st <= 1).mean()),
           "maxiter_frac": float((st == 2).mean()), "fail_frac": float((st == 3).mean()), "nan_frac": float((st == 4).mean()),
           "mean_iters": float(it.mean()), "p99_iters": float(np.percentile(it, 99))}
    rows.append(row)
    print(json.dumps(row), flush=True)

run("configs[1] kin-CBF static, 10k", "kin_cbf", scenarios.kin_cbf_static, 10000)
run("configs[2] kin-CBF moving (obs_prediction), 100k", "kin_cbf_pre", scenarios.kin_cbf_moving, 100000)
run("configs[3] dyn CBF, 100k", "dyn", lambda B, N=50: scenarios.dyn_static(B, N=N), 100000)
run("configs[3] dyn CBF, bound lists as shipped, 100k", "dyn", lambda B, N=50: scenarios.dyn_static(B, N=N), 100000, dyn_bounds="as_shipped")
for N in (20, 50, 100):
    run(f"configs[4] scaling sweep N={N}, 1M (one GPU share)", "kin_cbf_pre", scenarios.kin_cbf_moving, 1000000 if N < 100 else 500000, N=N)
run("row N3: discrete-time CBF rows gamma=0.5, moving obstacle, 100k", "kin_cbf_pre", scenarios.kin_cbf_moving, 100000, cbf_gamma=0.5)
run("row N3: per-stage cost targets (aa != 0), moving obstacle, 100k", "kin_cbf_pre", scenarios.kin_cbf_moving, 100000, ref="trajectory")
with open(out_path, "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
