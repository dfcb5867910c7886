"""Iterations / success / throughput as a function of the initial barrier parameter, over the BASELINE configurations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
cases = (("kin-CBF static N=50", scenarios.kin_cbf_static, "kin_cbf", 50, 20000), ("kin-CBF moving N=50", scenarios.kin_cbf_moving, "kin_cbf_pre", 50, 20000),
         ("dyn N=50", scenarios.dyn_static, "dyn", 50, 20000), ("kin no-CBF N=50", scenarios.kin_nocbf, "kin_nocbf", 50, 20000),
         ("kin-CBF moving N=20", scenarios.kin_cbf_moving, "kin_cbf_pre", 20, 20000), ("kin-CBF moving N=100", scenarios.kin_cbf_moving, "kin_cbf_pre", 100, 20000))
mus = [float(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "3,10,20,30,50,100,300,1000".split(","))]
for name, gen, kind, N, B in cases:
    x0, xs, obs = gen(B, N=N)
    a, b = torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev)
    c = torch.from_numpy(obs).to(dev) if obs.shape[1] else None
    row = []
    for mu0 in mus:
        s = BatchSolver(kind, N=N, mu_init=mu0)
        s.solve(a, b, c); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
        ok = (out["status"] <= 1).float().mean().item()
        row.append(f"mu0={mu0:g}: ok {ok:.4f} it {out['iters'].float().mean().item():.1f} {B / e0.elapsed_time(e1):.0f}k/s")
    print(f"{name:22s} | " + " | ".join(row), flush=True)
