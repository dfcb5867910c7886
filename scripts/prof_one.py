"""One warm-up launch + one profiled launch of the solve kernel (used under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2960
kind = sys.argv[2] if len(sys.argv) > 2 else "kin_cbf"
gen = {"kin_cbf": scenarios.kin_cbf_static, "kin_cbf_pre": scenarios.kin_cbf_moving, "kin_nocbf": scenarios.kin_nocbf, "dyn": scenarios.dyn_static}[kind]
x0, xs, obs = gen(B)
dev = torch.device("cuda:0")
s = BatchSolver(kind)
a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = s.solve(a, b, c if obs.shape[1] else None)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} {e0.elapsed_time(e1):.2f} ms, mean iters {out['iters'].float().mean().item():.1f}")
