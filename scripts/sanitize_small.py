"""Small batch of every kernel variant (used under compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
for kind, gen, B, N in (("kin_cbf_pre", scenarios.kin_cbf_moving, 24, 20), ("kin_nocbf", scenarios.kin_nocbf, 8, 50), ("dyn", scenarios.dyn_static, 8, 20)):
    x0, xs, obs = gen(B, N=N) if kind != "kin_nocbf" else gen(B)
    s = BatchSolver(kind, N=N)
    o = s.solve(torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev), torch.from_numpy(obs).to(dev) if obs.shape[1] else None, return_z=True)
    torch.cuda.synchronize()
    print(kind, "status", o["status"].cpu().tolist())
