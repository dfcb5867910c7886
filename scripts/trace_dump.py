"""Dump per-iteration traces of the bench batch (kin-CBF static, B=10k) for offline scheduling studies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
B = 10000
x0, xs, obs = scenarios.kin_cbf_static(B)
s = BatchSolver("kin_cbf")
tr = torch.zeros((B, 24, 8), dtype=torch.float64, device=dev)
s.set_trace(tr)
out = s.solve(*(torch.from_numpy(v).to(dev) for v in (x0, xs, obs)))
torch.cuda.synchronize()
np.savez_compressed("gpurun_out/trace_static_10k.npz", trace=tr.cpu().numpy().astype(np.float32), iters=out["iters"].cpu().numpy(), status=out["status"].cpu().numpy())
print("ok", out["iters"].float().mean().item())
