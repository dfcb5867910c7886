"""How much of the B = 10,000 makespan is scheduling?  Solve once, then again with the queue ordered by the
(now known) iteration counts, longest first: the upper bound of what any predictor could buy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
for B in (10000, 20000, 100000):
    x0, xs, obs = scenarios.kin_cbf_static(B)
    s = BatchSolver("kin_cbf")
    a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
    def timed():
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, out
    s.solve(a, b, c); torch.cuda.synchronize()
    t0, out = timed()
    it = out["iters"]
    s.set_order(torch.argsort(it, descending=True, stable=True).to(torch.int32))
    t1, _ = timed()
    s.set_order(torch.argsort(it, descending=False, stable=True).to(torch.int32))
    t2, _ = timed()
    s.set_order(None)
    print(f"B={B}: arrival order {B / t0 * 1e3:.0f} solves/s, longest first (oracle knowledge) {B / t1 * 1e3:.0f}, shortest first {B / t2 * 1e3:.0f}", flush=True)
