"""Throughput of the kin-CBF (moving obstacle) kernel at N = 20 / 50 / 100 (layout A/B runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
for N, B in ((50, 10000), (50, 100000), (100, 50000), (20, 200000), (128, 30000)):
    gen = scenarios.kin_cbf_static if B == 10000 else scenarios.kin_cbf_moving
    x0, xs, obs = gen(B, N=N)
    s = BatchSolver("kin_cbf" if B == 10000 else "kin_cbf_pre", N=N)
    a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
    s.solve(a[:4096], b[:4096], c[:4096]); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    li = s.launch_info()
    print(f"N={N} B={B}: {best:.2f} ms -> {B / best * 1e3:.0f} solves/s (block {li['block']}, blocks/SM {li['blocks_per_sm']}, smem {li['smem_bytes']}, regs {li['regs_per_thread']})", flush=True)
