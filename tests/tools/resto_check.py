"""GPU vs C oracle with the restoration phase on: status histograms, verdict agreement, parity of the commonly
converged scenarios, timing.  python tests/tools/resto_check.py [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpc_motion_planning_b200 import scenarios  # noqa: E402
from mpc_motion_planning_b200.solver import BatchSolver  # noqa: E402
from oracle import c_oracle  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dev = torch.device("cuda:0")
for name, gen, kind in (("static", "kin_cbf_static", "kin_cbf"), ("moving", "kin_cbf_moving", "kin_cbf_pre")):
    x0, xs, obs = getattr(scenarios, gen)(B)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for K, resto in ((1, False), (1, True), (0, True)):
        s = BatchSolver(kind, restoration=resto, resto_max_calls=K)
        for _ in range(2):
            out = s.solve(t(x0), t(xs), t(obs))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = s.solve(t(x0), t(xs), t(obs))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        g = {k: v.cpu().numpy() for k, v in out.items()}
        cfg = c_oracle.make_cfg(kind, restoration=resto, resto_max_calls=K)
        u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs, obs, nthreads=os.cpu_count())
        both = (g["status"] <= 1) & (st <= 1)
        same = (g["status"] <= 1) == (st <= 1)
        du = np.abs(g["u0"] - u0).max(axis=1)
        dc = np.abs(g["cost"] - cost) / np.abs(cost)
        bad = both & ((du > 1e-4) | (dc > 1e-6))
        print(f"{name} resto={resto} K={K}: gpu status {np.bincount(g['status'], minlength=6)} oracle {np.bincount(st, minlength=6)} "
              f"verdict-equal {same.mean():.4f} ({(~same).sum()} differ) both {both.sum()} outside-tol {bad.sum()} "
              f"iters equal {(g['iters'][both] == it[both]).mean():.4f} mean it gpu {g['iters'].mean():.2f} oracle {it.mean():.2f}; "
              f"{ms:.2f} ms -> {B / ms * 1e3:.0f} solves/s", flush=True)
