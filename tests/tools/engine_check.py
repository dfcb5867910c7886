"""Lane-per-scenario engine vs warp-per-scenario engine vs the C oracle (parity, iteration counts, timing).
  MPCB_ENGINE=lane|warp python tests/tools/engine_check.py [B] [kind]"""
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
kinds = sys.argv[2:] or ["kin_cbf", "kin_cbf_pre", "kin_nocbf"]
GEN = {"kin_cbf": "kin_cbf_static", "kin_cbf_pre": "kin_cbf_moving", "kin_nocbf": "kin_nocbf"}
dev = torch.device("cuda:0")
for kind in kinds:
    x0, xs, obs = getattr(scenarios, GEN[kind])(B)
    t = lambda a: None if a is None or a.shape[1] == 0 else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    s = BatchSolver(kind)
    tx0, txs, tobs = t(x0), t(xs), t(obs)
    for _ in range(2):
        out = s.solve(tx0, txs, tobs)
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = s.solve(tx0, txs, tobs)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    g = {k: v.cpu().numpy() for k, v in out.items()}
    info = s.launch_info()
    line = f"{os.environ.get('MPCB_ENGINE', 'auto')} {kind} B={B}: {np.mean(ms):.2f} ms -> {B / np.mean(ms) * 1e3:.0f} solves/s; grid {info['grid']}x{info['block']} regs {info['regs_per_thread']}; status {np.bincount(g['status'], minlength=6)} mean it {g['iters'].mean():.2f}"
    if B <= 20000:
        cfg = c_oracle.make_cfg(kind)
        u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs, obs if obs.shape[1] else None, nthreads=os.cpu_count())
        both = (g["status"] <= 1) & (st <= 1)
        same = (g["status"] <= 1) == (st <= 1)
        du = np.abs(g["u0"] - u0).max(axis=1)
        dc = np.abs(g["cost"] - cost) / np.abs(cost)
        bad = both & ((du > 1e-4) | (dc > 1e-6))
        line += f"; oracle {np.bincount(st, minlength=6)} verdict-equal {same.mean():.4f} both {both.sum()} outside-tol {bad.sum()} worst du {du[both].max():.1e} dc {dc[both].max():.1e} iters equal {(g['iters'][both] == it[both]).mean():.4f}"
    print(line, flush=True)
