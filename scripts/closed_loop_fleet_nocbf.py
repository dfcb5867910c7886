"""Closed-loop throughput of a fleet on the row-free NLP (BASELINE configs[0]'s problem, main_kin_c_sim.py, at fleet scale):
B vehicles x `steps` MPC steps on the device, warp engine against the automatically chosen lane engine.
    python scripts/closed_loop_fleet_nocbf.py [B] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.closed_loop import run_closed_loop
from mpc_motion_planning_b200.solver import BatchSolver

dev = torch.device("cuda:0")
B, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 56832, int(sys.argv[2]) if len(sys.argv) > 2 else 20
x0, xs, _ = scenarios.kin_nocbf(B)
tx0, txs = torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev)
ref = None
for engine in ("warp", "auto"):
    s = BatchSolver("kin_nocbf", engine=engine)
    run_closed_loop(s, tx0[:512], txs[:512], None, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run_closed_loop(s, tx0, txs, None, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = out["iters"].float()
    same = None if ref is None else float((out["x"] - ref).abs().max())
    ref = out["x"] if ref is None else ref
    print(json.dumps({"B": B, "steps": steps, "engine": engine, "lane_kernel": s.launch_info()["smem_bytes"] == 0, "ms_total": ms,
                      "closed_loop_steps_per_s": B * steps / ms * 1e3, "mean_iters_first_step": float(it[0].mean()),
                      "mean_iters_later_steps": float(it[1:].mean()), "ok_frac": float((out["status"] <= 1).float().mean()),
                      "max_abs_state_difference_to_warp_engine": same}), flush=True)
