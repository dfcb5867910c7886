"""Closed-loop throughput (row N1 of SURVEY.md section 8f): B vehicles, `steps` MPC steps each, everything on
the device (solve -> plant Euler step -> warm-start shift -> obstacle advance), device-timed."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.closed_loop import run_closed_loop, run_closed_loop_lanes
from mpc_motion_planning_b200.solver import BatchSolver

dev = torch.device("cuda:0")
B, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 20
x0, xs, obs = scenarios.kin_cbf_moving(B)
obs0 = torch.from_numpy(obs[:, :, 0, :].copy()).to(dev)
tx0, txs = torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev)
for obs_input in ("trajectory", "initial"):
    for lf in (False, True):
        s = BatchSolver("kin_cbf_pre", obs_input=obs_input)
        run_closed_loop(s, tx0[:512], txs[:512], obs0[:512], 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = run_closed_loop(s, tx0, txs, obs0, steps, longest_first=lf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        it = out["iters"].float()
        print(json.dumps({"B": B, "steps": steps, "obs_input": obs_input, "longest_first": lf, "ms_total": ms,
                          "closed_loop_steps_per_s": B * steps / ms * 1e3, "mean_iters_first_step": float(it[0].mean()),
                          "mean_iters_later_steps": float(it[1:].mean()),
                          "ok_frac": float((out["status"] <= 1).float().mean())}), flush=True)

# the fleet split into independent groups, one handle + stream each (batches in flight)
for lanes in (2, 3, 4):
    ss = [BatchSolver("kin_cbf_pre", obs_input="initial") for _ in range(lanes)]
    run_closed_loop_lanes(ss, tx0[:1024], txs[:1024], obs0[:1024], 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out2 = run_closed_loop_lanes(ss, tx0, txs, obs0, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    same = bool(torch.equal(out2["u"], out["u"]) and torch.equal(out2["status"], out["status"]))
    print(json.dumps({"B": B, "steps": steps, "obs_input": "initial", "lanes": lanes, "ms_total": ms,
                      "closed_loop_steps_per_s": B * steps / ms * 1e3, "identical_to_one_handle": same}), flush=True)
