"""Independent batches in flight on several handles/streams: does the next batch's head hide the last wave of the
previous one?  kin-CBF static, N=50, B per batch as given; device-timed over K batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
K = 24
for B in (2000, 10000, 20000):
    x0, xs, obs = scenarios.kin_cbf_static(B)
    obs = np.ascontiguousarray(obs[:, :, 0, :])
    a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
    for lanes in (1, 2, 3, 4):
        solvers = [BatchSolver("kin_cbf", obs_input="static") for _ in range(lanes)]
        streams = [torch.cuda.Stream(dev) for _ in range(lanes)]
        for s, st in zip(solvers, streams):
            with torch.cuda.stream(st):
                s.solve(a, b, c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for st in streams:
                st.wait_stream(torch.cuda.current_stream(dev))
            outs = []
            for k in range(K):
                with torch.cuda.stream(streams[k % lanes]):
                    outs.append(solvers[k % lanes].solve(a, b, c))
            for st in streams:
                torch.cuda.current_stream(dev).wait_stream(st)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ok = all(torch.equal(o["status"], outs[0]["status"]) and torch.equal(o["u0"], outs[0]["u0"]) for o in outs)
        print(f"B={B} lanes={lanes}: {best / K:.3f} ms/batch -> {B * K / best * 1e3:.0f} solves/s  results identical across batches: {ok}", flush=True)
        del solvers
