"""Soak of the batches-in-flight path: random batch sizes (buffer regrowth, small-batch kernel, tails) through
PipelinedSolver's host and device entries on 3 lanes, every result compared bit for bit with BatchSolver alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.pipeline import PipelinedSolver
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device("cuda:0")
rng = np.random.default_rng(2026)
total = bad = 0


def same_bits(a, b):
    """Bit-pattern equality: the dyn NLP reports a NaN cost for status 4, and NaN != NaN."""
    if a.dtype == torch.float64:
        return a.shape == b.shape and torch.equal(a.contiguous().view(torch.int64), b.contiguous().view(torch.int64))
    return torch.equal(a, b)


for kind, gen in (("kin_cbf_pre", scenarios.kin_cbf_moving), ("dyn", scenarios.dyn_static)):
    one = BatchSolver(kind)
    pipe = PipelinedSolver(3, kind)
    sizes = [int(v) for v in rng.choice([1, 7, 64, 500, 1036, 1037, 2500, 6000, 9000], size=40)]
    jobs = []
    for i, B in enumerate(sizes):
        x0, xs, obs = gen(B, seed=7000 + i)
        h = [torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for v in (x0, xs, obs)]
        o = (torch.empty((B, 2), dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory(),
             torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
        pipe.submit_host(B, h[0], h[1], h[2], None, *o)
        jobs.append((h, o))
    pipe.wait()
    dev_out = []
    tickets = []
    for i, (h, o) in enumerate(jobs):
        d = [t.to(dev, non_blocking=True) for t in h]
        if i >= pipe.lanes:                      # a lane holds one batch: collect it before the lane is reused
            dev_out.append({k: v.clone() for k, v in pipe.result(tickets[i - pipe.lanes]).items()})
        tickets.append(pipe.submit(*d))
    for t in tickets[-pipe.lanes:]:
        dev_out.append({k: v.clone() for k, v in pipe.result(t).items()})
    torch.cuda.synchronize()
    for (h, o), dv in zip(jobs, dev_out):
        w = one.solve(*(t.to(dev) for t in h))
        torch.cuda.synchronize()
        same = (same_bits(w["u0"].cpu(), o[0]) and same_bits(w["cost"].cpu(), o[1]) and same_bits(w["status"].cpu(), o[2])
                and same_bits(w["iters"].cpu(), o[3]) and same_bits(w["u0"], dv["u0"]) and same_bits(w["iters"], dv["iters"])
                and same_bits(w["cost"], dv["cost"]))
        total += 1
        bad += 0 if same else 1
        if not same and bad <= 3:
            B = h[0].shape[0]
            print("B", B, "host:", [bool(same_bits(w[k].cpu(), o[j])) for j, k in enumerate(("u0", "cost", "status", "iters"))],
                  "device:", [bool(same_bits(w[k], dv[k])) for k in ("u0", "cost", "status", "iters")],
                  "shapes", tuple(w["u0"].shape), tuple(dv["u0"].shape), flush=True)
    print(f"{kind}: {len(jobs)} batches, sizes {sorted(set(sizes))}", flush=True)
print(f"pipeline soak: {total} batches x (host path, device path), {bad} differ from the one-at-a-time result")
sys.exit(1 if bad else 0)
