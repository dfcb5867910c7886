"""BASELINE.json configs[4]: kin-CBF MPC with predicted obstacle trajectories at horizons N in {20, 50, 100},
a fixed total batch sharded over the ranks (contiguous shares, no data-path collective).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        scripts/scaling_sweep.py --total 4000000

Each rank generates its own share (rank-offset seed), keeps it resident in HBM, and the solve of the share is
timed on the device; the job time is the max over ranks.  One JSON line per horizon from rank 0, also written to
gpurun_out/scaling_sweep_G<G>.jsonl."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=4_000_000)
ap.add_argument("--horizons", default="20,50,100")
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows = []
for N in (int(v) for v in args.horizons.split(",")):
    total = args.total if N < 100 else args.total // 2     # N = 100: half the scenarios, same order of work
    share = -(-total // world)
    lo, hi = rank * share, min(total, (rank + 1) * share)
    B = hi - lo
    x0, xs, obs = scenarios.kin_cbf_moving(B, N=N, seed=scenarios.BASE_SEED + 3 + 1000 * rank)
    a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
    s = BatchSolver("kin_cbf_pre", N=N)          # the handle binds to the current device
    s.solve(a[:4096], b[:4096], c[:4096]); torch.cuda.synchronize()
    best, out = None, None
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        tmax, tsum = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        if best is None or tmax.item() < best[0]:
            best = (tmax.item(), tsum.item() / world)
    stats = torch.tensor([(out["status"] <= 1).sum().item(), out["iters"].sum().item(), B], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    if rank == 0:
        row = {"config": f"configs[4] scaling sweep, kin-CBF moving obstacle, N={N}", "n_gpus": world, "N": N, "total_scenarios": int(stats[2].item()),
               "ms_max_over_ranks": best[0], "max_over_mean": best[0] / best[1], "solves_per_s": stats[2].item() / best[0] * 1e3,
               "success_frac": stats[0].item() / stats[2].item(), "mean_iters": stats[1].item() / stats[2].item(), "timing": f"best of {args.reps}, CUDA events, inputs resident"}
        rows.append(row); print(json.dumps(row), flush=True)
    del a, b, c, s, out
    torch.cuda.empty_cache()
if rank == 0:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", f"scaling_sweep_G{world}.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
if world > 1:
    dist.destroy_process_group()
