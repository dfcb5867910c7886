"""Row (e) on real GPUs: `sharding.solve_sharded` under torchrun/NCCL must return, on every rank, exactly what one GPU
returns for the whole batch (scenarios are independent, so the comparison is bit for bit).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 scripts/sharded_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.sharding import solve_sharded
from mpc_motion_planning_b200.solver import BatchSolver

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for B in (10001, 37):
    x0, xs, obs = scenarios.kin_cbf_moving(B, seed=31337)
    a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
    s = BatchSolver("kin_cbf_pre")
    got = solve_sharded(lambda x, y, o, z: s.solve(x, y, o, z), a, b, c)
    want = s.solve(a, b, c)
    same = all(torch.equal(got[k], want[k]) for k in ("u0", "status", "iters")) and \
        torch.equal(got["cost"].view(torch.int64), want["cost"].contiguous().view(torch.int64))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = ok and bool(flag.item())
    if rank == 0:
        print(f"solve_sharded over {world} GPUs, B={B}: identical to the one-GPU solve on every rank: {bool(flag.item())}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
