"""world_size-2 gloo run of the scenario-sharded solve (host logic of the N>1 path).

The local solve is the CPU oracle here (tests may use it; the product path has no CPU
fallback): what is tested is sharding, shuffling, padding and the final gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.sharding import solve_sharded
    from oracle import c_oracle

    x0, xs, obs = scenarios.kin_cbf_moving(B, N=20)
    cfg = c_oracle.make_cfg("kin_cbf_pre", N=20)

    def solve_local(a, b, c, z):
        if a.shape[0] == 0:
            e = torch.zeros(0, dtype=torch.float64)
            return {"u0": torch.zeros((0, 2), dtype=torch.float64), "cost": e, "status": e.to(torch.int32), "iters": e.to(torch.int32)}
        u0, cost, st, it, _ = c_oracle.solve_batch(cfg, a.numpy(), b.numpy(), c.numpy(), nthreads=2)
        return {"u0": torch.from_numpy(u0), "cost": torch.from_numpy(cost), "status": torch.from_numpy(st), "iters": torch.from_numpy(it)}

    out = solve_sharded(solve_local, torch.from_numpy(x0), torch.from_numpy(xs), torch.from_numpy(obs))
    q.put((rank, {k: v.numpy() for k, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_process():
    B = 37  # not divisible by 2: exercises the padded last shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from mpc_motion_planning_b200 import scenarios
    from oracle import c_oracle

    x0, xs, obs = scenarios.kin_cbf_moving(B, N=20)
    u0, cost, st, it, _ = c_oracle.solve_batch(c_oracle.make_cfg("kin_cbf_pre", N=20), x0, xs, obs, nthreads=2)
    for r in (0, 1):  # every rank holds the full result, bit-identical to the unsharded solve
        assert np.array_equal(got[r]["u0"], u0)
        assert np.array_equal(got[r]["cost"], cost)
        assert np.array_equal(got[r]["status"], st)
        assert np.array_equal(got[r]["iters"], it)
