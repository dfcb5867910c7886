#!/usr/bin/env python
"""Benchmark of the MPC solve path (BASELINE.json metric: kin-CBF MPC solves/sec, N=50).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

A "step" = one pass of the hot path over one batch of synthetic scenarios: G x 10,000 kin-CBF NLPs
(static obstacle, N=50, M=1; BASELINE.json configs[1], seeded as SURVEY.md section 8d) are
solved from the reference's zero control guess.  One rank per GPU; the batch is the same on every
rank and goes through `sharding.solve_sharded`: fixed-seed shuffle, contiguous shard per rank, local
solve (no data-path collective), all-gather of the results (NCCL) - all inside the timed region.
Weak scaling: 10,000 scenarios per GPU.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kin_cbf_mpc_solves_per_sec_N50"
UNIT = "solves/s"
B_PER_GPU = 10000          # configs[1]: batch of 10k random initial states
N_HORIZON = 50
# algorithmic work per interior-point iteration of one kin-CBF scenario
FLOP_PER_ITER = 72.4e3
BYTES_PER_SOLVE_MIN = 912  # x0, xs, per-step obstacle (x,y)+(l,w) in; u0, cost, status, iters out
# dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of kin_solve_kernel<1,1,1>
# at B = 7104 (profiles/r01_final3_solve_kernel_kin_cbf_B7104.txt): 22.70 MB + 13.03 MB -> per scenario
# (inputs 2.5 KB + the part of the iterate slab that falls out of L2; L2 hit rate 96.6 %)
DRAM_BYTES_PER_SOLVE_NCU = (22704640 + 13032704) / 7104


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_batch(rank: int, B: int):
    from mpc_motion_planning_b200 import scenarios
    return scenarios.kin_cbf_static(B, N=N_HORIZON, seed=scenarios.BASE_SEED + 2 + 1000 * rank)


def bench_config(world: int, B: int) -> dict:
    """`config` of the JSON line - the same dict for this repo's arm and for the CPU reference arm"""
    return {"workload": f"kin-CBF static obstacle MPC, N={N_HORIZON}, M=1, B={B}/GPU (BASELINE configs[1])",
            "start": "zero controls, Euler roll-out states", "mu_init": 30.0, "tol": 1e-8, "max_iter": 100,
            "inputs": "x0 [B][4], xs [B][4], obstacle rows [B][1][6] as optimize_problem takes them",
            "l2": "flushed between timed steps (256 MiB memset)", "parallelism": f"scenario-sharded x{world} (solve_sharded)", "restoration": "off"}


def _stats(status, iters):
    return {"success_frac": float((status <= 1).mean()), "mean_iters": float(iters.mean()), "p99_iters": float(np.percentile(iters, 99)),
            "status_counts": np.bincount(status, minlength=6).tolist()}


def protocol_legs(dev, x0, xs, obs, K):
    """The reference's literal call protocol beside the headline (N = 1, rank 0): first guess ALL ZEROS taken as given
    (PKG/main_cbf_kin_c_sim.py:47-50,92), IPOPT's default mu_init = 0.1, restoration after a failed line search."""
    import torch

    from mpc_motion_planning_b200.solver import BatchSolver

    legs = {}
    B = x0.shape[0]
    for name, kw in (("rollout_mu30", dict()),
                     ("rollout_mu30_restoration", dict(restoration=True, resto_max_calls=1)),
                     ("zeros_mu30", dict(init="as_given")),
                     ("zeros_mu0.1", dict(init="as_given", mu_init=0.1)),
                     ("zeros_mu0.1_restoration_nocap", dict(init="as_given", mu_init=0.1, restoration=True, resto_max_calls=0))):
        s = BatchSolver("kin_cbf", N=N_HORIZON, M=1, obs_input="static", **kw)
        for _ in range(2):
            out = s.solve(x0, xs, obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            out = s.solve(x0, xs, obs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        st, it = out["status"].cpu().numpy(), out["iters"].cpu().numpy()
        leg = _stats(st, it)
        leg.update({"value": B / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "successful_solves_per_s": B / ms * 1e3 * leg["success_frac"],
                    "start": "all-zero z as given" if kw.get("init") == "as_given" else "zero controls, Euler roll-out states",
                    "mu_init": kw.get("mu_init", 30.0), "restoration": bool(kw.get("restoration", False)),
                    "resto_max_calls": kw.get("resto_max_calls")})
        legs[name] = leg
        s.close()
    return legs


def engine_legs(dev, peak_tflops):
    """The two engines side by side at 100,000 scenarios (one GPU): the row-free kinematic family (BASELINE configs[0]'s NLP),
    where the lanes of a warp do not diverge and MPCB_ENGINE_AUTO picks the lane engine, and the kin-CBF moving-obstacle
    family (configs[2]), where it stays on the warp engine.  Algorithmic flops per iteration from SURVEY.md section 8d."""
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver

    out = {}
    B = 100000
    for kind, gen, flop_it in (("kin_nocbf", scenarios.kin_nocbf, 69.8e3), ("kin_cbf_pre", scenarios.kin_cbf_moving, 72.4e3)):
        x0, xs, obs = gen(B)
        a, b_ = torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev)
        c = torch.from_numpy(obs).to(dev) if obs.shape[1] else None
        for engine in ("warp", "lane"):
            s = BatchSolver(kind, engine=engine)
            s.solve(a[:20000], b_[:20000], None if c is None else c[:20000])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = s.solve(a, b_, c)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            it = o["iters"].cpu().numpy()
            tf = float(it.sum()) * flop_it / (ms * 1e-3) / 1e12
            out[f"{kind}_B100k_{engine}"] = {"value": B / ms * 1e3, "unit": UNIT, "ms": ms, "mean_iters": float(it.mean()),
                                             "success_frac": float((o["status"].cpu().numpy() <= 1).mean()),
                                             "fp64_tflops": tf, "roofline_frac": tf / peak_tflops if peak_tflops > 0 else None}
            if engine == "lane" and kind == "kin_nocbf":
                # one scenario per resident lane (148 SMs x 3 blocks x 128 lanes on a B200): no partially filled last wave
                info = s.launch_info()
                Br = min(B, info["num_sms"] * info["blocks_per_sm"] * info["block"])
                s.solve(a[:Br], b_[:Br], None)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o = s.solve(a[:Br], b_[:Br], None)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                it = o["iters"].cpu().numpy()
                tf = float(it.sum()) * flop_it / (ms * 1e-3) / 1e12
                # this kernel is bound by HBM, not by the FP64 pipe: DRAM bytes per solve from the ncu capture of the same
                # launch (profiles/r02b_lane_kernel_kin_nocbf_B56832.txt: 62.94 GB read + 24.03 GB written for 56,832 solves)
                gbs = Br * (62.944738e9 + 24.032219e9) / 56832 / (ms * 1e-3) / 1e9
                hbm_peak, hbm_src = _peaks()
                out[f"{kind}_one_wave_{engine}"] = {"scenarios": Br, "value": Br / ms * 1e3, "unit": UNIT, "ms": ms, "mean_iters": float(it.mean()),
                                                    "fp64_tflops": tf, "roofline_frac": tf / peak_tflops if peak_tflops > 0 else None,
                                                    "hbm": {"achieved": gbs, "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s", "frac": gbs / hbm_peak,
                                                            "traffic_source": "ncu --set full dram__bytes of this launch, profiles/r02b_lane_kernel_kin_nocbf_B56832.txt"}}
            s.close()
        del a, b_, c
        torch.cuda.empty_cache()
    return out


def configs4_sweep(dev, world, rank, dist):
    """BASELINE.json configs[4]: kin-CBF moving-obstacle MPC, N in {20, 50, 100}, 4M / 4M / 2M scenarios sharded over the
    GPUs (strong scaling: the totals are fixed).  Obstacle states go in as [B][1][6] (prediction on the device); the shard is
    cut on the host from the same seeded, shuffled global batch on every rank; timed: local solve + all-gather of the results."""
    import torch

    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.sharding import balanced_permutation, shard_range
    from mpc_motion_planning_b200.solver import BatchSolver

    lines = []
    for N, total in ((20, 4_000_000), (50, 4_000_000), (100, 2_000_000)):
        rng = np.random.default_rng(scenarios.BASE_SEED + 5 + N)
        x0 = np.stack([rng.uniform(0, 20, total), rng.uniform(0, 4.5, total), rng.uniform(-0.05, 0.05, total), rng.uniform(10, 25, total)], axis=1)
        ob = np.stack([x0[:, 0] + rng.uniform(30, 80, total), rng.uniform(0, 4, total), rng.uniform(-0.05, 0.05, total), rng.uniform(5, 12, total),
                       np.full(total, 4.8), np.full(total, 1.8)], axis=1)[:, None, :]
        perm = balanced_permutation(total, 0)
        lo, hi = shard_range(total, world, rank)
        idx = perm[lo:hi]
        dx0 = torch.from_numpy(x0[idx]).to(dev)
        dxs = torch.tensor([400.0, 3.5, 0.0, 30.0], dtype=torch.float64, device=dev).repeat(hi - lo, 1)
        dob = torch.from_numpy(np.ascontiguousarray(ob[idx])).to(dev)
        del x0, ob, perm
        s = BatchSolver("kin_cbf_pre", N=N, M=1, obs_input="initial")
        per = -(-total // world)
        packed = torch.zeros((per, 5), dtype=torch.float64, device=dev)
        full = torch.empty((world * per, 5), dtype=torch.float64, device=dev) if world > 1 else packed
        s.solve(dx0[:20000], dxs[:20000], dob[:20000])  # warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out = s.solve(dx0, dxs, dob)
        e1.record()
        n = hi - lo
        packed[:n, 0:2] = out["u0"]; packed[:n, 2] = out["cost"]; packed[:n, 3] = out["status"].to(torch.float64); packed[:n, 4] = out["iters"].to(torch.float64)
        if world > 1:
            dist.all_gather_into_tensor(full, packed)
        e2.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = full[:, 3]
        lines.append({"N": N, "scenarios": total, "ms": float(t[0]), "solve_ms_max_rank": float(t[1]), "value": total / float(t[0]) * 1e3, "unit": UNIT,
                      "success_frac": float((st[: total if world == 1 else world * per] <= 1).double().mean()),
                      "gather_bytes": int(world * per * 40) if world > 1 else 0})
        s.close()
        del dx0, dxs, dob, packed, full, out
        torch.cuda.empty_cache()
    return lines


def cpu_solve(x0, xs, obs, nthreads):
    from oracle import c_oracle
    cfg = c_oracle.make_cfg("kin_cbf", N=N_HORIZON, M=1)
    t = time.perf_counter()
    u0, cost, st, it, _ = c_oracle.solve_batch(cfg, x0, xs, obs, nthreads=nthreads)
    return time.perf_counter() - t, st, it


def cpu_latency():
    """The B = 1 latency workload of `latency` (kin no-CBF closed loop, PKG/main_kin_c_sim.py) on ONE host core with the
    restated CPU IPM: 99 warm-started solves, plant Euler step and warm-start shift in numpy (not timed)."""
    from oracle import c_oracle
    cfg = c_oracle.make_cfg("kin_nocbf", N=N_HORIZON)
    N, T, L = N_HORIZON, float(cfg.T), float(cfg.Veh_l)
    x, xs, z = np.array([0.0, 0.0, 0.0, 20.0]), np.array([500.0, 3.5, 0.0, 30.0]), np.zeros(2 * N + 4 * (N + 1))
    lat = []
    for _ in range(100):
        t = time.perf_counter()
        zz, _, _ = c_oracle.solve(cfg, x, xs, None, z)
        lat.append((time.perf_counter() - t) * 1e3)
        U, X = zz[:2 * N].reshape(N, 2), zz[2 * N:].reshape(N + 1, 4)
        x = x + T * np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), x[3] * np.tan(U[0, 0]) / L, U[0, 1]])
        z = np.concatenate([np.vstack([U[1:], U[-1:]]).ravel(), np.vstack([X[1:], X[-1:]]).ravel()])
    lat = np.array(lat[1:])
    return float(np.percentile(lat, 50)), float(np.percentile(lat, 99))


def run_reference(args):
    """CPU arm.  The reference's own solver (CasADi+IPOPT) is not installable here (no wheel, no
    network), so this times the oracle port on all host cores, as the tier rules say."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 2000
    x0, xs, obs = make_batch(0, sample)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_solve(x0[:256], xs[:256], obs[:256], cores)
    tot = 0.0
    for _ in range(args.steps):
        dt, st, it = cpu_solve(x0, xs, obs, cores)
        tot += dt
    val = sample * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.gpus, B_PER_GPU),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample} scenarios of the workload per step, restated CPU IPM (oracle/mpc_oracle.c), not CasADi+IPOPT"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the literal-protocol / restoration legs")
    ap.add_argument("--no-sweep", action="store_true", help="skip the BASELINE configs[4] horizon sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from mpc_motion_planning_b200 import _lib
    from mpc_motion_planning_b200.solver import BatchSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries the one JSON line only: NCCL prints its version banner on fd 1 at the first collective, so
    # everything else that writes to fd 1 is sent to stderr and the line goes out through a saved descriptor
    out_stream = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch

    from mpc_motion_planning_b200.sharding import shard_range, solve_sharded

    # the SAME global batch of world x B scenarios on every rank (at world = 1: configs[1]'s 10,000)
    GB = world * B
    gx0, gxs, gobs_traj = make_batch(0, GB)
    # The static-obstacle module's optimize_problem takes the obstacle rows themselves, (M,6)
    # (PKG/MPC_CBF_optimize_kin.py:136,236-243; PKG/main_cbf_kin_c_sim.py:55,99): that is the input of the
    # bench, [B][M][6] (MPCB_OBS_STATIC).  The CPU arm gets the same rows repeated per step.
    gobs = np.ascontiguousarray(gobs_traj[:, :, 0, :])
    assert np.array_equal(gobs_traj, np.repeat(gobs[:, :, None, :], N_HORIZON + 1, axis=2))
    solver = BatchSolver("kin_cbf", N=N_HORIZON, M=1, obs_input="static")
    dx0, dxs, dobs = (torch.from_numpy(a).to(dev) for a in (gx0, gxs, gobs))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # local solve of this rank's shard with CUDA events around it: the gather's share of the step is the difference
    solve_ev = []

    def solve_local(a, b_, c, z):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = solver.solve(a, b_, c, z)
        e1.record()
        solve_ev.append((e0, e1))
        return o

    def step():
        return solve_sharded(solve_local, dx0, dxs, dobs, shuffle_seed=0)

    for _ in range(W):
        out = step()
    barrier()
    launches0 = solver.launch_info()["launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    solve_ev.clear()
    barrier()
    for k in range(K):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        ev[k][0].record()
        out = step()
        ev[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    step_ms = [a.elapsed_time(b) for a, b in ev]
    solve_ms = [a.elapsed_time(b) for a, b in solve_ev]
    t_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    print(f"[bench rank {rank}] device ms/step: {sum(step_ms) / K:.3f} (min {min(step_ms):.3f}, max {max(step_ms):.3f}), local solve {sum(solve_ms) / K:.3f}; "
          f"mean iterations {float(out['iters'].float().mean()):.2f}", file=sys.stderr, flush=True)
    rank_ms = [float(t_ms.item()) / K]
    rank_solve_ms = [sum(solve_ms) / K]
    if world > 1:
        both = torch.tensor([sum(step_ms) / K, sum(solve_ms) / K], dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(gathered, both)
        rank_ms = [float(g[0]) for g in gathered]
        rank_solve_ms = [float(g[1]) for g in gathered]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    launches = solver.launch_info()["launches"] - launches0
    launch_info = solver.launch_info()
    iters = out["iters"].cpu().numpy()    # the whole gathered batch, input order
    status = out["status"].cpu().numpy()
    # same work on every rank (rank 0's shard): separates machine variance from the variance of the random shards
    same_ms = None
    if world > 1:
        from mpc_motion_planning_b200.sharding import balanced_permutation
        idx0 = torch.from_numpy(balanced_permutation(GB, 0)[: shard_range(GB, world, 0)[1]]).to(dev)
        sx0, sxs, sobs = (t.index_select(0, idx0) for t in (dx0, dxs, dobs))
        for _ in range(2):
            solver.solve(sx0, sxs, sobs)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            solver.solve(sx0, sxs, sobs)
        e1.record()
        barrier()
        one = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(one) for _ in range(world)]
        dist.all_gather(gathered, one)
        same_ms = [round(float(g.item()), 3) for g in gathered]
    # this rank's own shard for the end-to-end legs below (host buffers)
    lo, hi = shard_range(GB, world, rank)
    from mpc_motion_planning_b200.sharding import balanced_permutation as _bp
    my = _bp(GB, 0)[lo:hi]
    x0, xs, obs, obs_traj = gx0[my], gxs[my], gobs[my], gobs_traj[my]
    B = hi - lo
    my_status = status[my]
    my_u0 = out["u0"].cpu().numpy()[my]
    dx0, dxs, dobs = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, xs, obs))

    # ---- end to end through the host-pointer C-ABI call (pinned host buffers, copies inside)
    hx0, hxs, hobs = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x0, xs, obs))
    hu0 = torch.empty((B, 2), dtype=torch.float64).pin_memory()
    hcost = torch.empty((B,), dtype=torch.float64).pin_memory()
    hst = torch.empty((B,), dtype=torch.int32).pin_memory()
    hit = torch.empty((B,), dtype=torch.int32).pin_memory()
    for _ in range(2):
        solver.solve_host_ptrs(B, hx0, hxs, hobs, None, hu0, hcost, hst, hit)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        solver.solve_host_ptrs(B, hx0, hxs, hobs, None, hu0, hcost, hst, hit)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_serial = world * B * K / float(e2e_s.item())

    # ---- the same K steps with two batches in flight (pipeline.PipelinedSolver: two handles, mpcb_submit_batch_host /
    # mpcb_wait): step k+1 is submitted before step k is waited for, so its head fills the SMs that step k's last
    # wave has released.  Every step still copies its inputs in and its results out inside the timed region.
    from mpc_motion_planning_b200.pipeline import PipelinedSolver
    LANES = 2
    pipe = PipelinedSolver(LANES, "kin_cbf", N=N_HORIZON, M=1, obs_input="static")
    houts = [(hu0, hcost, hst, hit)] + [tuple(torch.empty_like(t).pin_memory() for t in (hu0, hcost, hst, hit)) for _ in range(LANES - 1)]
    for k in range(2 * LANES):
        pipe.submit_host(B, hx0, hxs, hobs, None, *houts[k % LANES])
    pipe.wait()
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        pipe.submit_host(B, hx0, hxs, hobs, None, *houts[k % LANES])
    pipe.wait()
    e2e_p = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_p, op=dist.ReduceOp.MAX)
    e2e_val = world * B * K / float(e2e_p.item())
    for ho in houts:
        assert np.array_equal(ho[2].numpy(), my_status) and np.array_equal(ho[0].numpy(), my_u0), "pipelined results differ"
    # device-resident counterpart (inputs in HBM, CUDA events around the K submissions)
    for k in range(2 * LANES):
        pipe.submit(dx0, dxs, dobs)
    barrier()
    pl0 = pipe.launch_info()["launches"]
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for k in range(K):
        pipe.submit(dx0, dxs, dobs)
    for lane in range(LANES):
        pipe.result(lane)
    pe1.record()
    barrier()
    pipe_ms = torch.tensor([pe0.elapsed_time(pe1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pipe_ms, op=dist.ReduceOp.MAX)
    pipe_value = world * B * K / (float(pipe_ms.item()) * 1e-3)
    pipe_launches = pipe.launch_info()["launches"] - pl0
    h2d = int(hx0.numel() + hxs.numel() + hobs.numel()) * 8
    d2h = int(hu0.numel() + hcost.numel()) * 8 + int(hst.numel() + hit.numel()) * 4
    assert np.array_equal(hst.numpy(), my_status), "host-path and device-path verdicts differ"

    # ---- single-solve latency: BASELINE configs[0] (main_kin_c_sim.py: no-CBF kin MPC, closed loop, B = 1)
    latency = None
    if rank == 0:
        from mpc_motion_planning_b200.closed_loop import run_closed_loop  # noqa: F401
        s1 = BatchSolver("kin_nocbf", N=N_HORIZON)
        lx = torch.tensor([[0.0, 0.0, 0.0, 20.0]], dtype=torch.float64, device=dev)   # PKG/main_kin_c_sim.py:42
        lxs = torch.tensor([[500.0, 3.5, 0.0, 30.0]], dtype=torch.float64, device=dev)  # :46
        lz = torch.zeros((1, s1.nv), dtype=torch.float64, device=dev)
        lat = []
        for i in range(100):  # sim_time / T_S = 100 MPC steps (:55,70)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            o1 = s1.solve(lx, lxs, None, lz, return_z=True)
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t1) * 1e3)
            lz = o1["z"]
            s1.shift(lx, lz)
        lat = np.array(lat[1:])
        latency = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
                   "workload": "kin no-CBF closed loop (main_kin_c_sim.py), B=1, 99 warm-started solves, host wall clock incl. launch+sync"}

    # ---- the reference's literal call protocol and the restoration phase, beside the headline (one GPU only)
    legs = None
    if world == 1 and not args.no_legs:
        legs = protocol_legs(dev, dx0, dxs, dobs, max(2, min(K, 3)))
    engines = None
    if world == 1 and not args.no_legs:
        import ctypes as C0
        pk = C0.c_double(0.0)
        _lib.check(_lib.load().mpcb_fp64_peak_tflops(C0.byref(pk)), "fp64 peak")
        engines = engine_legs(dev, float(pk.value))
    # ---- BASELINE configs[4]: horizon sweep, 4M / 4M / 2M scenarios sharded over the GPUs (all ranks take part)
    sweep = None
    if not args.no_sweep:
        solver.close()
        del flush
        torch.cuda.empty_cache()
        sweep = configs4_sweep(dev, world, rank, dist if world > 1 else None)

    if rank == 0:
        value = world * B * K / (total_ms * 1e-3)
        ms_kernel = float(np.mean(solve_ms))  # the solve kernel of this rank's shard (CUDA events around the launch)
        iters_all, status_all = iters, status
        iters = iters_all[my]
        # roofline of the dominant (only) kernel: FP64 FMA pipe; HBM traffic reported beside it
        import ctypes as C
        peak = C.c_double(0.0)
        _lib.check(_lib.load().mpcb_fp64_peak_tflops(C.byref(peak)), "fp64 peak")
        flops = float(iters.sum()) * FLOP_PER_ITER
        ach_tf = flops / (ms_kernel * 1e-3) / 1e12
        hbm_peak, hbm_src = _peaks()
        ach_gbs = B * BYTES_PER_SOLVE_MIN / (ms_kernel * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(world, B),
            "solver": {"converged_frac": float((status_all <= 1).mean()), "acceptable_frac": float((status_all == 1).mean()),
                       "mean_iters": float(iters_all.mean()), "p99_iters": float(np.percentile(iters_all, 99)),
                       "status_counts": np.bincount(status_all, minlength=6).tolist(),
                       "successful_solves_per_s": value * float((status_all <= 1).mean()),
                       "restoration": "off (cfg.restoration = 0): a failed line search ends with status 3; see `protocol` for the legs with it"},
            "sharding": {"path": "sharding.solve_sharded: fixed-seed shuffle, contiguous shard per rank, local solve, all_gather_into_tensor of "
                                 "(u0, cost, status, iters) = 40 B/scenario, un-shuffle - all inside the timed step",
                         "global_batch": int(GB), "gather_bytes_per_step": int(GB * 40) if world > 1 else 0,
                         "solve_ms_per_rank": [round(v, 3) for v in rank_solve_ms],
                         "gather_and_bookkeeping_ms": round(total_ms / K - max(rank_solve_ms), 3),
                         "same_shard_on_every_rank_ms": same_ms},
            "protocol": legs,
            "engines": engines,
            "configs4": sweep,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "mode": f"{LANES} batches in flight: mpcb_submit_batch_host on {LANES} handles, step k+1 submitted before mpcb_wait of step k; "
                            "host wall clock over the K steps, every step's pinned-host copies in and out inside it",
                    "serial_value": e2e_serial,
                    "serial_mode": "one blocking mpcb_solve_batch_host per step"},
            "pipelined": {"lanes": LANES, "value": pipe_value, "unit": UNIT, "ms_per_step": float(pipe_ms.item()) / K,
                          "note": "same K steps, inputs resident, two handles on two streams, no L2 flush between steps; `value` above is the strict "
                                  "one-step-at-a-time figure", "gpu_launches": int(pipe_launches)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": float(peak.value), "unit": "TFLOP/s",
                         "frac": ach_tf / float(peak.value) if peak.value > 0 else None,
                         "traffic": int(B * DRAM_BYTES_PER_SOLVE_NCU),
                         "traffic_source": "ncu --set full capture at B=7104 scaled to this batch (profiles/r01_final3_*)",
                         "peak_source": "DFMA micro-benchmark measured in this run (mpcb_fp64_peak_tflops)",
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                 "peak_source": hbm_src + " MEASURED_PEAKS.json hbm_gbs"}},
            "ranks": {"ms_per_step": [round(v, 3) for v in rank_ms], "max_over_mean": max(rank_ms) / (sum(rank_ms) / len(rank_ms))},
            "launch": launch_info,
            "latency": latency,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sample = min(B, 4000)
            dt, st_c, it_c = cpu_solve(x0[:sample], xs[:sample], obs_traj[:sample], cores)
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} scenarios of the same batch, restated CPU IPM (oracle/mpc_oracle.c, "
                                              "scalar Riccati, one scenario per thread), not CasADi+IPOPT"}
            if latency is not None:
                p50, p99 = cpu_latency()
                latency["cpu_port"] = {"p50_ms": p50, "p99_ms": p99, "cores": 1,
                                       "note": "same closed loop, restated CPU IPM (oracle/mpc_oracle.c) called through ctypes, one host core"}
        print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
