"""Single-solve latency (BASELINE configs[0]: no-CBF kin closed loop, B = 1, warm-started) and small-batch
times with and without the all-shared small-batch kernel (MPCB_NO_LATENCY_VARIANT=1 disables it)."""
import os, sys, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    dev = torch.device("cuda:0")
    tag = "slab kernel" if os.environ.get("MPCB_NO_LATENCY_VARIANT") else "small-batch kernel"
    for kind, gen in (("kin_nocbf", scenarios.kin_nocbf), ("kin_cbf_pre", scenarios.kin_cbf_moving)):
        for B in (1, 8, 64, 512, 1036):
            x0, xs, obs = gen(B)
            s = BatchSolver(kind)
            a, b = torch.from_numpy(x0).to(dev), torch.from_numpy(xs).to(dev)
            c = torch.from_numpy(obs).to(dev) if obs.shape[1] else None
            out = s.solve(a, b, c, return_z=True); torch.cuda.synchronize()
            z = out["z"]
            ts = []
            for _ in range(30):  # warm-started re-solves from the solution (a closed-loop step without the shift)
                t0 = time.perf_counter(); o = s.solve(a, b, c, z); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            ts = np.array(ts) * 1e3
            print(f"{tag:18s} {kind:12s} B={B:5d}: p50 {np.percentile(ts, 50):.3f} ms  p99 {np.percentile(ts, 99):.3f} ms  iters {o['iters'].float().mean().item():.1f}  grid {s.launch_info()['grid']}", flush=True)
else:
    for env in ({}, {"MPCB_NO_LATENCY_VARIANT": "1"}):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env))
