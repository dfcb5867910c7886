"""A/B of the warps-per-block knob (MPCB_FORCE_W) for the dyn kernel."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    dev = torch.device("cuda:0")
    for B in (16384, 100000):
        x0, xs, obs = scenarios.dyn_static(B)
        s = BatchSolver("dyn")
        a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
        s.solve(a, b, c); torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        li = s.launch_info()
        print(f"W={os.environ.get('MPCB_FORCE_W','auto')} dyn B={B}: {best:.2f} ms -> {B / best * 1e3:.0f} solves/s (block {li['block']}, blocks/SM {li['blocks_per_sm']}, smem {li['smem_bytes']}, regs {li['regs_per_thread']})", flush=True)
else:
    for w in os.environ.get("AB_WS", "auto,4,2,1").split(","):
        env = dict(os.environ)
        if w != "auto":
            env["MPCB_FORCE_W"] = w
        subprocess.run([sys.executable, __file__, "child"], env=env)
