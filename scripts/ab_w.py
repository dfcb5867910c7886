"""A/B of the warps-per-block knob (MPCB_FORCE_W) on the bench workload."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from mpc_motion_planning_b200 import scenarios
    from mpc_motion_planning_b200.solver import BatchSolver
    dev = torch.device("cuda:0")
    for gen, kind, B in ((scenarios.kin_cbf_static, "kin_cbf", 10000), (scenarios.kin_cbf_moving, "kin_cbf_pre", 16384), (scenarios.kin_cbf_moving, "kin_cbf_pre", 100000)):
        x0, xs, obs = gen(B)
        s = BatchSolver(kind)
        a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
        s.solve(a, b, c); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        li = s.launch_info()
        print(f"W={os.environ.get('MPCB_FORCE_W','auto')} {kind} B={B}: {best:.2f} ms -> {B / best * 1e3:.0f} solves/s (block {li['block']}, blocks/SM {li['blocks_per_sm']})", flush=True)
else:
    for w in os.environ.get("AB_WS", "4,2,1").split(","):
        env = dict(os.environ, MPCB_FORCE_W=w)
        subprocess.run([sys.executable, __file__, "child"], env=env)
