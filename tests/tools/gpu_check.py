"""First-contact GPU check: CUDA path vs the C oracle on seeded scenarios + a rough timing."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
from oracle import c_oracle

def compare(kind, gen, B, **kw):
    x0, xs, obs = gen(B)
    s = BatchSolver(kind, **kw)
    print(kind, 'launch info', s.launch_info())
    dev = torch.device('cuda:0')
    t = lambda a: torch.from_numpy(a).to(dev)
    out = s.solve(t(x0), t(xs), t(obs) if obs.shape[1] else None, return_z=True)
    torch.cuda.synchronize()
    u0 = out['u0'].cpu().numpy(); cost = out['cost'].cpu().numpy(); st = out['status'].cpu().numpy(); it = out['iters'].cpu().numpy()
    cfg = c_oracle.make_cfg(kind, N=s.N, M=max(s.M, 1), mu_init=kw.get('mu_init', 30.0))
    ou0, ocost, ost, oit, _ = c_oracle.solve_batch(cfg, x0, xs, obs if obs.shape[1] else None, nthreads=os.cpu_count())
    same_st = (st <= 1) == (ost <= 1)
    both = (st <= 1) & (ost <= 1)
    du = np.abs(u0 - ou0).max(axis=1)
    dc = np.abs(cost - ocost) / np.maximum(1.0, np.abs(ocost))
    okp = both & (du <= 1e-4) & (dc <= 1e-6)
    print(f"{kind}: B={B} status equal {same_st.mean():.4f}  both-converged {both.mean():.4f}  parity among both {okp.sum()}/{both.sum()}"
          f"  max du0 {du[both].max() if both.any() else 0:.2e} max dcost {dc[both].max() if both.any() else 0:.2e}"
          f"  iters equal {(it == oit).mean():.3f} gpu mean iters {it.mean():.1f} oracle {oit.mean():.1f}")
    bad = np.where(~same_st | (both & ~okp))[0][:10]
    for i in bad:
        print('   mismatch', i, 'gpu', st[i], it[i], cost[i], u0[i], 'oracle', ost[i], oit[i], ocost[i], ou0[i])
    return s, (x0, xs, obs)

if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    compare('kin_nocbf', scenarios.kin_nocbf, 64)
    compare('dyn', scenarios.dyn_static, 256)
    compare('kin_cbf', scenarios.kin_cbf_static, 256)
    s, (x0, xs, obs) = compare('kin_cbf_pre', scenarios.kin_cbf_moving, 512)
    # rough timing
    for B in (2048, 16384):
        x0, xs, obs = scenarios.kin_cbf_moving(B)
        dev = torch.device('cuda:0')
        a, b, c = (torch.from_numpy(v).to(dev) for v in (x0, xs, obs))
        s.solve(a, b, c); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"B={B}: {ms:.2f} ms -> {B / ms * 1e3:.0f} solves/s; mean iters {out['iters'].float().mean().item():.1f}; conv {(out['status'] <= 1).float().mean().item():.3f}")
