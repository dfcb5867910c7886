"""GPU check of the discrete-time CBF rows / per-stage reference options + timing beside the plain rows."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
from oracle import c_oracle

dev = torch.device('cuda:0')
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

def stage_ref(xs, N, seed=77):
    rng = np.random.default_rng(seed)
    B = xs.shape[0]
    ref = np.repeat(xs[:, None, :], N, axis=1).copy()
    ramp = np.linspace(0.0, 1.0, N)[None, :]
    ref[:, :, 1] += rng.uniform(-1.0, 1.0, (B, 1)) * (1.0 - ramp)
    ref[:, :, 3] += rng.uniform(-3.0, 3.0, (B, 1)) * (1.0 - ramp)
    return ref

def compare(gamma, ref, B=512):
    x0, xs, obs = scenarios.kin_cbf_moving(B)
    xs_in = stage_ref(xs, 50) if ref == 'trajectory' else xs
    s = BatchSolver('kin_cbf_pre', cbf_gamma=gamma, ref=ref)
    out = s.solve(t(x0), t(xs_in), t(obs)); torch.cuda.synchronize()
    u0 = out['u0'].cpu().numpy(); cost = out['cost'].cpu().numpy(); st = out['status'].cpu().numpy(); it = out['iters'].cpu().numpy()
    cfg = c_oracle.make_cfg('kin_cbf_pre', cbf_gamma=gamma, ref_trajectory=(ref == 'trajectory'))
    ou0, ocost, ost, oit, _ = c_oracle.solve_batch(cfg, x0, xs_in, obs, nthreads=os.cpu_count())
    same = (st <= 1) == (ost <= 1); both = (st <= 1) & (ost <= 1)
    du = np.abs(u0 - ou0).max(axis=1); dc = np.abs(cost - ocost) / np.maximum(1.0, np.abs(ocost))
    okp = both & (du <= 1e-4) & (dc <= 1e-6)
    print(f"gamma={gamma} ref={ref}: verdict equal {same.mean():.4f} both {both.mean():.4f} parity {okp.sum()}/{both.sum()} "
          f"max du {du[both].max():.2e} dc {dc[both].max():.2e} iters equal {(it == oit).mean():.3f} mean it {it.mean():.1f}/{oit.mean():.1f}", flush=True)

def timing(gamma, ref, B=16384):
    x0, xs, obs = scenarios.kin_cbf_moving(B)
    xs_in = stage_ref(xs, 50) if ref == 'trajectory' else xs
    s = BatchSolver('kin_cbf_pre', cbf_gamma=gamma, ref=ref)
    a, b, c = t(x0), t(xs_in), t(obs)
    s.solve(a, b, c); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = s.solve(a, b, c); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"timing gamma={gamma} ref={ref} B={B}: {ms:.2f} ms -> {B / ms * 1e3:.0f} solves/s; iters {out['iters'].float().mean().item():.1f}; "
          f"ok {(out['status'] <= 1).float().mean().item():.3f}; {s.launch_info()}", flush=True)

if __name__ == '__main__':
    for g, r in [(None, 'terminal'), (1.0, 'terminal'), (0.4, 'terminal'), (None, 'trajectory'), (0.6, 'trajectory')]:
        compare(g, r)
    for g, r in [(None, 'terminal'), (0.5, 'terminal'), (None, 'trajectory')]:
        timing(g, r)
