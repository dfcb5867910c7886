import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
dev = torch.device('cuda:0')
for gen in ('kin_cbf_static', 'kin_cbf_moving'):
    B = 4000
    x0, xs, obs = getattr(scenarios, gen)(B)
    s = BatchSolver('kin_cbf_pre')
    tr = torch.zeros((B, 104, 8), dtype=torch.float64, device=dev)
    s.set_trace(tr)
    out = s.solve(*(torch.from_numpy(v).to(dev) for v in (x0, xs, obs)))
    torch.cuda.synchronize()
    st = out['status'].cpu().numpy(); it = out['iters'].cpu().numpy(); t = tr.cpu().numpy()
    fails = np.where(st == 3)[0]
    err = np.array([t[i, it[i], 2] for i in fails])
    print(gen, 'B', B, 'conv', (st == 0).mean(), 'fail3', len(fails), 'maxiter', (st == 2).sum(), 'near-solution fails (err<1e-5):', (err < 1e-5).sum(), (err<1e-3).sum())
    mi = np.where(st == 2)[0]
    errm = np.array([t[i, min(it[i],103), 2] for i in mi])
    print('   maxiter errs', np.sort(errm)[:10])
