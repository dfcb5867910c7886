import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
from oracle import c_oracle
np.set_printoptions(linewidth=200, precision=3)
gen, B, idxs = sys.argv[1], int(sys.argv[2]), [int(v) for v in sys.argv[3:]]
x0, xs, obs = getattr(scenarios, gen)(B)
dev = torch.device('cuda:0')
s = BatchSolver('kin_cbf_pre')
tr = torch.zeros((B, 104, 8), dtype=torch.float64, device=dev)
s.set_trace(tr)
out = s.solve(*(torch.from_numpy(v).to(dev) for v in (x0, xs, obs)))
torch.cuda.synchronize()
st = out['status'].cpu().numpy(); it = out['iters'].cpu().numpy()
cfg = c_oracle.make_cfg('kin_cbf_pre')
for i in idxs:
    z, lam, info = c_oracle.solve(cfg, x0[i], xs[i], obs[i])
    print('scenario', i, 'gpu status', st[i], 'iters', it[i], 'oracle', info.status, info.iters)
    t = tr[i].cpu().numpy()
    for k in range(max(0, it[i] - 12), it[i] + 1):
        print(k, ' '.join('%.3e' % v for v in t[k]))
