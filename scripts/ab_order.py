import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from mpc_motion_planning_b200 import scenarios
from mpc_motion_planning_b200.solver import BatchSolver
from mpc_motion_planning_b200.sharding import balanced_permutation
dev=torch.device("cuda:0")
x0,xs,obs=scenarios.kin_cbf_static(10000)
ob=np.ascontiguousarray(obs[:,:,0,:])
s=BatchSolver("kin_cbf",obs_input="static")
flush=torch.empty(256*1024*1024,dtype=torch.uint8,device=dev)
def run(idx,name):
    t=[torch.from_numpy(np.ascontiguousarray(a[idx])).to(dev) for a in (x0,xs,ob)]
    for _ in range(3): s.solve(*t)
    ms=[]
    for _ in range(8):
        flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(*t); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print(name, "ms mean %.3f min %.3f"%(np.mean(ms),np.min(ms)))
run(np.arange(10000),"arrival order")
for seed in (0,1,2,3):
    run(balanced_permutation(10000,seed),"shuffle seed %d"%seed)
