"""Times the objective path (awegpu_eval_objective, device pointers) for one configuration; used under ncu."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from awebox_b200.evaluator import Evaluator
from awebox_b200.ocp.objective import Objective
from awebox_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='ampyx_ap2')
ap.add_argument('--batch', type=int, default=4096)
ap.add_argument('--reps', type=int, default=10)
a = ap.parse_args()
ev = Evaluator(a.config)
ocp = ev.ocp
ob = Objective(ocp)
W = synthetic.batch(ocp, a.batch, seed=1)
p = ocp.default_p()
p[:ocp.n_v] = W[0]
p[ob.p_weights:ob.p_weights + ocp.n_var] = 1.0
p[ob.p_cost:ob.p_cost + ob.n_cost] = 1.0
dW, dP = torch.from_numpy(W).cuda(), torch.from_numpy(p).cuda()
dF = torch.empty((a.batch,), dtype=torch.float64, device='cuda')
dG = torch.empty((a.batch, ocp.n_v), dtype=torch.float64, device='cuda')
for _ in range(3):
    ev.eval_objective_device(dW, dP, dF, dG)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ev.eval_objective_device(dW, dP, dF, dG)
e1.record()
torch.cuda.synchronize()
print('objective: %.4f ms per call, B=%d, n_v=%d' % (e0.elapsed_time(e1) / a.reps, a.batch, ocp.n_v))
