"""Times the second-order hook (awegpu_eval_hessian, device pointers) for one configuration; used under ncu."""
import sys, os, argparse, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from awebox_b200.evaluator import Evaluator
from awebox_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='ampyx_ap2')
ap.add_argument('--batch', type=int, default=4096)
ap.add_argument('--reps', type=int, default=5)
a = ap.parse_args()
ev = Evaluator(a.config)
ocp = ev.ocp
H = ev.hessian_setup()
W = synthetic.batch(ocp, a.batch, seed=1)
p = ocp.default_p()
if ev.hess_objective:
    p[:ocp.n_v] = W[0]
    p[ocp.n_v:ocp.p_theta0] = 1.0
Lam = np.random.default_rng(2).uniform(-1.0, 1.0, (a.batch, ocp.n_g))
dW, dP, dL = torch.from_numpy(W).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(Lam).cuda()
dH = torch.empty((a.batch, H.nnz), dtype=torch.float64, device='cuda')
dS = torch.ones((a.batch,), dtype=torch.float64, device='cuda') if ev.hess_objective else None
for _ in range(2):
    ev.eval_hessian_device(dW, dP, dL, dH, sigma=dS)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ev.eval_hessian_device(dW, dP, dL, dH, sigma=dS)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
nodes = a.batch * ocp.n_k * (ocp.d + 1)
byts = 8.0 * a.batch * (ocp.n_v + ocp.n_g + H.nnz)
print(json.dumps({'config': a.config, 'B': a.batch, 'ms_per_call': ms, 'nnz_hess': int(H.nnz), 'n_hv': int(H.n_hv),
                  'node_hessians_per_s': nodes / ms * 1e3, 'algorithmic_GBps': byts / ms * 1e-6}))
