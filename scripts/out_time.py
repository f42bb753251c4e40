"""Times the outputs map (awegpu_eval_outputs, device pointers) for one configuration; used under ncu."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from awebox_b200.evaluator import Evaluator
from awebox_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='ampyx_ap2')
ap.add_argument('--batch', type=int, default=4096)
ap.add_argument('--reps', type=int, default=10)
a = ap.parse_args()
ev = Evaluator(a.config)
ocp = ev.ocp
W = synthetic.batch(ocp, a.batch, seed=1)
p = ocp.default_p()
n_out, n_nodes = len(ev.outputs_labels()), ocp.n_k * (ocp.d + 1)
dW, dP = torch.from_numpy(W).cuda(), torch.from_numpy(p).cuda()
dO = torch.empty((a.batch, n_out, n_nodes), dtype=torch.float64, device='cuda')
for _ in range(3):
    ev.eval_outputs_device(dW, dP, dO)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ev.eval_outputs_device(dW, dP, dO)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
gb = 8.0 * a.batch * (ocp.n_v + n_out * n_nodes) / 1e9
print('outputs: %.4f ms per call, B=%d, n_out=%d, nodes=%d, %.2f GB algorithmic -> %.0f GB/s' % (ms, a.batch, n_out, n_nodes,
                                                                                             gb, gb / ms * 1e3))
