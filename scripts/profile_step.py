"""A few steps of the hot path with device-resident inputs: the command ncu wraps (launch list / full capture).

    python scripts/profile_step.py [batch] [steps] [spec] [n_k]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from awebox_b200.evaluator import Evaluator
from awebox_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = sys.argv[3] if len(sys.argv) > 3 else 'ampyx_ap2'
n_k = int(sys.argv[4]) if len(sys.argv) > 4 else None
ev = Evaluator(spec, n_k=n_k)
ocp = ev.ocp
W = synthetic.batch(ocp, B, seed=20261018 + 2)
dW = torch.from_numpy(W).cuda()
dP = torch.from_numpy(ocp.default_p()).cuda()
dG = torch.empty((B, ocp.n_g), dtype=torch.float64, device='cuda')
dJ = torch.empty((B, ocp.nnz), dtype=torch.float64, device='cuda')
for _ in range(steps):
    ev.eval_device(dW, dP, dG, dJ)
    torch.cuda.synchronize()
    print('step %.3f ms  dominant kernel %.3f ms  launches %d' % (ev.last_eval_ms(), ev.last_kernel_ms(), ev.launch_count()))
