"""Parity (vs the C checker) and timing of the library named by AWEGPU_LIB (kernel-tuning experiments).

    AWEGPU_LIB=awebox_b200/lib/var/libawegpu_ampyx_ap2.<name>.so python scripts/variant_check.py [B] [spec] [n_k]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import torch
from awebox_b200.evaluator import Evaluator
from awebox_b200 import synthetic, cabi
import build_oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spec = sys.argv[2] if len(sys.argv) > 2 else 'ampyx_ap2'
n_k = int(sys.argv[3]) if len(sys.argv) > 3 else None
ev = Evaluator(spec, n_k=n_k)
ocp = ev.ocp
W = synthetic.batch(ocp, B, seed=20261018 + 2)
p = ocp.default_p()
nb = min(B, 70)
G0, J0 = build_oracle.Oracle(ocp).eval(W[:nb], p, flags=cabi.AWEGPU_P_SHARED)
G, Jv = ev.eval(W[:nb], p)
eg = np.max(np.abs(G - G0) / np.maximum(1, np.abs(G0)))
ej = np.max(np.abs(Jv - J0) / np.maximum(1, np.abs(J0)))
dW = torch.from_numpy(W).cuda()
dP = torch.from_numpy(p).cuda()
dG = torch.empty((B, ocp.n_g), dtype=torch.float64, device='cuda')
dJ = torch.empty((B, ocp.nnz), dtype=torch.float64, device='cuda')
ms, kms = [], []
for _ in range(8):
    ev.eval_device(dW, dP, dG, dJ)
    torch.cuda.synchronize()
    ms.append(ev.last_eval_ms())
    kms.append(ev.last_kernel_ms())
print('%-40s parity g %.1e J %.1e | B=%d step %.3f ms (min %.3f)  coll-s0 %.3f ms | %.3e node-evals/s'
      % (os.path.basename(os.environ.get('AWEGPU_LIB', 'default')), eg, ej, B, float(np.median(ms[3:])), min(ms),
         float(np.median(kms[3:])), B * ocp.n_k * ocp.d / (np.median(ms[3:]) * 1e-3)))
