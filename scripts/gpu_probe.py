"""first GPU probe: parity against the C checker + kernel timing for both Jacobian layouts."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
from awebox_b200.evaluator import Evaluator
from awebox_b200 import synthetic, cabi
import build_oracle

print(torch.cuda.get_device_name(0))
ev = Evaluator('ampyx_ap2', n_k=40)
ocp = ev.ocp
orc = build_oracle.Oracle(ocp)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = synthetic.batch(ocp, B, seed=20261018 + 2)
p = ocp.default_p()
t = time.time(); G0, J0 = orc.eval(W[:256], p, flags=cabi.AWEGPU_P_SHARED); t_cpu = time.time() - t
print('cpu checker 256 instances: %.3f s (%d threads)' % (t_cpu, os.cpu_count()))
G, Jv = ev.eval(W[:256], p)
print('parity g %.3e  J %.3e' % (np.max(np.abs(G - G0) / np.maximum(1, np.abs(G0))), np.max(np.abs(Jv - J0) / np.maximum(1, np.abs(J0)))))
Gt, Jt = ev.eval(W[:256], p, batch_minor=True)
print('parity batch-minor g %.3e  J %.3e' % (np.max(np.abs(Gt.T - G0)), np.max(np.abs(Jt.T - J0) / np.maximum(1, np.abs(J0)))))
dW = torch.from_numpy(W).cuda(); dP = torch.from_numpy(p).cuda()
dG = torch.empty(B * ocp.n_g, dtype=torch.float64, device='cuda'); dJ = torch.empty(B * ocp.nnz, dtype=torch.float64, device='cuda')
for bm in (False, True):
    for it in range(6):
        ev.eval_device(dW, dP, dG, dJ, batch_minor=bm)
        torch.cuda.synchronize()
        ms = ev.last_eval_ms()
    nodes = B * ocp.n_k * ocp.d
    byts = B * ocp.algorithmic_bytes()
    print('batch_minor=%s: %.3f ms  %.3e coll-node-evals/s  %.1f GB/s algorithmic' % (bm, ms, nodes / ms * 1e3, byts / ms / 1e6))
ev.eval_device(dW, dP, dG, None)
torch.cuda.synchronize(); print('residual only: %.3f ms' % ev.last_eval_ms())
