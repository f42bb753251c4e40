"""which degenerate inputs give a different set of non-finite results on the GPU and in the checker (diagnostic)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import get_ocp, get_oracle
from awebox_b200 import cabi, synthetic
from awebox_b200.evaluator import Evaluator
name, n_k = sys.argv[1], int(sys.argv[2])
ocp = get_ocp(name, n_k); ev = Evaluator(name, ocp=ocp)
W = synthetic.batch(ocp, 8, seed=31); p = ocp.default_p()
P = np.tile(p, (8, 1)) if ocp.needs_full_p else p
oq, _ = ocp.model.offset[('x', 'q10')]
W[1, ocp.v_cx[1][2] + oq:ocp.v_cx[1][2] + oq + 3] = 0.0
W[2, ocp.v_x[0] + oq:ocp.v_x[0] + oq + 3] = 0.0
W[3, int(ocp.tf_idx[0])] = 0.0
W[4, ocp.v_cx[0][0] + 1] = np.nan
W[5, ocp.v_u[1]] = np.inf
W[6, ocp.v_z[0]] = -np.inf
W[7, ocp.v_cx[2][1]:ocp.v_cx[2][1] + ocp.n_x] = 0.0
with np.errstate(all='ignore'):
    G0, J0 = get_oracle(name, n_k).eval(W, P, flags=0 if ocp.needs_full_p else cabi.AWEGPU_P_SHARED)
G, Jv = ev.eval(W, P)
for b in range(8):
    mg = np.isfinite(G[b]) != np.isfinite(G0[b]); mj = np.isfinite(Jv[b]) != np.isfinite(J0[b])
    print(b, 'g: nonfinite gpu %d ref %d mismatch %d | J: gpu %d ref %d mismatch %d' % ((~np.isfinite(G[b])).sum(), (~np.isfinite(G0[b])).sum(), mg.sum(), (~np.isfinite(Jv[b])).sum(), (~np.isfinite(J0[b])).sum(), mj.sum()))
    for e in np.nonzero(mj)[0][:4]:
        c = np.searchsorted(ocp.colind, e, side='right') - 1
        print('     slot', e, 'row', ocp.row[e], 'col', c, 'gpu', Jv[b, e], 'ref', J0[b, e])
    for r in np.nonzero(mg)[0][:4]:
        print('     g row', r, 'gpu', G[b, r], 'ref', G0[b, r])
