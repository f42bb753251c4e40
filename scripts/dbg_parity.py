import sys, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
from conftest import get_ocp, get_oracle, jac_tolerance
from awebox_b200.evaluator import Evaluator
from awebox_b200 import cabi
for name, nk in (('dual_kites', 3), ('dual_kites', 20), ('mpc_3dof', 3)):
    ocp = get_ocp(name, nk)
    ev = Evaluator(name, ocp=ocp)
    gold = np.load('/root/repo/tests/golden/%s_nk%d.npz' % (name, nk))
    G, Jv = ev.eval(gold['w'], gold['p'])
    bad = np.argwhere(np.abs(Jv - gold['jv']) > jac_tolerance(ocp, gold['jv']))
    badg = np.argwhere(np.abs(G - gold['g']) > 1e-12 * np.maximum(1, np.abs(gold['g'])))
    print(name, nk, 'bad J', len(bad), 'bad g', len(badg))
    colind = ocp.colind
    labels = None
    for b, s in bad[:12]:
        col = np.searchsorted(colind, s, side='right') - 1
        print('  inst', b, 'slot', s, 'row', ocp.row[s], 'col', col, 'got', Jv[b, s], 'want', gold['jv'][b, s])
