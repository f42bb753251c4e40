"""DAG statistics of the emitted node functions: op counts, live ranges in emission order."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awebox_b200.spec import ModelSpec
from awebox_b200.ocp.nlp import OcpProblem
from awebox_b200.codegen import NodeCode
from awebox_b200 import sym as S

def live_stats(order, outputs):
    pos = {n.idx: i for i, n in enumerate(order)}
    last = {}
    for i, n in enumerate(order):
        for c in (n.a, n.b):
            if c is not None and c.op not in (S.CONST, S.SYM):
                last[c.idx] = i
    for e in outputs:
        if e.idx in pos: last.setdefault(e.idx, pos[e.idx])
    ops = [n for n in order if n.op not in (S.CONST, S.SYM)]
    events = [0] * (len(order) + 2)
    for n in ops:
        events[pos[n.idx]] += 1
        events[last.get(n.idx, pos[n.idx]) + 1] -= 1
    live, mx, tot = 0, 0, 0
    hist = []
    for i in range(len(order) + 1):
        live += events[i]
        mx = max(mx, live)
        tot += live
        hist.append(live)
    return mx, tot / len(order), hist

if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else 'ampyx_ap2'
    ocp = OcpProblem(ModelSpec.load(name))
    nc = NodeCode(ocp)
    for label, exprs in (('coll', nc.coll_exprs), ('shoot', nc.shoot_exprs)):
        order = S.topo(exprs)
        nops = sum(1 for n in order if n.op not in (S.CONST, S.SYM))
        mx, avg, hist = live_stats(order, exprs)
        n_res = ocp.n_eq if label == 'coll' else len(ocp.shoot_rows)
        res_order = S.topo(exprs[:n_res])
        nres = sum(1 for n in res_order if n.op not in (S.CONST, S.SYM))
        print(label, 'ops', nops, 'outputs', len(exprs), 'residual-only ops', nres, 'max live', mx, 'avg live %.1f' % avg)
        print('  live profile (every 400):', hist[::400])
