"""Experiment: register-pressure-aware ordering + Belady spill simulation of the node functions."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awebox_b200.spec import ModelSpec
from awebox_b200.ocp.nlp import OcpProblem
from awebox_b200.codegen import NodeCode
from awebox_b200 import sym as S

def isop(n): return n.op not in (S.CONST, S.SYM)
def kids(n):
    out = []
    if n.a is not None and isop(n.a): out.append(n.a)
    if n.b is not None and isop(n.b) and n.b is not n.a: out.append(n.b)
    return out

def belady(sched, R):
    uses = collections.defaultdict(list)
    for i, n in enumerate(sched):
        for ch in kids(n): uses[ch.idx].append(i)
    ptr = collections.defaultdict(int)
    def next_use(v, i):
        u = uses.get(v, ()); p = ptr[v]
        while p < len(u) and u[p] < i: p += 1
        ptr[v] = p
        return u[p] if p < len(u) else 1 << 30
    inreg = set(); inmem = set(); stores = loads = 0; peak = 0
    for i, n in enumerate(sched):
        ops = [c.idx for c in kids(n)]
        for v in ops:
            if v not in inreg:
                loads += 1; inreg.add(v)
        for v in ops:
            if next_use(v, i + 1) == 1 << 30:
                inreg.discard(v); inmem.discard(v)
        if n.idx in uses: inreg.add(n.idx)
        while len(inreg) > R:
            victim = max(inreg, key=lambda v: next_use(v, i + 1))
            inreg.discard(victim)
            if victim not in inmem:
                inmem.add(victim); stores += 1
        peak = max(peak, len(inmem))
    return stores, loads, peak

if __name__ == '__main__':
    ocp = OcpProblem(ModelSpec.load('ampyx_ap2'))
    nc = NodeCode(ocp)
    exprs = nc.coll_exprs
    items = ocp.coll_out
    rng = ocp.rng
    def colkey(it):
        k = it[0]
        if k == 'res': return (0, it[1])
        if k == 'Xo': return (1, it[2], it[1])
        if k == 'Xc': return (1, it[2], it[1], it[3])
        if k == 'u': return (2, it[2], it[1])
        if k == 'z': return (3, it[2], it[1])
        if k == 'theta': return (4, it[2], it[1])
        return (5, it[2], it[1])
    def rowkey(it):
        return (it[1],) + tuple(colkey(it))
    for name, key in (('given', None), ('col-major', colkey), ('row-major', rowkey)):
        idx = list(range(len(exprs)))
        if key: idx.sort(key=lambda i: key(items[i]))
        order = [n for n in S.topo([exprs[i] for i in idx]) if isop(n)]
        for R in (64, 96, 110):
            print(name, 'R', R, belady(order, R))
