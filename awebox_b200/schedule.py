"""Instruction scheduling and register / scratch allocation for the emitted node functions.

The node functions are straight-line FP64 programs of several thousand operations whose natural
(depth-first) order keeps ~500 values alive at once — far more than the 255 32-bit registers of an
sm_100a thread.  Left to ptxas, the excess goes to local memory (L1/L2 traffic, which made kernel v1
L2-bound).  This module is the code generator's own back end:

* ``order``      — emission order: depth-first from the outputs in their (column-major) output
                   order, with the export values of every scheduling region (``sym.region``) visited
                   first so that the values private to a region die inside it;
* ``allocate``   — a Belady (furthest-next-use) register allocation over ``n_regs`` FP64 registers
                   with a look-ahead window for reloads, explicit spill stores / reloads to numbered
                   scratch slots, leaf (input) loads on demand, and output statements in output order.

The result is a list of abstract instructions that ``codegen`` prints with macros
(``SPILL_ST/SPILL_LD`` = per-thread shared-memory scratch in the CUDA kernels).  Build-time only.
"""
from . import sym as S


def is_op(n):
    return n.op not in (S.CONST, S.SYM)


def is_leaf(n):
    return n.op == S.SYM


def operands(n):
    """distinct non-constant operands of an op node."""
    out = []
    if n.a is not None and n.a.op != S.CONST:
        out.append(n.a)
    if n.b is not None and n.b.op != S.CONST and n.b is not n.a:
        out.append(n.b)
    return out


def order(outputs):
    """-> list of op nodes, dependencies first."""
    ops = [n for n in S.topo(outputs) if is_op(n)]
    users = {}
    for n in ops:
        for c in operands(n):
            users.setdefault(c.idx, []).append(n)
    outset = set(e.idx for e in outputs)
    regions = sorted(set(n.reg for n in ops if n.reg != 0))
    roots = []
    for r in regions:
        name = S.region_name(r)
        if not name.startswith('tether'):
            continue            # only additive, mutually independent regions are worth front-loading
        roots += [n for n in ops if n.reg == r and (n.idx in outset or any(u.reg != r for u in users.get(n.idx, [])))]
    seq = [n for n in S.topo(roots + list(outputs)) if is_op(n)]
    assert len(seq) == len(ops)
    return seq


ILP_WINDOW = 0      # > 0: latency-aware list scheduling inside a sliding window of the depth-first order (interleave)
ILP_LATENCY = 8     # statements between an operation and its first consumer the interleaving aims for (DFMA: 8 clocks)


def interleave(seq, window, latency):
    """Depth-first order keeps the register pressure low but strings dependent operations together: a warp then issues
    one FP64 instruction per pipeline latency.  This pass reorders inside a sliding window of ``window`` statements of
    the depth-first order: the next statement is the earliest one of the window whose operands were issued at least
    ``latency`` statements ago (else the one whose operands are oldest).  The first statement of the window is always
    eligible, so the order stays a topological one and no value is delayed by more than ``window`` statements."""
    import os
    window = int(os.environ.get('AWE_ILP_WINDOW', window))
    latency = int(os.environ.get('AWE_ILP_LATENCY', latency))
    if window <= 1:
        return seq
    inside = set(n.idx for n in seq)
    issued = {}
    out, win, it = [], [], iter(seq)
    for n in it:
        win.append(n)
        if len(win) >= window:
            break
    while win:
        t = len(out)
        best, best_at = None, None
        for c in win:
            at, ok = 0, True
            for o in operands(c):
                if o.idx in inside:
                    if o.idx not in issued:
                        ok = False
                        break
                    at = max(at, issued[o.idx] + latency)
            if not ok:
                continue
            if at <= t:
                best = c
                break
            if best_at is None or at < best_at:
                best, best_at = c, at
        win.remove(best)
        issued[best.idx] = t
        out.append(best)
        nxt = next(it, None)
        if nxt is not None:
            win.append(nxt)
    assert len(out) == len(seq)
    return out


class Program(object):
    """abstract instruction list:
        ('leaf', reg, node)         reg <- load of an input leaf
        ('op', reg, node, srcs)     reg <- node.op(srcs...)   srcs: register numbers or Expr constants
        ('st', slot, reg)           scratch[slot] <- reg
        ('ld', reg, slot)           reg <- scratch[slot]
        ('out', n, src)             output n <- register number or constant Expr
    """

    def __init__(self):
        self.ins = []
        self.n_regs = 0
        self.n_slots = 0
        self.stats = {}


def allocate(outputs, n_regs, lookahead=12, seq=None, leaf_lookahead=None, leaf_burst=0):
    """``leaf_burst`` = W > 0: the input loads are issued in *generations*: every W statements, all inputs first
    used between W and 2 W statements later are requested together.  A thread has only six scoreboard slots for its
    outstanding loads and a consumer waits for EVERY load that shares the slot of its operand; ptxas gives the loads of
    one burst one slot, so a steady trickle of single loads (``leaf_lookahead``) makes every wait include loads that
    were issued a moment ago and exposes the full memory latency each time, whereas a generation is waited for once,
    W .. 2 W statements after it was requested."""
    seq = interleave(order(outputs), ILP_WINDOW, ILP_LATENCY) if seq is None else seq
    # ---- schedule with output statements in output order
    pos_of = {n.idx: i for i, n in enumerate(seq)}
    sched = []          # ('op', node) | ('out', n, expr)
    out_iter = 0
    n_out = len(outputs)

    def ready(e, upto):
        return (not is_op(e)) or pos_of[e.idx] <= upto

    for i, n in enumerate(seq):
        sched.append(('op', n))
        while out_iter < n_out and ready(outputs[out_iter], i):
            sched.append(('out', out_iter, outputs[out_iter]))
            out_iter += 1
    if not seq:
        while out_iter < n_out:
            sched.append(('out', out_iter, outputs[out_iter]))
            out_iter += 1
    assert out_iter == n_out

    def srcs_of(item):
        if item[0] == 'op':
            return operands(item[1])
        e = item[2]
        return [] if e.op == S.CONST else [e]

    uses = {}
    for i, item in enumerate(sched):
        for v in srcs_of(item):
            uses.setdefault(v.idx, []).append(i)
    ptr = {}

    def next_use(v, i):
        """first use position >= i (inf if none)."""
        u = uses.get(v)
        if not u:
            return 1 << 30
        p = ptr.get(v, 0)
        while p < len(u) and u[p] < i:
            p += 1
        ptr[v] = p
        return u[p] if p < len(u) else 1 << 30

    leaf_lookahead = lookahead if leaf_lookahead is None else max(lookahead, leaf_lookahead)
    prog = Program()
    ins = prog.ins
    reg_of = {}                 # value idx -> register
    val_in = [None] * n_regs    # register -> value idx
    free_regs = list(range(n_regs - 1, -1, -1))
    node_of = {}
    slot_of = {}                # value idx -> scratch slot (while stored)
    free_slots = []
    n_slots = [0]
    counts = {'st': 0, 'ld': 0, 'leaf': 0, 'op': 0}
    peak_slots = [0]

    def take_slot():
        if free_slots:
            free_slots.sort(reverse=True)
            return free_slots.pop()
        n_slots[0] += 1
        return n_slots[0] - 1

    def evict(i, pinned):
        best, best_use = None, -1
        for r in range(n_regs):
            v = val_in[r]
            if v is None or v in pinned:
                continue
            nu = next_use(v, i)
            if nu > best_use:
                best, best_use = r, nu
        if best is None:
            raise RuntimeError('register allocation failed: %d registers are not enough' % n_regs)
        v = val_in[best]
        if best_use < (1 << 30) and is_op(node_of[v]) and v not in slot_of:
            s = take_slot()
            slot_of[v] = s
            ins.append(('st', s, best))
            counts['st'] += 1
            peak_slots[0] = max(peak_slots[0], len(slot_of))
        del reg_of[v]
        val_in[best] = None
        return best

    def get_reg(i, pinned):
        if free_regs:
            return free_regs.pop()
        return evict(i, pinned)

    def bring(v_node, i, pinned):
        v = v_node.idx
        if v in reg_of:
            return
        node_of[v] = v_node
        r = get_reg(i, pinned)
        if is_leaf(v_node):
            ins.append(('leaf', r, v_node))
            counts['leaf'] += 1
        else:
            ins.append(('ld', r, slot_of[v]))
            counts['ld'] += 1
        reg_of[v] = r
        val_in[r] = v

    def release_dead(vals, i):
        for v_node in vals:
            v = v_node.idx
            if next_use(v, i + 1) == (1 << 30):
                r = reg_of.pop(v, None)
                if r is not None:
                    val_in[r] = None
                    free_regs.append(r)
                s = slot_of.pop(v, None)
                if s is not None:
                    free_slots.append(s)

    n_sched = len(sched)
    burst_end = [0]             # leaves first used before this statement have been requested (leaf_burst > 0)

    def leaves_between(a, b):
        out, seen = [], set()
        for it in sched[a:min(n_sched, b)]:
            for v in srcs_of(it):
                if is_leaf(v) and v.idx not in seen:
                    seen.add(v.idx)
                    out.append(v)
        return out

    if leaf_burst > 0:
        first = set(v.idx for v in leaves_between(0, 2 * leaf_burst))
        for v in leaves_between(0, 2 * leaf_burst):
            if len(reg_of) < n_regs - 8:
                bring(v, 0, first)
        burst_end[0] = 2 * leaf_burst
    # opening burst: every input the first window needs is requested before the first operation, so that
    # the cold misses of a program overlap instead of queueing behind one another
    elif leaf_lookahead > lookahead:
        first = set()
        for it in sched[:min(n_sched, leaf_lookahead + 1)]:
            for v in srcs_of(it):
                if is_leaf(v):
                    first.add(v.idx)
        for it in sched[:min(n_sched, leaf_lookahead + 1)]:
            for v in srcs_of(it):
                if is_leaf(v) and len(reg_of) < n_regs - 8:
                    bring(v, 0, first)
    for i, item in enumerate(sched):
        # values needed in the look-ahead window are pinned; reloads for the far end are issued now.
        # Input leaves come from global memory (L2 latency): their window is longer than the one of
        # scratch reloads (shared memory).
        pinned = set()
        for it in sched[i:min(n_sched, i + lookahead + 1)]:
            for v in srcs_of(it):
                pinned.add(v.idx)
        if leaf_burst > 0:
            # requested generations stay in their registers until they are used
            for v in leaves_between(i + lookahead + 1, burst_end[0]):
                if v.idx in reg_of:
                    pinned.add(v.idx)
            if i > 0 and i % leaf_burst == 0 and burst_end[0] < n_sched:
                nxt = leaves_between(burst_end[0], i + 2 * leaf_burst)
                for v in nxt:
                    pinned.add(v.idx)
                for v in nxt:
                    # keep a few registers for the computation: what does not fit is loaded on demand
                    if v.idx not in reg_of and sum(1 for x in pinned if x in reg_of) < n_regs - 10:
                        bring(v, i, pinned)
                burst_end[0] = i + 2 * leaf_burst
        for it in sched[i + lookahead + 1:min(n_sched, i + leaf_lookahead + 1)]:
            for v in srcs_of(it):
                if is_leaf(v) and leaf_burst == 0:
                    pinned.add(v.idx)
        if i + leaf_lookahead < n_sched and leaf_lookahead > lookahead and leaf_burst == 0:
            for v in srcs_of(sched[i + leaf_lookahead]):
                if is_leaf(v):
                    bring(v, i, pinned)
        if i + lookahead < n_sched:
            for v in srcs_of(sched[i + lookahead]):
                produced_later = is_op(v) and v.idx not in reg_of and v.idx not in slot_of
                if not produced_later:
                    bring(v, i, pinned)
        src = srcs_of(item)
        for v in src:
            if v.idx not in reg_of:
                bring(v, i, pinned)
        if item[0] == 'op':
            n = item[1]
            node_of[n.idx] = n
            regs = []
            for c in (n.a, n.b):
                if c is None:
                    continue
                regs.append(c if c.op == S.CONST else reg_of[c.idx])
            release_dead(src, i)
            if next_use(n.idx, i + 1) == (1 << 30):
                # dead value (cannot happen for a reachable node) — still needs a destination
                pass
            r = get_reg(i + 1, pinned)
            reg_of[n.idx] = r
            val_in[r] = n.idx
            ins.append(('op', r, n, regs))
            counts['op'] += 1
        else:
            e = item[2]
            ins.append(('out', item[1], e if e.op == S.CONST else reg_of[e.idx]))
            release_dead(src, i)
    prog.n_regs = n_regs
    prog.n_slots = n_slots[0]
    prog.stats = dict(counts, n_slots=n_slots[0], peak_live_slots=peak_slots[0], n_regs=n_regs, lookahead=lookahead, leaf_lookahead=leaf_lookahead)
    return prog


def interpret(prog, leaf_value, n_out):
    """reference interpreter of an allocated program (host-side test of the allocator)."""
    import math
    regs = [None] * prog.n_regs
    scratch = {}
    out = [None] * n_out

    def val(s):
        return s.val if isinstance(s, S.Expr) else regs[s]

    for it in prog.ins:
        k = it[0]
        if k == 'leaf':
            regs[it[1]] = leaf_value(it[2])
        elif k == 'ld':
            regs[it[1]] = scratch[it[2]]
        elif k == 'st':
            scratch[it[1]] = regs[it[2]]
        elif k == 'out':
            out[it[1]] = val(it[2])
        else:
            n, srcs = it[2], it[3]
            a = val(srcs[0])
            b = val(srcs[1]) if len(srcs) > 1 else None
            op = n.op
            if op == S.ADD: r = a + b
            elif op == S.SUB: r = a - b
            elif op == S.MUL: r = a * b
            elif op == S.DIV: r = a / b
            elif op == S.NEG: r = -a
            elif op == S.SQRT: r = math.sqrt(a)
            elif op == S.RSQRT: r = 1.0 / math.sqrt(a)
            elif op in (S.POW, S.POWC): r = math.pow(a, b)
            elif op == S.SIN: r = math.sin(a)
            elif op == S.COS: r = math.cos(a)
            elif op == S.TAN: r = math.tan(a)
            elif op == S.ATAN: r = math.atan(a)
            elif op == S.ASIN: r = math.asin(a)
            elif op == S.ACOS: r = math.acos(a)
            elif op == S.ATAN2: r = math.atan2(a, b)
            elif op == S.EXP: r = math.exp(a)
            elif op == S.LOG: r = math.log(a)
            elif op == S.FABS: r = abs(a)
            elif op == S.SIGN: r = float((a > 0) - (a < 0))
            else: raise NotImplementedError(op)
            regs[it[1]] = r
    return out


def prefetch_leaves(prog, pool, ahead, prefetchable, step=4):
    """Software prefetch of the input loads through the per-thread scratch slots the spills leave free.

    Every ('leaf', reg, node) whose node satisfies ``prefetchable`` becomes
        ('pf', slot, node)                issued ~``ahead`` operations earlier (asynchronous copy of the
                                          input into scratch slot ``slot``; copies complete in issue order)
        ('leafr', reg, node, slot, n)     at the old place: wait until at most ``n`` younger copies are
                                          pending, then read the slot into the register
    so that the in-flight storage of an input is a scratch slot instead of a register and the look-ahead
    can be several L2 latencies long.  Slots come from one pool of ``pool`` slots shared with the spill
    slots of the allocator: a copy may use any slot that holds no spilled value and no other pending
    input between its issue and its read."""
    import numpy as np
    ins = prog.ins
    n_ins = len(ins)
    pool = max(pool, prog.n_slots)
    occ = np.zeros((pool, n_ins + 1), dtype=bool)
    start = {}
    for i, it in enumerate(ins):
        if it[0] == 'st':
            start[it[1]] = i
            occ[it[1], i] = True
        elif it[0] == 'ld':
            occ[it[2], start[it[2]]:i + 1] = True
    opcount, c = [], 0
    for it in ins:
        opcount.append(c)
        if it[0] == 'op':
            c += 1
    first_at = {}
    for i, oc in enumerate(opcount):
        first_at.setdefault(oc, i)
    leaves = [i for i, it in enumerate(ins) if it[0] == 'leaf' and prefetchable(it[2])]
    place, info = {}, {}
    last_p, exposed, dist = 0, 0, 0
    for n, t in enumerate(leaves):
        p = max(first_at.get(max(0, opcount[t] - ahead), 0), last_p)
        slot = None
        while True:
            free = np.nonzero(~occ[:, p:t + 1].any(axis=1))[0]
            if len(free):
                slot = int(free[-1])         # high slot numbers first: the spill allocator fills from 0
                break
            if p >= t:
                break
            p = min(t, p + step)
        if slot is None:
            # no slot is free even at the use: keep the plain load
            continue
        if p >= t:
            exposed += 1
        occ[slot, p:t + 1] = True
        place.setdefault(p, []).append(('pf', slot, ins[t][2]))
        info[t] = (slot, len(info))
        last_p = p
        dist += opcount[t] - opcount[p]
    out, issued = [], 0
    for i, it in enumerate(ins):
        for pf in place.get(i, []):
            out.append(pf)
            issued += 1
        if i in info:
            slot, n = info[i]
            out.append(('leafr', it[1], it[2], slot, issued - (n + 1)))
        else:
            out.append(it)
    prog.ins = out
    prog.n_slots = pool
    used = np.nonzero(occ.any(axis=1))[0]
    prog.stats = dict(prog.stats, pool=int(used.max()) + 1 if len(used) else 0, pf=len(info), pf_exposed=exposed,
                      pf_mean_ahead=(dist / max(1, len(info))))
    return prog


def prefetch_hints(prog, ahead, wanted):
    """Insert ('pfh', node) ``ahead`` operations before every ('leaf', reg, node) with ``wanted(node)``: a
    cache prefetch of the input (no register, no scratch slot, no scoreboard), so that the load proper, which
    the allocator issues a few dozen statements before the first use, finds the line in L2 instead of HBM."""
    ins = prog.ins
    opcount, c = [], 0
    for it in ins:
        opcount.append(c)
        if it[0] == 'op':
            c += 1
    first_at = {}
    for i, oc in enumerate(opcount):
        first_at.setdefault(oc, i)
    place = {}
    n = 0
    for t, it in enumerate(ins):
        if it[0] == 'leaf' and wanted(it[2]):
            p = first_at.get(max(0, opcount[t] - ahead), 0)
            if p < t:
                place.setdefault(p, []).append(('pfh', it[2]))
                n += 1
    out = []
    for i, it in enumerate(ins):
        out.extend(place.get(i, []))
        out.append(it)
    prog.ins = out
    prog.stats = dict(prog.stats, pf_hints=n)
    return prog
