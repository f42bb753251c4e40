"""Emission of the node functions as straight-line C / CUDA code from the expression graph.

Two functions per model (``awebox_b200/ocp/nlp.py`` defines their output lists):

* ``NODE_COLL``  — model equalities at a collocation node: residual rows and every Jacobian entry
  the node contributes to ``jac_g`` (chain rule through the collocation polynomial included:
  ``d r/d X_r = d r/d xdot * C[r, j] / (h t_f)  [+ d r/d x  for the node's own block]``,
  ``d r/d t_f = d r/d theta_tf - sum_m d r/d xdot_m * xdot_m / t_f``);
* ``NODE_SHOOT`` — equality subset + path inequalities at a shooting node with their direct
  Jacobian entries.

The emitted body only uses macros (``VIN(i)``, ``PIN(i)``, ``CCIN(r)``, ``OWNIN(r)``, ``ITF``,
``OUT(n, value)``, ``RSQRT``, ``SIGN``), so the same text is compiled into the CUDA kernels
(``awebox_b200/csrc/awegpu_kernels.cu``) and into the plain-C checker (``oracle/awebox_oracle.c``).
"""
import struct

import numpy as np

from . import sym as S
from . import schedule as SCH


def _fmt(v):
    if v != v:
        return '(0.0/0.0)'
    if v in (float('inf'), float('-inf')):
        return '(1.0/0.0)' if v > 0 else '(-1.0/0.0)'
    r = repr(float(v))
    if 'e' not in r and '.' not in r:
        r += '.0'
    return r


# back-end parameters of the CUDA emission (awebox_b200/schedule.py): FP64 registers handed to the
# allocator, reload look-ahead (statements), shared-memory scratch slots per thread, statements
# between two CTA-wide re-convergence points (instruction-cache locality)
N_REGS = 48          # FP64 values the allocator keeps in registers (-> ~128 registers per thread)
LOOKAHEAD = 8        # statements between a scratch reload and its first use
LEAF_LOOKAHEAD = 40  # statements between an input load (global memory, L2 latency) and its first use
HOT_SLOTS = 1 << 20  # no cold tier: the grouped programs need only a few dozen scratch slots
CHUNK = 16           # outputs staged per flush
POOL = 0             # per-thread scratch slots shared by spills and the asynchronous input prefetch (0: plain loads; measured slower on B200: the extra shared-memory traffic saturates the MIO queue, profiles/r01_ncu_prefetch_variant.txt)
PF_AHEAD = 150       # operations between the asynchronous copy of an input and its first use
HINT_AHEAD = 0       # > 0: L2 prefetch hint this many operations before the load of a record / export input
GROUP_OPS = 1400     # weight budget of one program group (~27 KB of SASS; the L1.5 I-cache has 32 KB)
MAX_EXPORTS = 64     # intermediate values exported per stage (scaled up for large models, see NodeCode.plan)
MAX_EXPORTS_SHOOT = 0 # > 0: the same for the shooting-node function (0: as MAX_EXPORTS)
EXPORT_MIN_GAIN = 40
EXPORT_LOAD_COST = 4
MAX_STAGES = 3       # main programs + two levels of export pre-stages (at most N_STAGE_FILES)
N_STAGE_FILES = 3    # stage files the kernels include (k_*_s0..s2.inc); unused stages are empty
BATCH_MINOR_MAX_OPS = 30000   # largest model (executed FP64 operations per coll + shoot node) that also gets the batch-minor kernels


# rough SASS instruction count per operation (program size budget / export heuristic)
OP_WEIGHT = {S.POW: 60, S.POWC: 60, S.EXP: 30, S.LOG: 30, S.SIN: 30, S.COS: 30, S.TAN: 40, S.ATAN: 40, S.ATAN2: 50,
             S.ASIN: 40, S.ACOS: 40, S.DIV: 8, S.SQRT: 9, S.RSQRT: 7}


def _rhs(op, a, b):
    if op == S.ADD: return '%s + %s' % (a, b)
    if op == S.SUB: return '%s - %s' % (a, b)
    if op == S.MUL: return '%s * %s' % (a, b)
    if op == S.DIV: return 'DIVF(%s, %s)' % (a, b)
    if op == S.NEG: return '-%s' % a
    if op == S.SQRT: return 'SQRTF(%s)' % a
    if op == S.RSQRT: return 'RSQRT(%s)' % a
    if op in (S.POWC, S.POW): return 'pow(%s, %s)' % (a, b)
    if op == S.ATAN2: return 'atan2(%s, %s)' % (a, b)
    if op == S.SIGN: return 'SIGN(%s)' % a
    return '%s(%s)' % ({S.SIN: 'sin', S.COS: 'cos', S.TAN: 'tan', S.ASIN: 'asin', S.ACOS: 'acos',
                        S.ATAN: 'atan', S.EXP: 'exp', S.LOG: 'log', S.FABS: 'fabs'}[op], a)


class NodeCode(object):
    def __init__(self, ocp):
        self.ocp = ocp
        m = ocp.model
        d = ocp.d
        eq, ineq = m.eq_exprs(), m.ineq_exprs()
        local_in = ocp.local_in
        self.cc = S.symvec('cc', d + 1)
        self.own = S.symvec('own', d + 1)
        self.itf = S.sym('itf')
        j_eq = S.jacobian(eq, local_in)
        j_ineq = S.jacobian(ineq, local_in)
        rng = ocp.rng
        xdot_sym = m.vars[rng['xdot'][0]:rng['xdot'][1]]
        tf_t = ocp.tf_local - rng['theta'][0]
        scale = ocp.constraint_scale

        def sc(e):
            return e if scale == 1.0 else scale * e

        # ---- collocation node
        outs = []
        for item in ocp.coll_out:
            kind, i = item[0], item[1]
            if kind == 'res':
                e = eq[i]
            elif kind == 'theta':
                t = item[2]
                e = j_eq[i].get(rng['theta'][0] + t, S.ZERO)
                if t == tf_t:
                    for mm in ocp.eq_split[i]['xdot']:
                        b = j_eq[i].get(rng['xdot'][0] + mm, S.ZERO)
                        e = e - b * xdot_sym[mm] * self.itf
            elif kind == 'phi':
                e = j_eq[i].get(rng['phi'][0] + item[2], S.ZERO)
            elif kind == 'u':
                e = j_eq[i].get(rng['u'][0] + item[2], S.ZERO)
            elif kind == 'z':
                e = j_eq[i].get(rng['z'][0] + item[2], S.ZERO)
            elif kind == 'Xo':
                e = j_eq[i].get(rng['x'][0] + item[2], S.ZERO)
            elif kind == 'Xc':
                _, _, mm, r = item
                b = j_eq[i].get(rng['xdot'][0] + mm, S.ZERO)
                e = b * self.cc[r]
                if mm in ocp.eq_split[i]['x']:
                    e = e + j_eq[i].get(rng['x'][0] + mm, S.ZERO) * self.own[r]
            else:
                raise KeyError(kind)
            outs.append(sc(S._e(e)))
        self.coll_exprs = outs

        # ---- shooting node
        outs = []
        for item in ocp.shoot_out:
            if item[0] == 'res':
                kind, i = ocp.shoot_rows[item[1]]
                e = eq[i] if kind == 'eq' else ineq[i]
            else:
                _, n, c = item
                kind, i = ocp.shoot_rows[n]
                e = (j_eq[i] if kind == 'eq' else j_ineq[i]).get(c, S.ZERO)
            outs.append(sc(S._e(e)))
        self.shoot_exprs = outs
        self._kernel_exprs()
        self._ktab = {}                 # literal constants of the scheduled programs: hex value -> index
        self._names = {}
        for i, v in enumerate(m.vars):
            self._names[v.idx] = 'VIN(%d)' % i
        for i, v in enumerate(m.params):
            self._names[v.idx] = 'PIN(%d)' % i
        for r in range(d + 1):
            self._names[self.cc[r].idx] = 'CCIN(%d)' % r
            self._names[self.own[r].idx] = 'OWNIN(%d)' % r
        self._names[self.itf.idx] = 'ITF'

    # ------------------------------------------------------------------ kernel-side expressions
    def _kernel_exprs(self):
        """The same outputs over the leaves the CUDA node kernels read from the gathered node-input
        records (awebox_b200/csrc/awegpu_kernels.cu): ``XO(m)`` state block of the node, ``XD(m)`` its
        time derivative (collocation nodes: contracted by the gather kernel, awebox/ocp/collocation.py:
        225, 239-250; shooting nodes: the lifted xdot variables), ``UIN/ZIN/THIN/PIN`` and the per-lane
        collocation coefficients ``CCIN/OWNIN/ITF`` of the chain rule."""
        ocp, m = self.ocp, self.ocp.model
        rng, d = ocp.rng, ocp.d
        leaf = {}
        self.k_leaf = leaf              # sym idx -> macro text
        xo = [S.sym('XO(%d)' % i) for i in range(ocp.n_x)]
        xd = [S.sym('XD(%d)' % i) for i in range(ocp.n_x)]
        ui = [S.sym('UIN(%d)' % i) for i in range(ocp.n_u)]
        zi = [S.sym('ZIN(%d)' % i) for i in range(ocp.n_z)]
        th = [S.sym('THIN(%d)' % i) for i in range(ocp.n_theta)]
        pi = [S.sym('PIN(%d)' % i) for i in range(m.n_param)]
        for group in (xo, xd, ui, zi, th, pi):
            for v in group:
                leaf[v.idx] = v.name
        for r in range(d + 1):
            leaf[self.cc[r].idx] = 'CCIN(%d)' % r
            leaf[self.own[r].idx] = 'OWNIN(%d)' % r
        leaf[self.itf.idx] = 'ITF'
        base = {}
        for i in range(ocp.n_x):
            base[m.vars[rng['x'][0] + i]] = xo[i]
            base[m.vars[rng['xdot'][0] + i]] = xd[i]
        for i in range(ocp.n_u):
            base[m.vars[rng['u'][0] + i]] = ui[i]
        for i in range(ocp.n_z):
            base[m.vars[rng['z'][0] + i]] = zi[i]
        for i in range(ocp.n_theta):
            base[m.vars[rng['theta'][0] + i]] = th[i]
        for i in range(m.n_param):
            base[m.params[i]] = pi[i]
        self.k_coll = S.substitute(self.coll_exprs, base)
        self.k_shoot = S.substitute(self.shoot_exprs, base)

    @staticmethod
    def plan(exprs, limit=None, max_exports=None, min_gain=None):
        """Split a node function into small programs.

        * the outputs are cut into consecutive *groups* whose operation cones weigh at most ``limit``
          (weights ~ SASS instructions: 1 per FP64 operation, more for division, roots, pow): one group =
          one straight-line program that fits the SM's 32 KB instruction cache and needs few registers;
          what two groups share is recomputed in both;
        * *exports* are intermediate values that are too expensive to recompute in every group (the
          wind / density power laws, norms, aerodynamic angles ...): they are computed once per node by
          a pre-stage program, stored lane-contiguously, and read back as inputs.  Chosen greedily by
          (groups sharing the value - 1) x (weight of its cone) minus the cost of the extra loads; every
          round takes the best candidates whose cones neither contain nor lie inside one another (their
          gains are independent), so that large models are planned in a few dozen rounds.

        Cones are rows of a packed bit matrix (numpy), weights are summed by matrix products.

        -> (export nodes in dependency order, [(first, last+1)] output groups)"""
        limit = GROUP_OPS if limit is None else limit
        min_gain = EXPORT_MIN_GAIN if min_gain is None else min_gain
        order = [n for n in S.topo(exprs) if SCH.is_op(n)]
        n_ops = len(order)
        if n_ops == 0:
            return [], [(0, len(exprs))]
        if max_exports is None:
            # measured on B200 (AP2): 64 exports per stage beat both fewer (recomputation) and more (side-buffer
            # traffic); a model several times larger needs proportionally more to keep the group count down
            max_exports = max(MAX_EXPORTS, n_ops // 40) if n_ops > 8000 else MAX_EXPORTS
        pos = {n.idx: i for i, n in enumerate(order)}
        wgt = np.array([OP_WEIGHT.get(n.op, 1) for n in order], dtype=np.float32)
        nb = (n_ops + 7) // 8
        child = [[pos[c.idx] for c in (n.a, n.b) if c is not None and SCH.is_op(c)] for n in order]
        unit = np.zeros((n_ops, nb), dtype=np.uint8)
        unit[np.arange(n_ops), np.arange(n_ops) // 8] = (128 >> (np.arange(n_ops) % 8)).astype(np.uint8)
        exported = np.zeros(n_ops, dtype=bool)
        is_out = np.zeros(n_ops, dtype=bool)
        for e in exprs:
            if SCH.is_op(e):
                is_out[pos[e.idx]] = True
        out_pos = [pos[e.idx] if SCH.is_op(e) else -1 for e in exprs]

        def cones(cut):
            c = np.zeros((n_ops, nb), dtype=np.uint8)
            for i in range(n_ops):
                if cut is not None and cut[i]:
                    continue
                row = unit[i].copy()
                for ch in child[i]:
                    if cut is None or not cut[ch]:
                        row |= c[ch]
                c[i] = row
            return c

        def wsum_rows(rows):
            """weighted popcount of packed rows [m, nb] -> [m]"""
            out = np.empty(rows.shape[0], dtype=np.float64)
            for a0 in range(0, rows.shape[0], 1024):
                bits = np.unpackbits(rows[a0:a0 + 1024], axis=1)[:, :n_ops]
                out[a0:a0 + 1024] = bits.astype(np.float32) @ wgt
            return out

        def wsum(row):
            return float(np.unpackbits(row)[:n_ops].astype(np.float32) @ wgt)

        def partition(c):
            groups, cur, first = [], np.zeros(nb, dtype=np.uint8), 0
            for i, op in enumerate(out_pos):
                ce = c[op] if (op >= 0 and not exported[op]) else np.zeros(nb, dtype=np.uint8)
                u = cur | ce
                if i > first and wsum(u) > limit:
                    groups.append((first, i, cur))
                    cur, first = ce.copy(), i
                else:
                    cur = u
            groups.append((first, len(exprs), cur))
            return groups

        full = cones(None)         # cones without any export cut: what a pre-stage computes for an export
        pre_bits = np.zeros(nb, dtype=np.uint8)
        per_round = max(1, n_ops // 800) if n_ops > 8000 else 1
        while True:
            c = cones(exported)
            groups = partition(c)
            n_exp = int(exported.sum())
            if n_exp >= max_exports:
                break
            use = np.zeros(n_ops, dtype=np.int64)
            for g in groups:
                use += np.unpackbits(g[2])[:n_ops]
            cand = np.nonzero((use >= 2) & ~exported & ~is_out)[0]
            if len(cand) == 0:
                break
            saved = (use[cand] - 1) * wsum_rows(c[cand])
            cost = wsum_rows(full[cand] & ~pre_bits) + EXPORT_LOAD_COST * (use[cand] + 1)
            gain = saved - cost
            best = np.argsort(-gain, kind='stable')
            taken_cone = np.zeros(nb, dtype=np.uint8)       # union of the cones taken this round
            taken = []
            for bi in best:
                if gain[bi] < min_gain or gain[bi] <= 0 or len(taken) >= per_round or n_exp + len(taken) >= max_exports:
                    break
                i = int(cand[bi])
                if (taken_cone[i // 8] >> (7 - i % 8)) & 1:
                    continue                                   # inside a cone taken this round
                if np.any(c[i] & taken_cone):
                    continue                                   # contains (part of) one
                taken.append(i)
                taken_cone |= c[i]
            if not taken:
                break
            for i in taken:
                exported[i] = True
                pre_bits |= full[i]
        exports = [n for i, n in enumerate(order) if exported[i]]
        return exports, [(a, b) for a, b, _ in groups]

    def emit_scheduled(self, exprs, n_regs=N_REGS, lookahead=LOOKAHEAD, hot_slots=HOT_SLOTS, chunk=CHUNK, first=0,
                       export=False, leaf_names=None, q0=0):
        """-> (text, stats): register-allocated straight-line code (awebox_b200/schedule.py) printed
        with statement macros:
            LD_XR(reg, r, m) LD_XO(reg, m) LD_XD(reg, m) LD_U(reg, c) LD_Z(reg, c) LD_TH(reg, t)
            LD_P(reg, i) LD_CC(reg, r) LD_OWN(reg, r) LD_ITF(reg)        input loads
            SPILL_ST(slot, reg) SPILL_LD(reg, slot)                       hot scratch tier (shared memory)
            OVF_ST(slot, reg) OVF_LD(reg, slot)                           cold scratch tier
            OUT(n, value) FLUSH(chunk, count) SYNC_HINT()                 staged outputs
        Scratch slots are ranked by access count; the ``hot_slots`` busiest go to the hot tier."""
        names_ = leaf_names or self.k_leaf
        if POOL > 0:
            # inputs travel global -> scratch slot asynchronously and reach a register as late as a spill reload
            prog = SCH.allocate(exprs, n_regs, lookahead, leaf_lookahead=None)
            prog = SCH.prefetch_leaves(prog, POOL, PF_AHEAD, lambda n: '(' in names_[n.idx] and
                                       names_[n.idx].split('(')[0] not in ('CCIN', 'OWNIN'))
        else:
            prog = SCH.allocate(exprs, n_regs, lookahead, leaf_lookahead=LEAF_LOOKAHEAD)
            if HINT_AHEAD > 0:
                prog = SCH.prefetch_hints(prog, HINT_AHEAD, lambda n: names_[n.idx].split('(')[0] in
                                          ('XO', 'XD', 'UIN', 'ZIN', 'EXIN'))
        access = {}
        for it in prog.ins:
            if it[0] == 'ld':
                access[it[2]] = access.get(it[2], 0) + 1
            elif it[0] == 'st':
                access[it[1]] = access.get(it[1], 0) + 1
        ranked = sorted(access, key=lambda s_: (-access[s_], s_))
        tier = {}
        for rank, s_ in enumerate(ranked):
            tier[s_] = ('SPILL', rank) if rank < hot_slots else ('OVF', rank - hot_slots)
        n_hot = min(len(ranked), hot_slots)
        n_cold = max(0, len(ranked) - hot_slots)
        if POOL > 0:
            # one pool: slot numbers are final (the prefetch pass has placed its copies between the spills)
            tier = {s_: ('SPILL', s_) for s_ in access}
            n_hot, n_cold = prog.stats.get('pool', 0), 0
        lines = ['double ' + ', '.join('r%d' % i for i in range(prog.n_regs)) + ';']
        if not export:
            lines.append('GROUP_INIT(%d, %d, %d);' % (q0, (len(exprs) + chunk - 1) // chunk, first))

        ktab = self._ktab

        def src(s_):
            if not isinstance(s_, S.Expr):
                return 'r%d' % s_
            v = float(s_.val)
            if v != v or v in (float('inf'), float('-inf')) or \
                    (struct.unpack('<Q', struct.pack('<d', v))[0] & 0xffffffff) == 0:
                return _fmt(v)           # encodable as an immediate operand of an FP64 instruction
            # everything else comes from the constant bank (one operand, no register, no move instructions)
            return 'KC(%d)' % ktab.setdefault(v.hex(), len(ktab))

        n_stmt = 0
        n_out = len(exprs)
        cold_acc = 0
        prev_leaf = False
        for it in prog.ins:
            k = it[0]
            if prev_leaf and k != 'leaf':
                lines.append('LEAF_FENCE();')      # keeps the loads above from being sunk to their uses
            prev_leaf = k == 'leaf'
            if k == 'pfh':
                kind, args = names_[it[1].idx][:-1].split('(')
                lines.append('PFH_%s(%s);' % (kind.replace('IN', ''), args))
            elif k == 'pf':
                kind, args = names_[it[2].idx][:-1].split('(')
                lines.append('PF_%s(%d, %s);' % (kind.replace('IN', ''), it[1], args))
            elif k == 'leafr':
                lines.append('LDR(r%d, %d, %d);' % (it[1], it[3], it[4]))
            elif k == 'leaf':
                name = (leaf_names or self.k_leaf)[it[2].idx]
                if '(' in name:
                    kind, args = name[:-1].split('(')
                    lines.append('LD_%s(r%d, %s);' % (kind.replace('IN', ''), it[1], args))
                else:
                    lines.append('LD_%s(r%d);' % (name, it[1]))
            elif k == 'ld':
                t, idx = tier[it[2]]
                cold_acc += t == 'OVF'
                lines.append('%s_LD(r%d, %d);' % (t, it[1], idx))
            elif k == 'st':
                t, idx = tier[it[1]]
                cold_acc += t == 'OVF'
                lines.append('%s_ST(%d, r%d);' % (t, idx, it[2]))
            elif k == 'out':
                n = it[1]
                if export:
                    lines.append('EXOUT(%d, %s);' % (first + n, src(it[2])))
                    continue
                lines.append('OUT(%d, %s);' % (n % chunk, src(it[2])))
                if (n + 1) % chunk == 0 or n + 1 == n_out:
                    last_chunk = n + 1 == n_out
                    lines.append('FLUSH(%d, %d, %d, %d, %d);' % (
                        q0 + n // chunk, first + (n // chunk) * chunk, n % chunk + 1,
                        -1 if last_chunk else q0 + n // chunk + 1, first + (n // chunk + 1) * chunk))
            else:
                r, n, srcs = it[1], it[2], it[3]
                lines.append('r%d = %s;' % (r, _rhs(n.op, src(srcs[0]), src(srcs[1]) if len(srcs) > 1 else None)))
                n_stmt += 1
        stats = dict(prog.stats)
        stats.update(S.count_ops(exprs))
        stats.update(outputs=n_out, chunks=(n_out + chunk - 1) // chunk, hot_slots=n_hot, cold_slots=n_cold,
                     cold_accesses=cold_acc)
        return '\n'.join(lines) + '\n', stats

    # ------------------------------------------------------------------
    def emit(self, exprs):
        """-> (text, stats) straight-line code computing OUT(n, expr_n) for n = 0..len-1."""
        names = dict(self._names)
        lines = []
        n_tmp = 0
        # reference counts to inline single-use constants only
        emitted = set()

        def ref(n):
            if n.op == S.CONST:
                return _fmt(n.val)
            return names[n.idx]

        out_of = {}
        for k, e in enumerate(exprs):
            out_of.setdefault(e.idx, []).append(k)
        done_out = set()
        order = S.topo(exprs)
        pos_last_out = {}
        for n in order:
            if n.op in (S.CONST,):
                pass
            elif n.op == S.SYM:
                if n.idx not in names:
                    raise ValueError('free symbol %s in node function' % n.name)
            else:
                rhs = _rhs(n.op, ref(n.a), ref(n.b) if n.b is not None else None)
                nm = 't%d' % n_tmp
                n_tmp += 1
                names[n.idx] = nm
                lines.append('const double %s = %s;' % (nm, rhs))
            for k in out_of.get(n.idx, []):
                lines.append('OUT(%d, %s);' % (k, ref(n)))
                done_out.add(k)
        assert len(done_out) == len(exprs)
        stats = dict(S.count_ops(exprs))
        stats['statements'] = n_tmp
        stats['outputs'] = len(exprs)
        return '\n'.join(lines) + '\n', stats

    def emit_all(self):
        coll, cs = self.emit(self.coll_exprs)
        shoot, ss = self.emit(self.shoot_exprs)
        ocp = self.ocp
        head = ('// generated by awebox_b200/codegen.py from spec "%s" -- do not edit\n'
                '#define AWE_MODEL_NAME "%s"\n'
                '#define AWE_N_VAR %d\n#define AWE_N_PARAM %d\n#define AWE_N_THETA0 %d\n#define AWE_N_X %d\n'
                '#define AWE_N_U %d\n#define AWE_N_Z %d\n#define AWE_N_THETA %d\n#define AWE_N_PHI %d\n'
                '#define AWE_D %d\n#define AWE_N_EQ %d\n#define AWE_N_INEQ %d\n#define AWE_N_SHOOT_ROWS %d\n'
                '#define AWE_N_COLL_OUT %d\n#define AWE_N_SHOOT_OUT %d\n#define AWE_TF_LOCAL %d\n'
                % (ocp.spec.name, ocp.spec.name, ocp.n_var, ocp.model.n_param, ocp.n_theta0, ocp.n_x, ocp.n_u, ocp.n_z,
                   ocp.n_theta, ocp.n_phi, ocp.d, ocp.n_eq, ocp.n_ineq, len(ocp.shoot_rows),
                   len(ocp.coll_out), len(ocp.shoot_out), ocp.tf_local))
        head += '#define AWE_FLOPS_COLL %d\n#define AWE_FLOPS_SHOOT %d\n' % (flops(cs), flops(ss))
        banner = '// generated by awebox_b200/codegen.py from spec "%s" -- do not edit\n' % ocp.spec.name
        files = {'model_defs.h': head, 'node_coll.inc': banner + coll, 'node_shoot.inc': banner + shoot}
        stats = {'coll': cs, 'shoot': ss}
        # ---- CUDA emission: grouped, scheduled, register-allocated programs; full (g + jac_g) and
        #      residual-only variants.  File layout: ``case <g>: { ...program... } break;``
        n_sh = len(ocp.shoot_rows)
        variants = {'k_coll': self.k_coll, 'k_shoot': self.k_shoot,
                    'k_coll_res': self.k_coll[:ocp.n_eq], 'k_shoot_res': self.k_shoot[:n_sh]}
        kdefs = ''
        for name, exprs in variants.items():
            # stage 0 = the outputs; stage s+1 = the exports stage s (and the stages before it) read.
            # Executed in the order MAX_STAGES-1 .. 0, one launch per stage.
            stages, ex_index, ex_nodes = [], {}, []
            cur = list(exprs)
            for level in range(MAX_STAGES):
                cap = None
                if name.startswith('k_shoot') and MAX_EXPORTS_SHOOT > 0 and len(cur) and len([n for n in S.topo(cur) if SCH.is_op(n)]) <= 8000:
                    cap = MAX_EXPORTS_SHOOT
                exports, groups = self.plan(cur, max_exports=0 if level == MAX_STAGES - 1 else cap)
                new = [n for n in exports if n.idx not in ex_index]
                for n in new:
                    ex_index[n.idx] = len(ex_nodes)
                    ex_nodes.append(n)
                stages.append((cur, groups))
                if not new:
                    break
                cur = new
            leaf_names = dict(self.k_leaf)
            ex_syms = [S.sym('EXIN(%d)' % i) for i in range(len(ex_nodes))]
            for v in ex_syms:
                leaf_names[v.idx] = v.name
            sstats = []
            first_export = 0
            for level in range(N_STAGE_FILES):
                text, gstats = '', []
                if level < len(stages):
                    outs, groups = stages[level]
                    # everything exported by a deeper stage is an input leaf here
                    deeper = set()
                    for lv in range(level + 1, len(stages)):
                        deeper.update(n.idx for n in stages[lv][0])
                    mapping = {n: ex_syms[ex_index[n.idx]] for n in ex_nodes if n.idx in deeper}
                    g_exprs = S.substitute(outs, mapping) if mapping else outs
                    q0 = 0
                    chunk_first, chunk_cnt = [], []
                    for g, (a, b) in enumerate(groups):
                        if level == 0:
                            body, st = self.emit_scheduled(g_exprs[a:b], first=a, leaf_names=leaf_names, q0=q0)
                            for c0 in range(a, b, CHUNK):
                                chunk_first.append(c0)
                                chunk_cnt.append(min(CHUNK, b - c0))
                            q0 += (b - a + CHUNK - 1) // CHUNK
                        else:
                            body, st = self.emit_scheduled(g_exprs[a:b], first=ex_index[outs[a].idx], export=True,
                                                           leaf_names=leaf_names)
                            assert [ex_index[n.idx] for n in outs[a:b]] == list(range(ex_index[outs[a].idx],
                                                                                      ex_index[outs[a].idx] + b - a))
                        text += 'case %d: {\n%s} break;\n' % (g, body)
                        st['first'], st['last'] = a, b
                        gstats.append(st)
                files['%s_s%d.inc' % (name, level)] = banner + text
                sstats.append(gstats)
                if level == 0:
                    chunks0 = (chunk_first, chunk_cnt)
            allg = [st for gs in sstats for st in gs]
            hot = max(st['hot_slots'] for st in allg)
            stats[name] = {'stages': sstats, 'n_exports': len(ex_nodes),
                           'ops_total': sum(flops(st) for st in allg), 'hot_slots': hot}
            up = name.upper()
            kdefs += '#define AWE_%s_EXPORTS %d\n#define AWE_%s_HOT %d\n#define AWE_%s_OPS %d\n' % (
                up, len(ex_nodes), up, hot, up, stats[name]['ops_total'])
            for level in range(N_STAGE_FILES):
                kdefs += '#define AWE_%s_S%d_GROUPS %d\n' % (up, level, len(sstats[level]))
            kdefs += '#define AWE_%s_NCHUNKS %d\n' % (up, len(chunks0[0]))
            kdefs += '#define AWE_%s_MAXQ %d\n' % (up, max((b - a + CHUNK - 1) // CHUNK for a, b in stages[0][1]))
            kdefs += 'static const int awe_%s_chunk_first[] = {%s};\n' % (name, ', '.join(map(str, chunks0[0])))
            kdefs += 'static const int awe_%s_chunk_cnt[] = {%s};\n' % (name, ', '.join(map(str, chunks0[1])))
            gq, acc = [0], 0
            for a_, b_ in stages[0][1]:
                acc += (b_ - a_ + CHUNK - 1) // CHUNK
                gq.append(acc)
            kdefs += '#define AWE_%s_GROUP_Q0 {%s}\n' % (up, ', '.join(map(str, gq)))
            stats[name]['n_chunks'] = len(chunks0[0])
        kvals = [float.fromhex(h) for h, _ in sorted(self._ktab.items(), key=lambda kv: kv[1])] or [0.0]
        kdefs += '#define AWE_N_KCONST %d\n#define AWE_KCONST_INIT {%s}\n' % (len(kvals), ', '.join(_fmt(v) for v in kvals))
        # the batch-minor output layout doubles the output-stage kernels: compiled for small models only
        total_ops = stats['k_coll']['ops_total'] + stats['k_shoot']['ops_total']
        kdefs += '#define AWE_BATCH_MINOR %d\n' % (1 if total_ops <= BATCH_MINOR_MAX_OPS else 0)
        kdefs += '#define AWE_MAX_STAGES %d\n' % MAX_STAGES
        kdefs += '#define AWE_CHUNK %d\n' % CHUNK
        kdefs += '#define AWE_POOL %d\n' % POOL
        files['model_defs.h'] = head + kdefs
        return files, stats

    def write(self, out_dir):
        import os
        os.makedirs(out_dir, exist_ok=True)
        files, stats = self.emit_all()
        for name, text in files.items():
            path = os.path.join(out_dir, name)
            old = open(path).read() if os.path.exists(path) else None
            if old != text:
                with open(path, 'w') as f:
                    f.write(text)
        return stats


def flops(stats):
    """algorithmic flop count of an emitted function: every listed operation counts once."""
    return sum(v for k, v in stats.items() if k in S._NAMES.values())
