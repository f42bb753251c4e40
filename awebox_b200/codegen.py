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
N_REGS = 48          # FP64 values the allocator keeps in registers (-> 168 registers per thread, three quads per SM)
N_REGS_FAT = 80      # the same when only two quads fit an SM (255 registers per thread)
SMEM_PER_SM = 227 * 1024
LOOKAHEAD = 8        # statements between a scratch reload and its first use
LEAF_LOOKAHEAD = 40  # statements between an input load (global memory, L2 latency) and its first use
LEAF_BURST = 40       # > 0: input loads are issued in generations of this many statements (schedule.allocate)
HOT_SLOTS = 1 << 20  # no cold tier: the grouped programs need only a few dozen scratch slots
CHUNK = 16           # outputs staged per flush
OWN_CAP = 24         # entries staged per flush of a node's own CCS range (whole columns; raised to the largest column)
POOL = 0             # per-thread scratch slots shared by spills and the asynchronous input prefetch (0: plain loads; measured slower on B200: the extra shared-memory traffic saturates the MIO queue, profiles/r01_ncu_prefetch_variant.txt)
PF_AHEAD = 150       # operations between the asynchronous copy of an input and its first use
HINT_AHEAD = 0       # > 0: L2 prefetch hint this many operations before the load of a record / export input
GROUP_OPS = 900      # weight budget of one program group (~27 KB of SASS; the L1.5 I-cache has 32 KB)
_GROUP_OPS_DEFAULT = GROUP_OPS    # (an experiment that changes GROUP_OPS overrides the per-model table below)
GROUP_OPS_SHOOT = 0  # > 0: separate budget for the shooting-node programs
MAX_EXPORTS = 64     # intermediate values exported per stage (scaled up for large models, see NodeCode.plan)
EXPORTS_PER_OPS = 40  # large models: one export per this many operations of the function (see NodeCode.plan)
_EXPORTS_PER_OPS_DEFAULT = EXPORTS_PER_OPS
_exports_per_ops = [EXPORTS_PER_OPS]     # value in force for the model being generated (per-model table below)
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


# per-model overrides of the parameters above, chosen by measurement on B200 (the planner's outcome is not monotone in
# its budget: scripts/build_variants.py + scripts/gpu_variants.sh): mpc_3dof 0.81 ms per 8192-scenario step with the
# default group budget, 0.66 ms with 1400; dual_kites 8.05 / 7.91 / 7.70 / 7.49 / 11.2 ms per 4096-instance step with 700 / 900 /
# 1200 / 1600 / 2200; with one export per 60 instead of 40 operations (328 instead of 449 exports per node) 6.89 ms
TUNED = {'mpc_3dof': {'GROUP_OPS': 1400}, 'dual_kites': {'GROUP_OPS': 1600, 'EXPORTS_PER_OPS': 60}}


class NodeCode(object):
    def __init__(self, ocp):
        self.ocp = ocp
        self.group_ops = TUNED.get(ocp.spec.name, {}).get('GROUP_OPS')
        # an experiment that changes EXPORTS_PER_OPS overrides the per-model table (as for GROUP_OPS)
        _exports_per_ops[0] = (TUNED.get(ocp.spec.name, {}).get('EXPORTS_PER_OPS', EXPORTS_PER_OPS)
                               if EXPORTS_PER_OPS == _EXPORTS_PER_OPS_DEFAULT else EXPORTS_PER_OPS)
        self.group_ops_shoot = TUNED.get(ocp.spec.name, {}).get('GROUP_OPS_SHOOT')
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

        # ---- collocation node, output list of the CUDA kernels (ocp.kcoll_out: head + own CCS range)
        self.cco = S.sym('cco')                     # C[j+1, j+1] / (h t_f)
        self.ccs = S.symvec('ccs', d)               # ccs[dl] = C[j+1, sib+1] / (h t_f), sib = (j + dl) mod d
        self.sib_x = {dl: S.symvec('sibx%d_' % dl, ocp.n_x) for dl in range(1, d)}
        self.sib_xd = {dl: S.symvec('sibxd%d_' % dl, ocp.n_x) for dl in range(1, d)}
        self.sib_z = {dl: S.symvec('sibz%d_' % dl, ocp.n_z) for dl in range(1, d)}
        b_of = {}
        for i in range(len(eq)):
            for mm in ocp.eq_split[i]['xdot']:
                b_of[(i, mm)] = j_eq[i].get(rng['xdot'][0] + mm, S.ZERO)
        b_keys = sorted(b_of)
        b_sib = {}
        for dl in range(1, d):
            mp = {}
            for i_ in range(ocp.n_x):
                mp[m.vars[rng['x'][0] + i_]] = self.sib_x[dl][i_]
                mp[m.vars[rng['xdot'][0] + i_]] = self.sib_xd[dl][i_]
            for c_ in range(ocp.n_z):
                mp[m.vars[rng['z'][0] + c_]] = self.sib_z[dl][c_]
            sub = S.substitute([S._e(b_of[key]) for key in b_keys], mp)
            for key, e in zip(b_keys, sub):
                b_sib[(dl,) + key] = e
        by_old = {item: e for item, e in zip(ocp.coll_out, self.coll_exprs)}
        outs = []
        for item in ocp.kcoll_out:
            kind = item[0]
            if kind in ('res', 'u', 'theta', 'phi', 'z') or (kind == 'Xc'):
                outs.append(by_old[item])            # same expression as in the plain emission (scaled already)
                continue
            if kind == 'O':
                _, i, mm = item
                e = S.ZERO
                if mm in ocp.eq_split[i]['x']:
                    e = e + j_eq[i].get(rng['x'][0] + mm, S.ZERO)
                if mm in ocp.eq_split[i]['xdot']:
                    e = e + b_of[(i, mm)] * self.cco
            else:
                _, dl, i, mm = item
                e = b_sib[(dl, i, mm)] * self.ccs[dl]
            outs.append(sc(S._e(e)))
        self.kcoll_exprs = outs
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
        leaf[self.cco.idx] = 'CCO'
        for dl in range(1, d):
            leaf[self.ccs[dl].idx] = 'CCS(%d)' % dl
            for i in range(ocp.n_x):
                leaf[self.sib_x[dl][i].idx] = 'XS(%d, %d)' % (dl, i)
                leaf[self.sib_xd[dl][i].idx] = 'XDS(%d, %d)' % (dl, i)
            for c in range(ocp.n_z):
                leaf[self.sib_z[dl][c].idx] = 'ZS(%d, %d)' % (dl, c)
        self.k_coll = S.substitute(self.kcoll_exprs, base)
        self.k_shoot = S.substitute(self.shoot_exprs, base)

    @staticmethod
    def plan(exprs, limit=None, max_exports=None, min_gain=None, cuts=None, forced=()):
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

        ``cuts``: output indices at which a new group may start (None: anywhere); ``forced``: indices at which one
        must start.

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
            max_exports = max(MAX_EXPORTS, n_ops // _exports_per_ops[0]) if n_ops > 8000 else MAX_EXPORTS
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
                if i > first and (i in forced or (wsum(u) > limit and (cuts is None or i in cuts))):
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
                       export=False, leaf_names=None, q0=0, chunks=None, own_init=None):
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
            prog = SCH.allocate(exprs, n_regs, lookahead, leaf_lookahead=LEAF_LOOKAHEAD, leaf_burst=LEAF_BURST)
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
        if not export and chunks is None:
            # staged outputs in chunks of ``chunk``, cooperative / per-warp flush
            chunks = [('head', c0, min(chunk, len(exprs) - c0), q0 + c0 // chunk) for c0 in range(0, len(exprs), chunk)]
        if not export:
            n_headq = sum(1 for c in chunks if c[0] == 'head')
            if n_headq:
                lines.append('GROUP_INIT(%d, %d, %d);' % (q0, n_headq, first))
            if own_init is not None:
                lines.append('OWN_INIT(%d, %d, %d, %d);' % own_init)
            chunk_of, own_idx = {}, 0
            for ci, (kind, c0, cnt, qid) in enumerate(chunks):
                for n_ in range(c0, c0 + cnt):
                    chunk_of[n_] = ci

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
                ci = chunk_of[n]
                kind, c0, cnt, qid = chunks[ci]
                if kind == 'head':
                    lines.append('OUT(%d, %s);' % (n - c0, src(it[2])))
                    if n + 1 == c0 + cnt:
                        nxt = chunks[ci + 1] if ci + 1 < len(chunks) and chunks[ci + 1][0] == 'head' else None
                        lines.append('FLUSH(%d, %d, %d, %d, %d);' % (
                            qid, first + c0, cnt, -1 if nxt is None else nxt[3], first + c0 + cnt))
                else:
                    # own CCS range: position table index = output number inside the group's own outputs
                    if n == c0:
                        lines.append('OWN_BEGIN(%d);' % qid)
                    lines.append('OUTO(%d, %s);' % (n - own_init[2] + first if own_init else n, src(it[2])))
                    if n + 1 == c0 + cnt:
                        lines.append('FLUSH_OWN(%d);' % qid)
            else:
                r, n, srcs = it[1], it[2], it[3]
                lines.append('r%d = %s;' % (r, _rhs(n.op, src(srcs[0]), src(srcs[1]) if len(srcs) > 1 else None)))
                n_stmt += 1
        stats = dict(prog.stats)
        stats.update(S.count_ops(exprs))
        stats.update(outputs=n_out, chunks=len(chunks) if chunks else 0, hot_slots=n_hot, cold_slots=n_cold,
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

    def emit_all(self, n_regs_of=None):
        """``n_regs_of``: {variant name: FP64 registers handed to the allocator} (default N_REGS everywhere)."""
        n_regs_of = n_regs_of or {}
        self._ktab = {}
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
        # ---- objective integrands (awebox_b200/ocp/objective.py): power integrand, side-slip angles and their sparse
        #      gradients with respect to the node column, plain emission (a few hundred operations)
        from .ocp.objective import NodeIntegrands
        integ = NodeIntegrands(ocp.model)
        flat, layout = integ.flat_outputs()
        cost_text, cost_stats = self.emit([S._e(e) for e in flat])
        files['node_cost.inc'] = banner + cost_text
        stats['cost'] = cost_stats
        cols = []
        for first, cc_ in layout:
            cols += [first, len(cc_)] + list(cc_)
        head += '#define AWE_COST_N_OUT %d\n#define AWE_COST_N_FUN %d\n' % (len(flat), len(layout))
        head += ('/* per integrand: first output, entries, variable indices */\n#ifdef __CUDACC__\n__device__\n#endif\n'
                 'static const int awe_cost_layout[] = {%s};\n' % ', '.join(map(str, cols)))
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
                cuts = forced = None
                if name == 'k_coll' and level == 0:
                    # groups of own outputs consist of whole CCS columns; head and own outputs do not share a group
                    n_head = ocp.n_kcoll_head
                    cuts = set(range(n_head + 1)) | set(n_head + f for f, _ in ocp.kcoll_cols)
                    forced = {n_head}
                lim = self.group_ops if GROUP_OPS == _GROUP_OPS_DEFAULT else None
                if name.startswith('k_shoot') and (GROUP_OPS_SHOOT > 0 or self.group_ops_shoot):
                    lim = GROUP_OPS_SHOOT if GROUP_OPS_SHOOT > 0 else self.group_ops_shoot
                exports, groups = self.plan(cur, limit=lim,
                                            max_exports=0 if level == MAX_STAGES - 1 else cap, cuts=cuts,
                                            forced=forced or ())
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
                    own_q = []           # own chunks: (first own output, count, first local column, columns)
                    group_own = []       # per group: (first own chunk, chunks, first own output, outputs)
                    n_head = ocp.n_kcoll_head if name == 'k_coll' else len(outs)
                    for g, (a, b) in enumerate(groups):
                        if level == 0 and a >= n_head:
                            # own CCS range: chunks of whole columns, at most OWN_CAP staged entries (constant
                            # entries of the columns included: at most 2 per column)
                            cols = [ci for ci, (f, c) in enumerate(ocp.kcoll_cols) if a <= n_head + f < b and c > 0]
                            chunks, cur_cols = [], []
                            cap = max(OWN_CAP, max(c for _, c in ocp.kcoll_cols) + 2)

                            def close():
                                if cur_cols:
                                    f0 = ocp.kcoll_cols[cur_cols[0]][0]
                                    cnt = sum(ocp.kcoll_cols[ci][1] for ci in cur_cols)
                                    chunks.append(('own', n_head + f0 - a, cnt, len(own_q)))
                                    own_q.append((f0, cnt, cur_cols[0], cur_cols[-1] - cur_cols[0] + 1))
                                    del cur_cols[:]

                            oq0 = len(own_q)
                            for ci in cols:
                                if cur_cols and (sum(ocp.kcoll_cols[x][1] + 2 for x in cur_cols) + ocp.kcoll_cols[ci][1] + 2 > cap
                                                 or ci != cur_cols[-1] + 1):
                                    close()
                                cur_cols.append(ci)
                            close()
                            assert sum(c[2] for c in chunks) == b - a, 'empty columns inside an own group'
                            group_own.append((oq0, len(own_q) - oq0, a - n_head, b - a))
                            body, st = self.emit_scheduled(g_exprs[a:b], first=a, leaf_names=leaf_names, q0=q0, chunks=chunks,
                                                           own_init=(oq0, len(own_q) - oq0, a, b - a),
                                                           n_regs=n_regs_of.get(name, N_REGS))
                        elif level == 0:
                            body, st = self.emit_scheduled(g_exprs[a:b], first=a, leaf_names=leaf_names, q0=q0,
                                                           n_regs=n_regs_of.get(name, N_REGS))
                            for c0 in range(a, b, CHUNK):
                                chunk_first.append(c0)
                                chunk_cnt.append(min(CHUNK, b - c0))
                            q0 += (b - a + CHUNK - 1) // CHUNK
                            group_own.append((len(own_q), 0, 0, 0))
                        else:
                            body, st = self.emit_scheduled(g_exprs[a:b], first=ex_index[outs[a].idx], export=True,
                                                           leaf_names=leaf_names, n_regs=n_regs_of.get(name, N_REGS))
                            assert [ex_index[n.idx] for n in outs[a:b]] == list(range(ex_index[outs[a].idx],
                                                                                      ex_index[outs[a].idx] + b - a))
                        text += 'case %d: {\n%s} break;\n' % (g, body)
                        st['first'], st['last'] = a, b
                        gstats.append(st)
                files['%s_s%d.inc' % (name, level)] = banner + text
                sstats.append(gstats)
                if level == 0:
                    chunks0 = (chunk_first, chunk_cnt)
                    own0 = (own_q, group_own)
            allg = [st for gs in sstats for st in gs]
            hot = max(st['hot_slots'] for st in allg)
            stats[name] = {'stages': sstats, 'n_exports': len(ex_nodes),
                           'ops_total': sum(flops(st) for st in allg), 'hot_slots': hot}
            up = name.upper()
            kdefs += '#define AWE_%s_EXPORTS %d\n#define AWE_%s_HOT %d\n#define AWE_%s_OPS %d\n' % (
                up, len(ex_nodes), up, hot, up, stats[name]['ops_total'])
            for level in range(N_STAGE_FILES):
                kdefs += '#define AWE_%s_S%d_GROUPS %d\n' % (up, level, len(sstats[level]))
                # relative cost of the groups (share of the SMs a group gets in the persistent node kernels):
                # operations + input loads + staged outputs
                wts = [int(st['op'] + 2 * st['leaf'] + 3 * (st['st'] + st['ld']) + (8 if level == 0 else 2) * st['outputs'])
                       for st in sstats[level]] or [0]
                kdefs += '#define AWE_%s_S%d_WEIGHTS {%s}\n' % (up, level, ', '.join(map(str, wts)))
            kdefs += '#define AWE_%s_NCHUNKS %d\n' % (up, max(1, len(chunks0[0])))
            n_head_ = ocp.n_kcoll_head if name == 'k_coll' else len(stages[0][0])
            kdefs += '#define AWE_%s_MAXQ %d\n' % (up, max([1] + [(min(b, n_head_) - a + CHUNK - 1) // CHUNK
                                                                  for a, b in stages[0][1] if a < n_head_]))
            kdefs += 'static const int awe_%s_chunk_first[] = {%s};\n' % (name, ', '.join(map(str, chunks0[0])))
            kdefs += 'static const int awe_%s_chunk_cnt[] = {%s};\n' % (name, ', '.join(map(str, chunks0[1])))
            gq, acc = [0], 0
            for a_, b_ in stages[0][1]:
                if a_ < n_head_:
                    acc += (b_ - a_ + CHUNK - 1) // CHUNK
                gq.append(acc)
            if name == 'k_coll':
                own_q, group_own = own0
                kdefs += '#define AWE_N_KCOLL_OUT %d\n#define AWE_N_KCOLL_HEAD %d\n#define AWE_N_KCOLL_OWN %d\n' % (
                    len(ocp.kcoll_out), ocp.n_kcoll_head, len(ocp.kcoll_own))
                kdefs += '#define AWE_K_COLL_NOWNQ %d\n' % max(1, len(own_q))
                rs = max([c + 2 * nc for _, c, _, nc in own_q] or [2]) + 2
                kdefs += '#define AWE_OWN_RS %d\n' % (rs + (rs & 1))
                kdefs += '#define AWE_K_COLL_MAXOWN %d\n' % max([1] + [n for _, _, _, n in group_own])
                kdefs += '#define AWE_K_COLL_MAXOWNQ %d\n' % max([1] + [n for _, n, _, _ in group_own])
                for nm_, col in (('first', 0), ('cnt', 1), ('col0', 2), ('ncol', 3)):
                    kdefs += 'static const int awe_k_coll_own_%s[] = {%s};\n' % (nm_, ', '.join(str(q[col]) for q in own_q) or '0')
                stats[name]['own_chunks'] = len(own_q)
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

    def emit_fitted(self):
        """emit_all with the register budget of the full collocation / shooting programs chosen so that the persistent
        node kernels keep as many warps per SM as possible: three quads of 168-register threads if the spill slots,
        staging rows and tables of three quads fit the shared memory of an SM, else two quads of 255-register threads
        (N_REGS_FAT values in registers: far fewer spill slots)."""
        files, stats = self.emit_all()
        fat = {}
        for name in ('k_coll', 'k_shoot'):
            if quad_smem_bytes(name, stats[name]['hot_slots'], files['model_defs.h']) * 3 > SMEM_PER_SM - 64:
                fat[name] = N_REGS_FAT
        if fat:
            files, stats = self.emit_all(n_regs_of=fat)
        stats['n_regs'] = {name: fat.get(name, N_REGS) for name in ('k_coll', 'k_shoot')}
        return files, stats

    def write(self, out_dir):
        import os
        os.makedirs(out_dir, exist_ok=True)
        files, stats = self.emit_fitted()
        for name, text in files.items():
            path = os.path.join(out_dir, name)
            old = open(path).read() if os.path.exists(path) else None
            if old != text:
                with open(path, 'w') as f:
                    f.write(text)
        return stats


def quad_smem_bytes(name, hot, defs):
    """shared memory of one quad of the persistent node kernel (mirrors node_smem_bytes in awegpu_kernels.cu)."""
    import re

    def macro(key, default=0):
        m_ = re.search(r'#define %s (\d+)' % key, defs)
        return int(m_.group(1)) if m_ else default

    stg = 4 * CHUNK * 33
    if name == 'k_shoot':
        return 8 * (hot * 128 + stg)
    n_x, n_z = macro('AWE_N_X'), macro('AWE_N_Z')
    own_tab = 16 * macro('AWE_K_COLL_MAXOWNQ', 1) + ((2 * macro('AWE_K_COLL_MAXOWN', 1) + 15) // 16) * 16 + ((4 * (n_x + n_z) + 15) // 16) * 16
    head = stg + 8 * 33 + macro('AWE_K_COLL_MAXQ', 1) * (4 * CHUNK + 32)
    own = 128 * (macro('AWE_OWN_RS', 2) | 1)
    return 8 * (hot * 128 + (4 * own_tab + 7) // 8 + max(head, own))


def flops(stats):
    """algorithmic flop count of an emitted function: every listed operation counts once."""
    return sum(v for k, v in stats.items() if k in S._NAMES.values())
