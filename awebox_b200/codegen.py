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
import numpy as np

from . import sym as S


def _fmt(v):
    if v != v:
        return '(0.0/0.0)'
    if v in (float('inf'), float('-inf')):
        return '(1.0/0.0)' if v > 0 else '(-1.0/0.0)'
    r = repr(float(v))
    if 'e' not in r and '.' not in r:
        r += '.0'
    return r


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
        self._names = {}
        for i, v in enumerate(m.vars):
            self._names[v.idx] = 'VIN(%d)' % i
        for i, v in enumerate(m.params):
            self._names[v.idx] = 'PIN(%d)' % i
        for r in range(d + 1):
            self._names[self.cc[r].idx] = 'CCIN(%d)' % r
            self._names[self.own[r].idx] = 'OWNIN(%d)' % r
        self._names[self.itf.idx] = 'ITF'

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
                a = ref(n.a)
                b = ref(n.b) if n.b is not None else None
                op = n.op
                if op == S.ADD: rhs = '%s + %s' % (a, b)
                elif op == S.SUB: rhs = '%s - %s' % (a, b)
                elif op == S.MUL: rhs = '%s * %s' % (a, b)
                elif op == S.DIV: rhs = '%s / %s' % (a, b)
                elif op == S.NEG: rhs = '-%s' % a
                elif op == S.SQRT: rhs = 'sqrt(%s)' % a
                elif op == S.RSQRT: rhs = 'RSQRT(%s)' % a
                elif op in (S.POWC, S.POW): rhs = 'pow(%s, %s)' % (a, b)
                elif op == S.ATAN2: rhs = 'atan2(%s, %s)' % (a, b)
                elif op == S.SIGN: rhs = 'SIGN(%s)' % a
                else:
                    rhs = '%s(%s)' % ({S.SIN: 'sin', S.COS: 'cos', S.TAN: 'tan', S.ASIN: 'asin', S.ACOS: 'acos',
                                       S.ATAN: 'atan', S.EXP: 'exp', S.LOG: 'log', S.FABS: 'fabs'}[op], a)
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
        return files, {'coll': cs, 'shoot': ss}

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
    return sum(v for k, v in stats.items() if k not in ('statements', 'outputs'))
