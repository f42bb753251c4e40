"""Second-order hook: the Hessian of the constraint part of the Lagrangian,
``H(w, p, lam) = d^2/dw^2 (lam' g(w, p))``, upper triangle in CCS, for a batch of OCP instances.

What it replaces: the ``nlp_hess_l`` function CasADi derives inside ``cas.nlpsol`` when the reference solves with the
exact Hessian (``awebox/opts/default.py:323`` ``hessian_approximation = False``; the Lagrangian as the reference itself
forms it: ``awebox/tools/debug_operations.py:567-596``).  ``g`` is assembled from the node functions as in
``awebox/ocp/constraints.py:210-373``; continuity (``collocation.py:319-336``), periodicity
(``operation.py:245-267``), ``t_f`` bounds and the MPC initial conditions are affine in ``w`` and do not contribute.

Decomposition (host side, build time):

* collocation node (k, j): ``r(x, xdot, u, z, theta, phi)`` with ``xdot = (sum_r C[r, j+1] X_{k,r}) / h / t_f``
  (``collocation.py:202-258``).  With the *extended node variables* ``v = (x, xsum, u, z, theta, phi)``,
  ``xsum = sum_r C[r, j+1] X_{k,r}``, the map ``w -> v`` is **linear** (``t_f`` is an entry of ``theta``), so
  ``d^2/dw^2 (lam' r) = M' [d^2/dv^2 (lam' r(x, xsum n_k / t_f, ...))] M`` with a constant ``M``: all second-order
  terms of the ``1/t_f`` chain are inside the node Hessian, which ``sym.hessian_triu`` derives from the model's
  expression graph (reverse-mode gradient, forward-mode Jacobian of it).
* shooting node k: equality subset + path inequalities over ``(x_k, xdot_k, u_k, z_k, theta, phi)``, ``M`` = selection.

``LagrangianHessian(ocp, objective=True)`` is the complete ``hess_lag`` of nlpsol, ``lam_f f + lam_g' g`` (see the
constructor); the libraries are generated with it for power-cycle trajectories (``has_objective``).

The CUDA path evaluates the node Hessians (generated programs ``node_hess_coll.inc`` / ``node_hess_shoot.inc``, one
thread per node and instance) into a lane-contiguous buffer and a second kernel forms ``H = A hv`` with the constant
sparse matrix ``A`` (``asm_ptr / asm_idx / asm_coef``: for every CCS slot its node values and coefficients, fixed
summation order: deterministic, no atomics).  ``evaluate`` is the numpy restatement of the same two steps (host side,
tests and tables only; the product path is the kernel).
"""
import numpy as np

from .. import sym as S


def has_objective(ocp):
    """the generated programs carry the objective part for power-cycle trajectories (the MPC trial's tracking cost belongs
    to awebox/pmpc.py and has a constant, diagonal Hessian)"""
    return ocp.traj_type == 'power_cycle'


HESS_SECTIONS = 16        # programs a node Hessian is cut into (whole columns each); see LagrangianHessian.emit
HESS_SYNC_EVERY = 0       # > 0: a CTA-wide re-convergence point (HSYNC) every this many statements and after every section:
                          # the warps that run the same program stay inside one instruction-cache window


class LagrangianHessian(object):
    def __init__(self, ocp, objective=False):
        """objective = True: the complete ``hess_lag`` of nlpsol, ``lam_f f + lam_g' g`` (power-cycle trajectories: the
        objective of awebox/ocp/objective.py as restated in awebox_b200/ocp/objective.py).  Every term of ``f`` but the
        time cost is a sum over the collocation nodes of a function of the node's extended variables, the homotopy
        parameter psi (an entry of phi) and the time period T = sum_i c_i t_f,i (one more extended variable, linear in
        w): it is added to the node's scalar with the multiplier lam_f.  The time cost c (T - T_ref)^2 rides on the
        first node (leaf ``first_node`` = 1 there, 0 elsewhere)."""
        self.ocp = ocp
        self.objective = bool(objective)
        m, d, n_k = ocp.model, ocp.d, ocp.n_k
        rng = ocp.rng
        n_x = ocp.n_x
        scale = ocp.constraint_scale
        eq, ineq = m.eq_exprs(), m.ineq_exprs()
        local_in = list(ocp.local_in)                     # node variables + phi
        self.n_local = len(local_in)
        # ---- collocation node over the extended variables
        self.xsum = S.symvec('xsum', n_x)
        tf = m.vars[ocp.tf_local]
        self.nk_sym = S.sym('n_k')                        # run-time leaf: one library serves every n_k
        inv_h_tf = self.nk_sym / tf
        xdot_sym = m.vars[rng['xdot'][0]:rng['xdot'][1]]
        self.xdot_of = [self.xsum[i] * inv_h_tf for i in range(n_x)]
        eq_ext = S.substitute(eq, {xdot_sym[i]: self.xdot_of[i] for i in range(n_x)})
        self.ext_in = list(local_in)
        for i in range(n_x):
            self.ext_in[rng['xdot'][0] + i] = self.xsum[i]
        self.lam_c = S.symvec('lamc', len(eq))
        phi_c = S.dot(self.lam_c, eq_ext)
        if scale != 1.0:
            phi_c = scale * phi_c
        if self.objective:
            phi_c = phi_c + self._objective_term()
        hc = S.hessian_triu(phi_c, self.ext_in)
        self.coll_keys = sorted(hc)
        self.coll_exprs = [hc[key] for key in self.coll_keys]
        # ---- shooting node
        self.shoot_rows = list(ocp.shoot_rows)
        self.lam_s = S.symvec('lams', len(self.shoot_rows))
        terms = [eq[i] if kind == 'eq' else ineq[i] for kind, i in self.shoot_rows]
        phi_s = S.dot(self.lam_s, terms)
        if scale != 1.0:
            phi_s = scale * phi_s
        hs = S.hessian_triu(phi_s, local_in)
        self.shoot_keys = sorted(hs)
        self.shoot_exprs = [hs[key] for key in self.shoot_keys]
        self.local_in = local_in
        self.n_ec, self.n_es = len(self.coll_keys), len(self.shoot_keys)
        self.n_hv = n_k * (d * self.n_ec + self.n_es)          # node values per instance
        self._assemble()

    # ------------------------------------------------------------------ objective part of a collocation node
    def _objective_term(self):
        from .objective import Objective, COST_NAMES
        ocp = self.ocp
        m, rng, n_x = ocp.model, ocp.rng, ocp.n_x
        obj = self.obj = Objective(ocp)
        n_var = ocp.n_var
        self.sig, self.T_sym, self.Tref_sym = S.sym('lam_f'), S.sym('T_period'), S.sym('T_period_ref')
        self.qj, self.first = S.sym('q_j'), S.sym('first_node')
        self.ref, self.pw, self.cf = S.symvec('ref', n_var), S.symvec('pw', n_var), S.symvec('cf', len(COST_NAMES))
        self.T_terms = obj.time_period_terms()
        self.ext_in = self.ext_in + [self.T_sym]             # one more extended variable (index n_local)
        psi = m.phi('psi')
        tf = m.vars[ocp.tf_local]
        # the normalisations are proportional to n_k (awebox/opts/funcs.py:177-182): n_k is a run-time leaf of the programs
        inv_norm = {c: (float(ocp.n_k) / obj.norm[c]) / self.nk_sym for c in obj.norm}
        reg = {}
        for i, c in enumerate(obj.cat_of):
            if c is None:
                continue
            if rng['xdot'][0] <= i < rng['xdot'][1]:
                diff = self.xdot_of[i - rng['xdot'][0]]                   # the reference of xdot is zero
            else:
                diff = m.vars[i] - self.ref[i]
            wgt = self.pw[i] * self.cf[COST_NAMES.index(c)] * inv_norm[c]
            reg[c] = reg.get(c, S.ZERO) + wgt * diff * diff
        cost = psi * reg.get('tracking', S.ZERO)
        for c, e in reg.items():
            if c != 'tracking':
                cost = cost + e
        xdot_sym = m.vars[rng['xdot'][0]:rng['xdot'][1]]
        sub = S.substitute([obj.integrands.power] + list(obj.integrands.betas),
                           {xdot_sym[i]: self.xdot_of[i] for i in range(n_x)})
        power, betas = sub[0], sub[1:]
        if betas:
            cost = cost + (self.cf[COST_NAMES.index('beta')] * inv_norm['beta']) * S.dot(betas, betas)
        # power cost: -c_P / T * sum h q_j p,  h = t_f / n_k
        cost = cost + (S.ONE - psi) * (-self.cf[COST_NAMES.index('power')]) * (tf / self.nk_sym) * power / self.T_sym
        dT = self.T_sym - self.Tref_sym
        return self.sig * (self.qj * cost + self.first * self.cf[COST_NAMES.index('t_f')] * dT * dT)

    # ------------------------------------------------------------------ w -> node variables (constant, linear)
    def _maps(self, k, j):
        """[(V index, coefficient)] per node variable; j = None: shooting node k"""
        ocp = self.ocp
        rng, d = ocp.rng, ocp.d
        C = ocp.coll.coeff_collocation
        out = [None] * (self.n_local + 1)
        if self.objective:
            out[self.n_local] = [(int(i), float(c)) for i, c in self.T_terms]
        for i in range(ocp.n_x):
            if j is None:
                out[rng['x'][0] + i] = [(ocp.v_x[k] + i, 1.0)]
                out[rng['xdot'][0] + i] = [(ocp.v_xdot[k] + i, 1.0)]
            else:
                out[rng['x'][0] + i] = [(ocp.x_base(k, j + 1) + i, 1.0)]
                out[rng['xdot'][0] + i] = [(ocp.x_base(k, r) + i, float(C[r, j + 1])) for r in range(d + 1)
                                           if C[r, j + 1] != 0.0]
        for c in range(ocp.n_u):
            out[rng['u'][0] + c] = [(ocp.v_u[k] + c, 1.0)]
        for c in range(ocp.n_z):
            out[rng['z'][0] + c] = [((ocp.v_z[k] if j is None else ocp.v_cz[k][j]) + c, 1.0)]
        for t in range(ocp.n_theta):
            out[rng['theta'][0] + t] = [(int(ocp.theta_idx[k, t]), 1.0)]
        for c in range(ocp.n_phi):
            out[rng['phi'][0] + c] = [(ocp.v_phi + c, 1.0)]
        return out

    def node_offset(self, k, j):
        """first node value of node (k, j) (j = None: shooting node) in the per-instance vector ``hv``"""
        per_k = self.ocp.d * self.n_ec + self.n_es
        return k * per_k + (self.ocp.d * self.n_ec if j is None else j * self.n_ec)

    def _assemble(self):
        ocp = self.ocp
        n_v = ocp.n_v
        ti, tj, tv, tc = [], [], [], []          # triplets: row, column (row <= column), node value, coefficient
        for k in range(ocp.n_k):
            for j in list(range(ocp.d)) + [None]:
                maps = self._maps(k, j)
                keys = self.shoot_keys if j is None else self.coll_keys
                base = self.node_offset(k, j)
                for e, (a, b) in enumerate(keys):
                    for p, cp in maps[a]:
                        for q, cq in maps[b]:
                            c = cp * cq
                            if a == b:
                                if p > q:
                                    continue            # the ordered pair (q, p) of the same loop covers it
                            elif p == q:
                                c *= 2.0                # (a, b) and (b, a) land on the same diagonal entry
                            ti.append(min(p, q))
                            tj.append(max(p, q))
                            tv.append(base + e)
                            tc.append(c)
        ti, tj, tv, tc = np.array(ti), np.array(tj), np.array(tv), np.array(tc)
        # merge duplicates (same slot, same node value)
        key = (tj * n_v + ti) * self.n_hv + tv
        order = np.argsort(key, kind='stable')
        key, ti, tj, tv, tc = key[order], ti[order], tj[order], tv[order], tc[order]
        first = np.concatenate([[True], key[1:] != key[:-1]])
        grp = np.cumsum(first) - 1
        coef = np.zeros(grp[-1] + 1)
        np.add.at(coef, grp, tc)
        ti, tj, tv = ti[first], tj[first], tv[first]
        # CCS of the upper triangle: columns ascending, rows ascending inside a column
        slot_key = tj * n_v + ti
        sfirst = np.concatenate([[True], slot_key[1:] != slot_key[:-1]])
        slot = np.cumsum(sfirst) - 1
        self.nnz = int(slot[-1]) + 1
        self.row = ti[sfirst].astype(np.int64)
        cols = tj[sfirst]
        self.colind = np.zeros(n_v + 1, dtype=np.int64)
        np.add.at(self.colind, cols + 1, 1)
        self.colind = np.cumsum(self.colind)
        self.asm_ptr = np.zeros(self.nnz + 1, dtype=np.int64)
        np.add.at(self.asm_ptr, slot + 1, 1)
        self.asm_ptr = np.cumsum(self.asm_ptr)
        self.asm_idx = tv.astype(np.int64)
        self.asm_coef = coef

    # ------------------------------------------------------------------ numpy restatement (host; tests / tables)
    def node_inputs(self, w, p):
        """-> (coll [n_local + n_theta0? no: ext inputs, n_k*d], shoot [local inputs, n_k]) gathered from w"""
        ocp = self.ocp
        rng, d, n_k = ocp.rng, ocp.d, ocp.n_k
        C = ocp.coll.coeff_collocation
        xc = np.zeros((self.n_local, n_k * d))
        xs = np.zeros((self.n_local, n_k))
        for k in range(n_k):
            th = w[ocp.theta_idx[k]]
            ph = w[ocp.v_phi:ocp.v_phi + ocp.n_phi]
            X = [w[ocp.x_base(k, r):ocp.x_base(k, r) + ocp.n_x] for r in range(d + 1)]
            for j in range(d):
                col = xc[:, k * d + j]
                col[rng['x'][0]:rng['x'][1]] = X[j + 1]
                col[rng['xdot'][0]:rng['xdot'][1]] = sum(C[r, j + 1] * X[r] for r in range(d + 1))
                col[rng['u'][0]:rng['u'][1]] = w[ocp.v_u[k]:ocp.v_u[k] + ocp.n_u]
                col[rng['z'][0]:rng['z'][1]] = w[ocp.v_cz[k][j]:ocp.v_cz[k][j] + ocp.n_z]
                col[rng['theta'][0]:rng['theta'][1]] = th
                col[rng['phi'][0]:rng['phi'][1]] = ph
            col = xs[:, k]
            col[rng['x'][0]:rng['x'][1]] = X[0]
            col[rng['xdot'][0]:rng['xdot'][1]] = w[ocp.v_xdot[k]:ocp.v_xdot[k] + ocp.n_x]
            col[rng['u'][0]:rng['u'][1]] = w[ocp.v_u[k]:ocp.v_u[k] + ocp.n_u]
            col[rng['z'][0]:rng['z'][1]] = w[ocp.v_z[k]:ocp.v_z[k] + ocp.n_z]
            col[rng['theta'][0]:rng['theta'][1]] = th
            col[rng['phi'][0]:rng['phi'][1]] = ph
        return xc, xs

    def node_multipliers(self, lam):
        ocp = self.ocp
        d, n_k = ocp.d, ocp.n_k
        lc = np.zeros((ocp.n_eq, n_k * d))
        ls = np.zeros((len(self.shoot_rows), n_k))
        for k in range(n_k):
            for j in range(d):
                lc[:, k * d + j] = lam[ocp.row_coll[k][j]:ocp.row_coll[k][j] + ocp.n_eq]
            ls[:ocp.n_shoot_eq, k] = lam[ocp.row_shoot[k]:ocp.row_shoot[k] + ocp.n_shoot_eq]
            ls[ocp.n_shoot_eq:, k] = lam[ocp.row_path[k]:ocp.row_path[k] + ocp.n_ineq]
        return lc, ls

    def objective_inputs(self, w, p, lam_f):
        """leaves of the objective part per collocation node: T, ref, pw, cf, q_j, first_node, T_ref, lam_f"""
        ocp, obj = self.ocp, self.obj
        d, n_k, n_var = ocp.d, ocp.n_k, ocp.n_var
        n_c = n_k * d
        ref_v = p[:ocp.n_v]
        refc, _ = self.node_inputs(ref_v, p)
        T = sum(c * w[i] for i, c in self.T_terms)
        T_ref = sum(c * ref_v[i] for i, c in self.T_terms)
        pw = p[obj.p_weights:obj.p_weights + n_var]
        cf = p[obj.p_cost:obj.p_cost + obj.n_cost]
        first = np.zeros((1, n_c))
        first[0, 0] = 1.0
        return np.concatenate([np.full((1, n_c), T), refc[:n_var], np.repeat(pw[:, None], n_c, 1), np.repeat(cf[:, None], n_c, 1),
                               np.tile(obj.quad, n_k)[None, :], first, np.full((1, n_c), T_ref), np.full((1, n_c), float(lam_f))])

    def node_values(self, w, p, lam, lam_f=0.0):
        """per-instance node Hessian values ``hv`` [n_hv]"""
        ocp = self.ocp
        m = ocp.model
        theta0 = np.asarray(p[ocp.p_theta0:ocp.p_theta0 + ocp.n_theta0], dtype=float)
        xc, xs = self.node_inputs(w, p)
        lc, ls = self.node_multipliers(lam)
        par = list(m.params[:m.n_theta0])
        n_c, n_s = xc.shape[1], xs.shape[1]
        leaves = self.ext_in[:self.n_local] + par + list(self.lam_c) + [self.nk_sym]
        values = [xc, np.repeat(theta0[:, None], n_c, 1), lc, np.full((1, n_c), float(ocp.n_k))]
        if self.objective:
            leaves += [self.T_sym] + list(self.ref) + list(self.pw) + list(self.cf) + [self.qj, self.first, self.Tref_sym, self.sig]
            values.append(self.objective_inputs(w, p, lam_f))
        vc = S.evaluate(self.coll_exprs, leaves, np.concatenate(values))
        vs = S.evaluate(self.shoot_exprs, self.local_in + par + list(self.lam_s),
                        np.concatenate([xs, np.repeat(theta0[:, None], n_s, 1), ls]))
        hv = np.zeros(self.n_hv)
        for k in range(ocp.n_k):
            for j in range(ocp.d):
                o = self.node_offset(k, j)
                hv[o:o + self.n_ec] = vc[:, k * ocp.d + j]
            o = self.node_offset(k, None)
            hv[o:o + self.n_es] = vs[:, k]
        return hv

    def evaluate(self, w, p, lam, lam_f=0.0):
        """non-zeros of triu(H) in CCS order, one instance"""
        hv = self.node_values(np.asarray(w, dtype=float), np.asarray(p, dtype=float), np.asarray(lam, dtype=float), lam_f)
        prod = self.asm_coef * hv[self.asm_idx]
        return np.add.reduceat(prod, self.asm_ptr[:-1])

    # ------------------------------------------------------------------ generated node programs (CUDA path)
    def emit(self, exprs, lam, xsum=None, keys=None, sections=None):
        """straight-line code ``HOUT(n, value)`` over the leaves ``VIN(i)`` (node variables; ``XSIN(i)`` in place of
        xdot for a collocation node), ``PIN(i)`` (theta0 | phi), ``LIN(i)`` (multipliers), ``HNK`` (n_k).

        The entries are cut into ``sections`` groups of whole Hessian columns, each emitted as a program of its own
        (a C scope; what two sections share is computed twice).  One program for all entries keeps ~750 values alive
        at its peak (the adjoints of a reverse sweep), far beyond the register file: ptxas spilled 8.8 KB per thread,
        more than the L2 holds for the resident threads, and the kernel ran at the speed of DRAM.  Per section the peak
        is 20 - 480 values for 1.3x the operations."""
        from ..codegen import _rhs, _fmt
        m = self.ocp.model
        leaf = {}
        for i, v in enumerate(m.vars):
            leaf[v.idx] = 'VIN(%d)' % i
        for i, v in enumerate(m.params):
            leaf[v.idx] = 'PIN(%d)' % i
        for i, v in enumerate(lam):
            leaf[v.idx] = 'LIN(%d)' % i
        if xsum is not None:
            for i, v in enumerate(xsum):
                leaf[v.idx] = 'XSIN(%d)' % i
        leaf[self.nk_sym.idx] = 'HNK'
        if self.objective:
            for sym_, txt in ((self.sig, 'SIGF'), (self.T_sym, 'TPER'), (self.Tref_sym, 'TREF'), (self.qj, 'QJ'), (self.first, 'W0')):
                leaf[sym_.idx] = txt
            for i in range(len(self.ref)):
                leaf[self.ref[i].idx] = 'REF(%d)' % i
                leaf[self.pw[i].idx] = 'PW(%d)' % i
            for i in range(len(self.cf)):
                leaf[self.cf[i].idx] = 'CF(%d)' % i
        sections = HESS_SECTIONS if sections is None else sections
        groups = [list(range(len(exprs)))]
        if keys is not None and sections > 1:
            cols = sorted(set(b for _, b in keys))
            per = (len(cols) + sections - 1) // sections
            groups = []
            for g in range(0, len(cols), per):
                cs = set(cols[g:g + per])
                groups.append([n for n, (_, b) in enumerate(keys) if b in cs])
        lines, n_ops = [], 0
        import os
        sync_every = int(os.environ.get('AWE_HESS_SYNC_EVERY', HESS_SYNC_EVERY))
        since = 0
        for grp in groups:
            names = dict(leaf)
            n_tmp = 0
            out_of = {}
            for k in grp:
                out_of.setdefault(exprs[k].idx, []).append(k)

            def ref(n):
                return _fmt(n.val) if n.op == S.CONST else names[n.idx]
            lines.append('{')
            for n in S.topo([exprs[k] for k in grp]):
                if n.op == S.SYM:
                    if n.idx not in names:
                        raise ValueError('free symbol %s in the node Hessian' % n.name)
                elif n.op != S.CONST:
                    names[n.idx] = 't%d' % n_tmp
                    lines.append('const double t%d = %s;' % (n_tmp, _rhs(n.op, ref(n.a), ref(n.b) if n.b is not None else None)))
                    n_tmp += 1
                    since += 1
                    if sync_every > 0 and since >= sync_every:
                        lines.append('HSYNC')
                        since = 0
                for k in out_of.get(n.idx, []):
                    lines.append('HOUT(%d, %s);' % (k, ref(n)))
            lines.append('}')
            if sync_every > 0:
                lines.append('HSYNC')
                since = 0
            n_ops += n_tmp
        return '\n'.join(lines) + '\n', n_ops

    def write(self, out_dir):
        """node_hess_coll.inc, node_hess_shoot.inc, hess_defs.h -> out_dir; returns statistics"""
        import os
        coll, n_c = self.emit(self.coll_exprs, self.lam_c, self.xsum, keys=self.coll_keys)
        shoot, n_s = self.emit(self.shoot_exprs, self.lam_s, keys=self.shoot_keys)
        defs = ('// generated by awebox_b200/ocp/hessian.py -- do not edit\n'
                '#define AWE_H_NEC %d\n#define AWE_H_NES %d\n#define AWE_H_NLAM_S %d\n#define AWE_H_OBJ %d\n'
                % (self.n_ec, self.n_es, len(self.shoot_rows), 1 if self.objective else 0))
        for name, text in (('node_hess_coll.inc', coll), ('node_hess_shoot.inc', shoot), ('hess_defs.h', defs)):
            path = os.path.join(out_dir, name)
            if not os.path.exists(path) or open(path).read() != text:
                with open(path, 'w') as f:
                    f.write(text)
        return {'hess_coll_ops': n_c, 'hess_shoot_ops': n_s, 'hess_coll_entries': self.n_ec, 'hess_shoot_entries': self.n_es}
