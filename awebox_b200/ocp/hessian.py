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

The CUDA path evaluates the node Hessians (generated programs ``node_hess_coll.inc`` / ``node_hess_shoot.inc``, one
thread per node and instance) into a lane-contiguous buffer and a second kernel forms ``H = A hv`` with the constant
sparse matrix ``A`` (``asm_ptr / asm_idx / asm_coef``: for every CCS slot its node values and coefficients, fixed
summation order: deterministic, no atomics).  ``evaluate`` is the numpy restatement of the same two steps (host side,
tests and tables only; the product path is the kernel).
"""
import numpy as np

from .. import sym as S


class LagrangianHessian(object):
    def __init__(self, ocp):
        self.ocp = ocp
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

    # ------------------------------------------------------------------ w -> node variables (constant, linear)
    def _maps(self, k, j):
        """[(V index, coefficient)] per node variable; j = None: shooting node k"""
        ocp = self.ocp
        rng, d = ocp.rng, ocp.d
        C = ocp.coll.coeff_collocation
        out = [None] * self.n_local
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

    def node_values(self, w, p, lam):
        """per-instance node Hessian values ``hv`` [n_hv]"""
        ocp = self.ocp
        m = ocp.model
        theta0 = np.asarray(p[ocp.p_theta0:ocp.p_theta0 + ocp.n_theta0], dtype=float)
        xc, xs = self.node_inputs(w, p)
        lc, ls = self.node_multipliers(lam)
        par = list(m.params[:m.n_theta0])
        n_c, n_s = xc.shape[1], xs.shape[1]
        vc = S.evaluate(self.coll_exprs, self.ext_in + par + list(self.lam_c) + [self.nk_sym],
                        np.concatenate([xc, np.repeat(theta0[:, None], n_c, 1), lc, np.full((1, n_c), float(ocp.n_k))]))
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

    def evaluate(self, w, p, lam):
        """non-zeros of triu(H) in CCS order, one instance"""
        hv = self.node_values(np.asarray(w, dtype=float), np.asarray(p, dtype=float), np.asarray(lam, dtype=float))
        prod = self.asm_coef * hv[self.asm_idx]
        return np.add.reduceat(prod, self.asm_ptr[:-1])

    # ------------------------------------------------------------------ generated node programs (CUDA path)
    def emit(self, exprs, lam, xsum=None):
        """straight-line code ``HOUT(n, value)`` over the leaves ``VIN(i)`` (node variables; ``XSIN(i)`` in place of
        xdot for a collocation node), ``PIN(i)`` (theta0 | phi), ``LIN(i)`` (multipliers), ``HNK`` (n_k)."""
        from ..codegen import _rhs, _fmt
        m = self.ocp.model
        names = {}
        for i, v in enumerate(m.vars):
            names[v.idx] = 'VIN(%d)' % i
        for i, v in enumerate(m.params):
            names[v.idx] = 'PIN(%d)' % i
        for i, v in enumerate(lam):
            names[v.idx] = 'LIN(%d)' % i
        if xsum is not None:
            for i, v in enumerate(xsum):
                names[v.idx] = 'XSIN(%d)' % i
        names[self.nk_sym.idx] = 'HNK'
        lines, n_tmp = [], 0
        out_of = {}
        for k, e in enumerate(exprs):
            out_of.setdefault(e.idx, []).append(k)

        def ref(n):
            return _fmt(n.val) if n.op == S.CONST else names[n.idx]
        for n in S.topo(exprs):
            if n.op == S.SYM:
                if n.idx not in names:
                    raise ValueError('free symbol %s in the node Hessian' % n.name)
            elif n.op != S.CONST:
                names[n.idx] = 't%d' % n_tmp
                lines.append('const double t%d = %s;' % (n_tmp, _rhs(n.op, ref(n.a), ref(n.b) if n.b is not None else None)))
                n_tmp += 1
            for k in out_of.get(n.idx, []):
                lines.append('HOUT(%d, %s);' % (k, ref(n)))
        return '\n'.join(lines) + '\n', n_tmp

    def write(self, out_dir):
        """node_hess_coll.inc, node_hess_shoot.inc, hess_defs.h -> out_dir; returns statistics"""
        import os
        coll, n_c = self.emit(self.coll_exprs, self.lam_c, self.xsum)
        shoot, n_s = self.emit(self.shoot_exprs, self.lam_s)
        defs = ('// generated by awebox_b200/ocp/hessian.py -- do not edit\n'
                '#define AWE_H_NEC %d\n#define AWE_H_NES %d\n#define AWE_H_NLAM_S %d\n'
                % (self.n_ec, self.n_es, len(self.shoot_rows)))
        for name, text in (('node_hess_coll.inc', coll), ('node_hess_shoot.inc', shoot), ('hess_defs.h', defs)):
            path = os.path.join(out_dir, name)
            if not os.path.exists(path) or open(path).read() != text:
                with open(path, 'w') as f:
                    f.write(text)
        return {'hess_coll_ops': n_c, 'hess_shoot_ops': n_s, 'hess_coll_entries': self.n_ec, 'hess_shoot_entries': self.n_es}
