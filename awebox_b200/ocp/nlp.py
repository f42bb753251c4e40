"""OCP assembly: layout of the decision vector ``w = V.cat`` and parameter vector ``p = P.cat``, row
order of ``g``, structural CCS pattern of ``jac_g`` and the scatter tables that take node-local
Jacobian entries to CCS slots.

Follows (direct collocation, Radau/Legendre, zero-order-hold controls):
  awebox/ocp/var_struct.py:39-110           V = [theta, phi, xi, (x[nk+1], u[nk], xdot[nk], z[nk], coll_var[nk,d]{x,z})]
  awebox/ocp/discretization.py:129-179      P = [p{ref(|V|), weights(|vars|)}, cost(20), theta0]
  awebox/tools/struct_operations.py:69-116, 230-302, 363-471   node gathers, t_f selection
  awebox/ocp/collocation.py:202-258, 319-336                    xdot from the polynomial, continuity
  awebox/ocp/constraints.py:210-373         rows: per k  shooting(eq) path(ineq) {collocation}_d continuity
  awebox/ocp/operation.py:245-267           periodicity rows  x(0) - x(coll nk-1, d-1)
  awebox/ocp/constraints.py:147-168         t_f bound rows (single_reelout)
"""
import numpy as np

from .. import sym as S
from ..mdl.model import NodeModel, VAR_TYPES
from .collocation import Collocation

N_COST = 20          # entries of the cost struct, awebox/ocp/discretization.py:129-154


class OcpProblem(object):
    def __init__(self, spec, node_model=None, n_k=None):
        self.spec = spec
        self.model = node_model if node_model is not None else NodeModel(spec)
        m = self.model
        nlp = spec.nlp
        if nlp.get('discretization', 'direct_collocation') != 'direct_collocation':
            raise NotImplementedError('only direct collocation is supported')
        if nlp.get('collocation.u_param', 'zoh') != 'zoh':
            raise NotImplementedError("only u_param='zoh' is supported")
        self.n_k = int(nlp['n_k'] if n_k is None else n_k)
        self.d = int(nlp['collocation.d'])
        self.coll = Collocation(self.n_k, self.d, nlp.get('collocation.scheme', 'radau'))
        self.phase_fix = nlp.get('phase_fix', 'simple')
        self.single_reelout = (self.phase_fix == 'single_reelout')
        # operation conditions (awebox/ocp/operation.py:52-84, tools/performance_operations.py:63-69)
        self.traj_type = nlp.get('trajectory.type', 'power_cycle')
        if self.traj_type not in ('power_cycle', 'mpc'):
            raise NotImplementedError('trajectory type %r is not supported' % self.traj_type)
        self.periodic = self.traj_type == 'power_cycle'
        self.initial_conditions = self.traj_type == 'mpc'
        if self.traj_type == 'mpc' and nlp.get('mpc.terminal_point_constr', False):
            raise NotImplementedError('mpc terminal point constraint')
        self.constraint_scale = float(nlp.get('constraint_scale', 1.0))
        self.n_x, self.n_u, self.n_z, self.n_theta = m.n_x, m.n_u, m.n_z, m.n_theta
        self.n_var, self.n_phi, self.n_theta0 = m.n_var, len(spec.phi_names), m.n_theta0
        self._layout()
        self._local_patterns()
        self._rows()
        self._pattern()

    # ------------------------------------------------------------------ V / P layout
    def _layout(self):
        m, n_k, d = self.model, self.n_k, self.d
        off = 0
        self.theta_names = m.names('theta')
        self.v_theta = {}
        for nm in self.theta_names:
            sz = 2 if (nm == 't_f' and self.single_reelout) else 1
            self.v_theta[nm] = (off, sz)
            off += sz
        self.n_v_theta = off
        self.v_phi = off
        off += self.n_phi
        self.v_xi = off
        off += 2
        self.v_x, self.v_u, self.v_xdot, self.v_z = [], [], [], []
        self.v_cx = [[None] * d for _ in range(n_k)]
        self.v_cz = [[None] * d for _ in range(n_k)]
        for k in range(n_k + 1):
            self.v_x.append(off)
            off += self.n_x
            if k < n_k:
                self.v_u.append(off)
                off += self.n_u
                self.v_xdot.append(off)
                off += self.n_x
                self.v_z.append(off)
                off += self.n_z
                for j in range(d):
                    self.v_cx[k][j] = off
                    off += self.n_x
                    self.v_cz[k][j] = off
                    off += self.n_z
        self.n_v = off
        self.p_theta0 = self.n_v + self.n_var + N_COST
        self.n_p = self.p_theta0 + self.n_theta0
        # theta gather per interval (get_V_theta / calculate_tf)
        switch = int(round(n_k * float(self.spec.nlp.get('phase_fix_reelout', 0.7))))
        self.theta_idx = np.zeros((n_k, self.n_theta), dtype=np.int32)
        self.tf_idx = np.zeros(n_k, dtype=np.int32)
        for k in range(n_k):
            for t, nm in enumerate(self.theta_names):
                o, sz = self.v_theta[nm]
                self.theta_idx[k, t] = o + (1 if (sz == 2 and k >= switch) else 0)
            self.tf_idx[k] = self.theta_idx[k, self.theta_names.index('t_f')]
        self.tf_local = m.offset[('theta', 't_f')][0]      # position of t_f in the node input column

    def default_p(self):
        p = np.zeros(self.n_p)
        p[self.p_theta0:] = self.spec.theta0_values
        return p

    def x_base(self, k, r):
        """V offset of the state block X_r of interval k (r=0: x_k, r>=1: coll_var[k, r-1].x)."""
        return self.v_x[k] if r == 0 else self.v_cx[k][r - 1]

    # ------------------------------------------------------------------ node-local patterns
    def _local_patterns(self):
        m = self.model
        self.local_in = np.concatenate([m.vars, m.params[m.n_theta0:]])     # 59 vars + phi
        self.n_local = len(self.local_in)
        eq, ineq = m.eq_exprs(), m.ineq_exprs()
        self.n_eq, self.n_ineq = len(eq), len(ineq)
        self.dep_eq = S.dependency(eq, self.local_in)
        self.dep_ineq = S.dependency(ineq, self.local_in)
        # rows of the shooting-node equality subset (constraints.py:170-207): rows whose structural
        # Jacobian w.r.t. every non-x variable is empty are not enforced at the shooting nodes
        n_x = self.n_x
        self.shoot_eq_rows = [i for i, dcols in enumerate(self.dep_eq) if any((c >= n_x and c < self.n_var) for c in dcols)]
        o = m.offset
        self.rng = {}
        k = 0
        for vt in VAR_TYPES:
            sz = sum(s for _, s in m.var_list[vt])
            self.rng[vt] = (k, k + sz)
            k += sz
        self.rng['phi'] = (k, k + self.n_phi)

    def _split(self, cols):
        out = {key: [] for key in ('x', 'xdot', 'u', 'z', 'theta', 'phi')}
        for c in cols:
            for key, (a, b) in self.rng.items():
                if a <= c < b:
                    out[key].append(c - a)
        return out

    # ------------------------------------------------------------------ g rows
    def _rows(self):
        n_k, d = self.n_k, self.d
        self.n_shoot_eq = len(self.shoot_eq_rows)
        r = 0
        # initial conditions  x(0) - x_ref(0)  come first (constraints.py:60-65, operation.py:303-326).  The
        # reference iterates over a Python *set* of state names there, so its row order inside this block
        # depends on the interpreter's string hash seed; ``nlp['initial_order']`` states the order to
        # reproduce (default: model order).
        m = self.model
        self.row_initial = 0
        self.initial_idx = []
        if self.initial_conditions:
            order = self.spec.nlp.get('initial_order') or m.names('x')
            if sorted(order) != sorted(m.names('x')):
                raise ValueError('initial_order must be a permutation of the state names')
            for nm in order:
                if nm.rstrip('0123456789') == 'r':
                    raise NotImplementedError('initial conditions on a rotation matrix state (6-DOF mpc)')
                o, sz = m.offset[('x', nm)]
                self.initial_idx += list(range(o, o + sz))
            r += len(self.initial_idx)
        self.row_shoot, self.row_path, self.row_coll, self.row_cont = [], [], [], []
        for k in range(n_k):
            self.row_shoot.append(r)
            r += self.n_shoot_eq
            self.row_path.append(r)
            r += self.n_ineq
            self.row_coll.append([])
            for j in range(d):
                self.row_coll[k].append(r)
                r += self.n_eq
            self.row_cont.append(r)
            r += self.n_x
        # periodicity: every non-integral state, names sorted (operation.py:245-267)
        self.periodic_idx = []
        for nm in (m.sorted_names('x') if self.periodic else []):
            if nm in ('e',):
                continue
            o, sz = m.offset[('x', nm)]
            self.periodic_idx += list(range(o, o + sz))
        self.row_periodic = r
        r += len(self.periodic_idx)
        self.row_tf_bounds = None
        if self.single_reelout:
            self.row_tf_bounds = r
            r += 2
        self.n_g = r

    # ------------------------------------------------------------------ jac_g pattern + scatter tables
    def _pattern(self):
        n_k, d = self.n_k, self.d
        C, D = self.coll.coeff_collocation, self.coll.coeff_continuity
        ent_r, ent_c = [], []          # every structural entry once

        def add(r, c):
            ent_r.append(r)
            ent_c.append(c)

        # ---- collocation-node local output list (same for every (k, j)); see codegen
        #   ('res', i) | ('X', i, m, r) | ('u', i, c) | ('z', i, c) | ('theta', i, t) | ('phi', i, c)
        eq_split = [self._split(cols) for cols in self.dep_eq]
        self.coll_out = []
        for i in range(self.n_eq):
            self.coll_out.append(('res', i))
        #   ('Xc', i, m, r): d r_i/d X_r[m] through xdot (plus the direct part when r is the node's own
        #   block and (i, m) is also a direct dependency);  ('Xo', i, m): direct-only dependency.
        #   Order: state component m outermost.  Everything derived from d r/d x_m and d r/d xdot_m is
        #   emitted together (short live ranges in the generated code), and one output chunk of the d
        #   nodes of an interval then holds the complete interval part of the CCS columns X_r[m],
        #   r = 0..d — what the interval-cooperative flush of the CUDA kernels turns into full sectors.
        for mm in range(self.n_x):
            for i in range(self.n_eq):
                if mm in eq_split[i]['x'] and mm not in eq_split[i]['xdot']:
                    self.coll_out.append(('Xo', i, mm))
            for r in range(d + 1):
                for i in range(self.n_eq):
                    if mm in eq_split[i]['xdot']:
                        self.coll_out.append(('Xc', i, mm, r))
        for c in range(self.n_u):
            for i in range(self.n_eq):
                if c in eq_split[i]['u']:
                    self.coll_out.append(('u', i, c))
        for c in range(self.n_z):
            for i in range(self.n_eq):
                if c in eq_split[i]['z']:
                    self.coll_out.append(('z', i, c))
        for t in range(self.n_theta):
            for i in range(self.n_eq):
                direct = t in eq_split[i]['theta']
                via_tf = (t == self.tf_local - self.rng['theta'][0]) and len(eq_split[i]['xdot']) > 0
                if direct or via_tf:
                    self.coll_out.append(('theta', i, t))
        for c in range(self.n_phi):
            for i in range(self.n_eq):
                if c in eq_split[i]['phi']:
                    self.coll_out.append(('phi', i, c))
        self.eq_split = eq_split

        # ---- collocation-node output list of the CUDA kernels (awebox_b200/csrc/awegpu_kernels.cu).  The thread
        #   that evaluates node (k, j) OWNS the CCS columns of its own state block X_{j+1} and of z_{k,j}: it writes
        #   every entry of those columns, the rows of the d-1 sibling nodes of the interval included
        #   (d r_sib / d X_{j+1}[m] = d r/d xdot_m (at the sibling) * C[j+1, sib+1] / (h t_f): the d r/d xdot block is a
        #   few dozen cheap entries, re-evaluated from the sibling's inputs).  With the constant continuity /
        #   periodicity entries those columns are one contiguous range of the CCS value array per instance, which
        #   leaves the SM as whole, aligned runs.  Everything else the node contributes ("head": residual rows, the
        #   entries in the x_k, u_k, theta and phi columns) goes through the interval-cooperative flush.
        #     head: ('res', i) | ('Xc', i, m, 0) | ('u', i, c) | ('theta', i, t) | ('phi', i, c)
        #     own:  ('O', i, m)       row i of the node itself, column X_{j+1}[m]   (direct + through-xdot part)
        #           ('S', dl, i, m)   row i of sibling (j + dl) mod d, column X_{j+1}[m]
        #           ('z', i, c)       row i of the node itself, column z_{k,j}[c]
        #   own outputs are listed column by column; kcoll_cols[c] = (first own output, count) of local column c
        #   (c < n_x: X_{j+1}[c], then z)
        self.kcoll_head = [('res', i) for i in range(self.n_eq)]
        for mm in range(self.n_x):
            for i in range(self.n_eq):
                if mm in eq_split[i]['xdot']:
                    self.kcoll_head.append(('Xc', i, mm, 0))
        self.kcoll_head += [it for it in self.coll_out if it[0] in ('u', 'theta', 'phi')]
        self.kcoll_own, self.kcoll_cols = [], []
        for mm in range(self.n_x):
            first = len(self.kcoll_own)
            for i in range(self.n_eq):
                if mm in eq_split[i]['x'] or mm in eq_split[i]['xdot']:
                    self.kcoll_own.append(('O', i, mm))
            for dl in range(1, d):
                for i in range(self.n_eq):
                    if mm in eq_split[i]['xdot']:
                        self.kcoll_own.append(('S', dl, i, mm))
            self.kcoll_cols.append((first, len(self.kcoll_own) - first))
        for c in range(self.n_z):
            first = len(self.kcoll_own)
            for i in range(self.n_eq):
                if c in eq_split[i]['z']:
                    self.kcoll_own.append(('z', i, c))
            self.kcoll_cols.append((first, len(self.kcoll_own) - first))
        self.kcoll_out = self.kcoll_head + self.kcoll_own
        self.n_kcoll_head = len(self.kcoll_head)

        # ---- shooting-node local outputs: residual rows then direct entries
        sh_rows = [('eq', i) for i in self.shoot_eq_rows] + [('ineq', i) for i in range(self.n_ineq)]
        self.shoot_rows = sh_rows
        self.shoot_out = [('res', n) for n in range(len(sh_rows))]
        sh_dep = [self.dep_eq[i] if kind == 'eq' else self.dep_ineq[i] for kind, i in sh_rows]
        self.shoot_dep = sh_dep
        for c in range(self.n_local):
            for n, cols in enumerate(sh_dep):
                if c in cols:
                    self.shoot_out.append(('jac', n, c))

        # ---- global entries
        coll_keys = np.full((n_k, d, len(self.coll_out), 2), -1, dtype=np.int64)
        shoot_keys = np.full((n_k, len(self.shoot_out), 2), -1, dtype=np.int64)
        for k in range(n_k):
            # shooting node k
            xb, ub, xdb, zb = self.v_x[k], self.v_u[k], self.v_xdot[k], self.v_z[k]
            for n, item in enumerate(self.shoot_out):
                if item[0] != 'jac':
                    continue
                _, rr, c = item
                row = (self.row_shoot[k] + rr) if rr < self.n_shoot_eq else (self.row_path[k] + rr - self.n_shoot_eq)
                col = self._local_to_v(c, k, xb, xdb, ub, zb)
                shoot_keys[k, n] = (row, col)
                add(row, col)
            for j in range(d):
                row0 = self.row_coll[k][j]
                for n, item in enumerate(self.coll_out):
                    kind = item[0]
                    if kind == 'res':
                        continue
                    i = item[1]
                    row = row0 + i
                    if kind == 'theta':
                        col = int(self.theta_idx[k, item[2]])
                    elif kind == 'phi':
                        col = self.v_phi + item[2]
                    elif kind == 'u':
                        col = self.v_u[k] + item[2]
                    elif kind == 'z':
                        col = self.v_cz[k][j] + item[2]
                    elif kind == 'Xo':
                        col = self.x_base(k, j + 1) + item[2]
                    else:
                        _, _, mm, r = item
                        has_b = C[r, j + 1] != 0.0
                        has_a = (mm in self.eq_split[i]['x']) and (r == j + 1)
                        if not (has_a or has_b):
                            continue
                        col = self.x_base(k, r) + mm
                    coll_keys[k, j, n] = (row, col)
                    add(row, col)
        kcoll_keys = np.full((n_k, d, len(self.kcoll_out), 2), -1, dtype=np.int64)
        for k in range(n_k):
            for j in range(d):
                for n, item in enumerate(self.kcoll_out):
                    kind = item[0]
                    if kind == 'res':
                        continue
                    if kind == 'S':
                        _, dl, i, mm = item
                        sib = (j + dl) % d
                        if C[j + 1, sib + 1] == 0.0:
                            continue
                        kcoll_keys[k, j, n] = (self.row_coll[k][sib] + i, self.x_base(k, j + 1) + mm)
                        continue
                    i = item[1]
                    row = self.row_coll[k][j] + i
                    if kind == 'O':
                        mm = item[2]
                        if not (mm in self.eq_split[i]['x'] or C[j + 1, j + 1] != 0.0):
                            continue
                        col = self.x_base(k, j + 1) + mm
                    elif kind == 'Xc':
                        if C[0, j + 1] == 0.0:
                            continue
                        col = self.x_base(k, 0) + item[2]
                    elif kind == 'theta':
                        col = int(self.theta_idx[k, item[2]])
                    elif kind == 'phi':
                        col = self.v_phi + item[2]
                    elif kind == 'u':
                        col = self.v_u[k] + item[2]
                    else:
                        col = self.v_cz[k][j] + item[2]
                    kcoll_keys[k, j, n] = (row, col)
        # ---- constant entries: continuity, periodicity, t_f bounds
        const_r, const_c, const_v = [], [], []
        const_begin = []
        for k in range(n_k):
            const_begin.append(len(const_r))
            for mm in range(self.n_x):
                row = self.row_cont[k] + mm
                const_r.append(row), const_c.append(self.v_x[k + 1] + mm), const_v.append(1.0)
                for r in range(d + 1):
                    if D[r] != 0.0:
                        const_r.append(row), const_c.append(self.x_base(k, r) + mm), const_v.append(-D[r])
        last = self.v_cx[n_k - 1][d - 1]
        for n, mm in enumerate(self.periodic_idx):
            row = self.row_periodic + n
            const_r.append(row), const_c.append(self.v_x[0] + mm), const_v.append(1.0)
            const_r.append(row), const_c.append(last + mm), const_v.append(-1.0)
        # ---- affine rows  g[aff_row[n]] = scale * (sum_i aff_coef[i] * w[aff_col[i]] + sum_i affp_coef[i] * p[affp_col[i]]
        #                                         + aff_const[n])
        aff_row, aff_begin, aff_col, aff_coef, aff_const = [], [0], [], [], []
        affp_begin, affp_col, affp_coef = [0], [], []

        def affine(row, w_terms, p_terms, const):
            aff_row.append(row)
            for c, a in w_terms:
                aff_col.append(c), aff_coef.append(a)
            for c, a in p_terms:
                affp_col.append(c), affp_coef.append(a)
            aff_begin.append(len(aff_col)), affp_begin.append(len(affp_col)), aff_const.append(const)

        # initial conditions of an mpc trial: x_0 - p.ref[x_0]  (operation.py:303-326; P['p','ref'] is a V-shaped
        # block at the start of P, discretization.py:157-179)
        for n, mm in enumerate(self.initial_idx):
            affine(self.row_initial + n, [(self.v_x[0] + mm, 1.0)], [(self.v_x[0] + mm, -1.0)], 0.0)
        # t_f bound rows of phase_fix='single_reelout' (constraints.py:147-168 with ocp_outputs.py:118-141):
        # period = t_f[0] * n_ro / n_k + t_f[1] * (n_k - n_ro) / n_k, rows (period - ub) / s and (lb - period) / s,
        # s = phase_fix_reelout
        if self.single_reelout:
            s_ro = float(self.spec.nlp.get('phase_fix_reelout', 0.7))
            n_ro = int(round(n_k * s_ro))
            o_tf, sz_tf = self.v_theta['t_f']
            assert sz_tf == 2
            sc_tf = float(self.model.scale('theta', 't_f')[0])
            lb, ub = [float(b) / sc_tf for b in self.spec.opt('system_bounds.theta.t_f')]
            w0, w1 = float(n_ro) / n_k, float(n_k - n_ro) / n_k
            for n, (sign, bound) in enumerate(((1.0, -ub), (-1.0, lb))):
                affine(self.row_tf_bounds + n, [(o_tf, sign * w0 / s_ro), (o_tf + 1, sign * w1 / s_ro)], [], bound / s_ro)
        for n in range(len(aff_row)):
            for i in range(aff_begin[n], aff_begin[n + 1]):
                const_r.append(aff_row[n]), const_c.append(aff_col[i]), const_v.append(aff_coef[i])
        self.aff_row = np.array(aff_row, dtype=np.int32)
        self.aff_begin = np.array(aff_begin, dtype=np.int32)
        self.aff_col = np.array(aff_col, dtype=np.int32)
        self.aff_coef = np.array(aff_coef, dtype=float)
        self.aff_const = np.array(aff_const, dtype=float)
        self.affp_begin = np.array(affp_begin, dtype=np.int32)
        self.affp_col = np.array(affp_col, dtype=np.int32)
        self.affp_coef = np.array(affp_coef, dtype=float)
        self.n_aff = len(aff_row)
        self.needs_full_p = len(affp_col) > 0
        const_begin.append(len(const_r))       # periodicity / affine-row entries belong to the last interval
        self.const_begin = np.array(const_begin, dtype=np.int32)
        for r_, c_ in zip(const_r, const_c):
            add(r_, c_)

        ent_r = np.array(ent_r, dtype=np.int64)
        ent_c = np.array(ent_c, dtype=np.int64)
        key = ent_c * self.n_g + ent_r
        ukey = np.unique(key)
        if len(ukey) != len(key):
            raise AssertionError('duplicate structural entries in jac_g assembly')
        self.nnz = len(ukey)
        self.row = (ukey % self.n_g).astype(np.int64)
        cols = (ukey // self.n_g).astype(np.int64)
        self.colind = np.zeros(self.n_v + 1, dtype=np.int64)
        np.add.at(self.colind, cols + 1, 1)
        self.colind = np.cumsum(self.colind)

        def slots(keys):
            flat = keys.reshape(-1, 2)
            out = np.full(flat.shape[0], -1, dtype=np.int32)
            ok = flat[:, 0] >= 0
            kk = flat[ok, 1] * self.n_g + flat[ok, 0]
            out[ok] = np.searchsorted(ukey, kk).astype(np.int32)
            return out.reshape(keys.shape[:-1])

        self.coll_slot = slots(coll_keys)            # [n_k, d, n_coll_out]   (-1: not an entry)
        self.kcoll_slot = slots(kcoll_keys)          # [n_k, d, n_kcoll_out]  the same entries, kernel output order
        a_ = np.sort(self.coll_slot[self.coll_slot >= 0])
        b_ = np.sort(self.kcoll_slot[self.kcoll_slot >= 0])
        if len(a_) != len(b_) or np.any(a_ != b_):
            raise AssertionError('kernel output list does not cover the collocation entries exactly once')
        self.shoot_slot = slots(shoot_keys)          # [n_k, n_shoot_out]
        ck = np.stack([np.array(const_r, dtype=np.int64), np.array(const_c, dtype=np.int64)], axis=1)
        self.const_slot = slots(ck)
        self.const_val = np.array(const_v, dtype=float) * self.constraint_scale

    def _local_to_v(self, c, k, xb, xdb, ub, zb):
        for key, (a, b) in self.rng.items():
            if a <= c < b:
                if key == 'x':
                    return xb + (c - a)
                if key == 'xdot':
                    return xdb + (c - a)
                if key == 'u':
                    return ub + (c - a)
                if key == 'z':
                    return zb + (c - a)
                if key == 'theta':
                    return int(self.theta_idx[k, c - a])
                return self.v_phi + (c - a)
        raise IndexError(c)

    # ------------------------------------------------------------------ summaries
    def algorithmic_bytes(self):
        """compulsory HBM traffic per OCP instance and call: every datum once."""
        return 8 * (self.n_v + self.n_theta0 + self.n_g + self.nnz)
