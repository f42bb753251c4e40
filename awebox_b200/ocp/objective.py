"""Objective of the discretised OCP, ``f(w, p)`` and its gradient with respect to ``w``.

Follows (direct collocation, ``trajectory.type = 'power_cycle'``):
  awebox/ocp/objective.py:44-96, 120-171     general regularisation: weighted squares of (variables - reference) at
                                             every collocation node, sorted into cost categories
  awebox/ocp/objective.py:173-245            ... mapped over the n_k * d collocation nodes with the quadrature weights
  awebox/ocp/objective.py:261-300            homotopy-parameter costs, time cost, power cost
  awebox/ocp/objective.py:392-418            side-slip (beta) cost
  awebox/ocp/objective.py:330-390, 422-455   problem costs and the objective
      objective = psi * tracking + (1 - psi) * power + general + homotopy
  awebox/ocp/collocation.py:185-200, 272-316 quadrature weights, integral outputs (the energy e)
  awebox/mdl/dynamics.py:206-232, 318-330    integrand of e: lambda10 * l_t * dl_t / scaling(e)
  awebox/ocp/ocp_outputs.py:118-141          time period
  awebox/ocp/discretization.py:129-154       cost-factor block of P

What is evaluated per collocation node, on the node's variable column ``v = [x; xdot; u; z; theta]`` (xdot from the
collocation polynomial):
    reg_c  = sum_i [category(i) = c] weight_i (v_i - ref_i)^2,   weight_i = P.weights_i * P.cost_c / normalization_c
    beta^2 (6-DOF kites), power integrand
and then combined per instance.  ``NodeIntegrands`` holds the two model expressions and their gradients;
``Objective.evaluate`` is the host restatement used to pin the layout (numpy, one instance at a time: test
infrastructure and documentation of the arithmetic); the product path is ``csrc/awegpu_objective.cu`` behind
``awegpu_eval_objective`` (include/awegpu.h).
"""
import numpy as np

from .. import sym as S

REG_CATEGORIES = ['tracking', 'xdot_regularisation', 'u_regularisation', 'fictitious', 'theta_regularisation']
COST_NAMES = ['tracking', 'u_regularisation', 'xdot_regularisation', 'gamma', 'iota', 'psi', 'tau', 'eta', 'nu',
              'upsilon', 'fictitious', 'power', 'power_derivative', 't_f', 'theta_regularisation', 'nominal_landing',
              'compromised_battery', 'transition', 'beta', 'P_max']          # discretization.py:129-154


class NodeIntegrands(object):
    """model expressions the objective integrates over the collocation nodes: the power integrand and the side-slip
    angles, with their sparse gradients with respect to the node's variable column"""

    def __init__(self, model):
        m = self.model = model
        lam = m.si('z', 'lambda10')[0]
        l_t = m.si(m.var_type_of('l_t'), 'l_t')[0]
        dl_t = m.si(m.var_type_of('dl_t'), 'dl_t')[0]
        e_scale = float(m.spec.opt('scaling.x.e'))
        self.power = lam * l_t * dl_t / e_scale                         # dynamics.py:206-211, 318-330
        self.betas = [m.aero[k]['beta'] for k in sorted(m.aero)] if int(m.spec.opt('kite_dof')) == 6 else []
        exprs = [self.power] + self.betas
        jac = S.jacobian(exprs, m.vars)
        self.exprs = exprs
        self.grad = [sorted(j.items()) for j in jac]                   # [(var index, expr)]

    def flat_outputs(self):
        """-> (expressions [value_0, d value_0 ..., value_1, ...], layout [(first, var indices)])"""
        out, layout = [], []
        for e, g in zip(self.exprs, self.grad):
            layout.append((len(out), [c for c, _ in g]))
            out.append(e)
            out += [ge for _, ge in g]
        return out, layout


class Objective(object):
    def __init__(self, ocp):
        self.ocp = ocp
        m = ocp.model
        nlp = ocp.spec.nlp
        if ocp.traj_type != 'power_cycle':
            raise NotImplementedError('objective: trajectory type %r (the tracking-MPC cost is built by awebox/pmpc.py, '
                                      'not by ocp/objective.py)' % ocp.traj_type)
        if nlp.get('cost.P_max', False) or nlp.get('cost.PDGA', False) or not nlp.get('cost.output_quadrature', True):
            raise NotImplementedError('objective: cost.P_max / cost.PDGA / cost.output_quadrature=False')
        # the normalisations are n_k times a count of nodes / kites (awebox/opts/funcs.py:177-182): the spec records them
        # for its own n_k
        rescale = float(ocp.n_k) / float(nlp['n_k'])
        self.norm = {c: float(nlp['cost.normalization.' + c]) * rescale for c in REG_CATEGORIES + ['beta']}
        # category of every entry of the node's variable column (objective.py:120-143)
        sorting = {'x': ('tracking', {'e': None}), 'xdot': ('xdot_regularisation', {}),
                   'u': ('u_regularisation', {'f_fict': 'fictitious', 'm_fict': 'fictitious'}),
                   'z': ('tracking', {}), 'theta': ('theta_regularisation', {'t_f': None})}
        for vt, name, cat in (nlp.get('cost.adjustments_to_general_regularization_distribution') or []):
            sorting[vt][1][name] = cat
        cat_of = []
        for vt in ('x', 'xdot', 'u', 'z', 'theta'):
            for name, size in m.var_list[vt]:
                base = name.rstrip('0123456789')
                default, exceptions = sorting[vt]
                cat = exceptions[base] if base in exceptions else default
                cat_of += [None if cat is None else cat.replace('_cost', '')] * size
        self.cat_of = cat_of
        self.cat_index = np.array([-1 if c is None else REG_CATEGORIES.index(c) for c in cat_of], dtype=np.int32)
        self.quad = np.array(ocp.coll.quad_weights, dtype=float).ravel()
        self.integrands = NodeIntegrands(m)
        self.n_cost = len(COST_NAMES)
        self.p_weights = ocp.n_v                       # P = [p.ref (n_v) | p.weights (n_var) | cost (20) | theta0]
        self.p_cost = ocp.n_v + ocp.n_var
        # V index of the reference / variable of every node-column entry: filled per node in node_columns()
        self.phi_idx = [ocp.v_phi + i for i in range(ocp.n_phi)]
        self.phi_names = list(ocp.spec.phi_names)

    # ------------------------------------------------------------------ layout helpers
    def node_gather(self, k, j):
        """V indices of the node's column entries that are read directly: -> (idx [n_var], kind [n_var]) with kind 0 =
        V entry, 1 = xdot entry m (idx = m; comes from the collocation polynomial)"""
        ocp = self.ocp
        idx, kind = [], []
        for mm in range(ocp.n_x):
            idx.append(ocp.v_cx[k][j] + mm), kind.append(0)
        for mm in range(ocp.n_x):
            idx.append(mm), kind.append(1)
        for c in range(ocp.n_u):
            idx.append(ocp.v_u[k] + c), kind.append(0)
        for c in range(ocp.n_z):
            idx.append(ocp.v_cz[k][j] + c), kind.append(0)
        for t in range(ocp.n_theta):
            idx.append(int(ocp.theta_idx[k, t])), kind.append(0)
        return np.array(idx), np.array(kind)

    def time_period_terms(self):
        """T = sum_i coef_i * V[idx_i]  (ocp_outputs.py:118-141)"""
        ocp = self.ocp
        o_tf, sz = ocp.v_theta['t_f']
        if sz == 2:
            n_ro = int(round(ocp.n_k * float(ocp.spec.nlp.get('phase_fix_reelout', 0.7))))
            return [(o_tf, float(n_ro) / ocp.n_k), (o_tf + 1, float(ocp.n_k - n_ro) / ocp.n_k)]
        return [(o_tf, 1.0)]

    # ------------------------------------------------------------------ host restatement (test infrastructure)
    def evaluate(self, w, p, components=False):
        """-> f, grad [n_v] (and the dictionary of cost components) for one instance."""
        ocp, m = self.ocp, self.ocp.model
        d, n_k, n_x = ocp.d, ocp.n_k, ocp.n_x
        C = ocp.coll.coeff_collocation
        w = np.asarray(w, dtype=float)
        p = np.asarray(p, dtype=float)
        ref, pw, cost = p[:ocp.n_v], p[self.p_weights:self.p_weights + ocp.n_var], p[self.p_cost:self.p_cost + self.n_cost]
        cf = dict(zip(COST_NAMES, cost))
        wgt = np.array([0.0 if c is None else pw[i] * cf[c] / self.norm[c] for i, c in enumerate(self.cat_of)])
        reg = np.zeros(len(REG_CATEGORIES))
        g_reg = np.zeros((len(REG_CATEGORIES), ocp.n_v))
        beta_sum, g_beta = 0.0, np.zeros(ocp.n_v)
        energy, g_energy = 0.0, np.zeros(ocp.n_v)
        params = np.concatenate([p[ocp.p_theta0:], w[ocp.v_phi:ocp.v_phi + ocp.n_phi]])
        flat, layout = self.integrands.flat_outputs()
        leaves = list(m.vars) + list(m.params)
        for k in range(n_k):
            tf = w[ocp.tf_idx[k]]
            for j in range(d):
                idx, kind = self.node_gather(k, j)
                cc = C[:, j + 1] * n_k / tf
                Xr = np.array([w[ocp.x_base(k, r):ocp.x_base(k, r) + n_x] for r in range(d + 1)])
                xdot = cc @ Xr
                v = np.where(kind == 0, w[np.where(kind == 0, idx, 0)], 0.0)
                v[kind == 1] = xdot
                r_ = np.where(kind == 0, ref[np.where(kind == 0, idx, 0)], 0.0)          # reference of xdot: 0
                diff = v - r_
                q = self.quad[j]
                dv = np.zeros(ocp.n_var)                  # d (node cost) / d v, per category applied below

                def scatter(g_vec, gv):
                    """chain a gradient with respect to the node column into V"""
                    for i in range(ocp.n_var):
                        if gv[i] == 0.0:
                            continue
                        if kind[i] == 0:
                            g_vec[idx[i]] += gv[i]
                        else:
                            mm = idx[i]
                            for r in range(d + 1):
                                g_vec[ocp.x_base(k, r) + mm] += gv[i] * cc[r]
                            g_vec[ocp.tf_idx[k]] -= gv[i] * xdot[mm] / tf

                for ci in range(len(REG_CATEGORIES)):
                    sel = self.cat_index == ci
                    reg[ci] += q * np.sum(wgt[sel] * diff[sel] ** 2)
                    gv = np.where(sel, 2.0 * q * wgt * diff, 0.0)
                    scatter(g_reg[ci], gv)
                vals = S.evaluate(flat, leaves, np.concatenate([v, params]))
                pw_first, pw_cols = layout[0]
                h = tf / n_k
                energy += h * q_energy(self, j) * vals[pw_first]
                gv = np.zeros(ocp.n_var)
                gv[pw_cols] = h * q_energy(self, j) * np.array(vals[pw_first + 1:pw_first + 1 + len(pw_cols)])
                scatter(g_energy, gv)
                g_energy[ocp.tf_idx[k]] += q_energy(self, j) * vals[pw_first] / n_k
                for b_first, b_cols in layout[1:]:
                    beta = vals[b_first]
                    beta_sum += q * beta ** 2
                    gv = np.zeros(ocp.n_var)
                    gv[b_cols] = 2.0 * q * beta * np.array(vals[b_first + 1:b_first + 1 + len(b_cols)])
                    scatter(g_beta, gv)
                del dv
        comp = {}
        for ci, c in enumerate(REG_CATEGORIES):
            comp[c + '_cost'] = reg[ci]
        beta_f = cf['beta'] / self.norm['beta']
        comp['beta_cost'] = beta_f * beta_sum
        T_terms = self.time_period_terms()
        T = sum(c * w[i] for i, c in T_terms)
        T_ref = sum(c * ref[i] for i, c in T_terms)
        comp['time_cost'] = cf['t_f'] * (T - T_ref) ** 2
        comp['power_cost'] = -cf['power'] * energy / T
        hom = 0.0
        for n, name in enumerate(self.phi_names):
            comp[name + '_cost'] = cf[name] * w[self.phi_idx[n]]
            hom += comp[name + '_cost']
        comp['homotopy_cost'] = hom
        comp['tracking_problem_cost'] = comp['tracking_cost']
        comp['power_problem_cost'] = comp['power_cost']
        comp['general_problem_cost'] = (comp['fictitious_cost'] + comp['u_regularisation_cost'] + comp['xdot_regularisation_cost']
                                        + comp['theta_regularisation_cost'] + comp['beta_cost'] + comp['time_cost'])
        i_psi = self.phi_idx[self.phi_names.index('psi')]
        psi = w[i_psi]
        f = psi * comp['tracking_cost'] + (1.0 - psi) * comp['power_cost'] + comp['general_problem_cost'] + hom
        comp['objective'] = f
        # gradient
        ti = REG_CATEGORIES.index('tracking')
        g = psi * g_reg[ti]
        for ci, c in enumerate(REG_CATEGORIES):
            if ci != ti:
                g = g + g_reg[ci]
        g = g + beta_f * g_beta
        gT = np.zeros(ocp.n_v)
        for i, c in T_terms:
            gT[i] = c
        g = g + 2.0 * cf['t_f'] * (T - T_ref) * gT
        g = g + (1.0 - psi) * (-cf['power']) * (g_energy / T - energy / T ** 2 * gT)
        g[i_psi] += comp['tracking_cost'] - comp['power_cost']
        for n, name in enumerate(self.phi_names):
            g[self.phi_idx[n]] += cf[name]
        return (f, g, comp) if components else (f, g)


def q_energy(obj, j):
    """weight of collocation node j in the integral of an interval (collocation.py:272-316 with Radau points:
    the continuity coefficients pick the last node, so the weights are the last column of Lambda = the quadrature
    weights)"""
    return obj.quad[j]


class MpcObjective(object):
    """The tracking cost of the periodic MPC (awebox/pmpc.py:304-358, 691-706; BASELINE configs[4]) over the MPC trial:

        f = 1/N sum_k sum_j q_j [ (x_kj - ref)' Q (x_kj - ref) + |z_kj - ref|^2 + (u_k - ref)' R (u_k - ref) ]
            + (x_N - ref)' P (x_N - ref)

    with diagonal Q, R, P and u_param = 'zoh'.  The parameter vector is Pmpc's own (pmpc.py:152-162):
    ``p = [x0 (n_x) | ref (n_v, laid out like V) | u_ref (1) | Q (n_x) | R (n_u) | P (n_x)]``.
    Every entry i of V carries at most one term: ``fac[i] * W[kind[i]][idx[i]] * (w_i - ref_i)^2``; the tables below are
    what the CUDA kernel reads.  ``kind``: 0 none, 1 Q, 2 R, 3 P, 4 unit weight (algebraic variables)."""

    KIND_NONE, KIND_Q, KIND_R, KIND_P, KIND_ONE = 0, 1, 2, 3, 4

    def __init__(self, ocp):
        if ocp.traj_type != 'mpc':
            raise NotImplementedError('MpcObjective: trajectory type %r' % ocp.traj_type)
        self.ocp = ocp
        n_v, n_x, n_u = ocp.n_v, ocp.n_x, ocp.n_u
        self.p_x0, self.p_ref = 0, n_x
        self.p_uref = n_x + n_v
        self.p_Q = self.p_uref + 1
        self.p_R = self.p_Q + n_x
        self.p_P = self.p_R + n_u
        self.n_p = self.p_P + n_x
        quad = np.array(ocp.coll.quad_weights, dtype=float).ravel()
        kind = np.zeros(n_v, dtype=np.int32)
        idx = np.zeros(n_v, dtype=np.int32)
        fac = np.zeros(n_v)
        N = float(ocp.n_k)
        for k in range(ocp.n_k):
            for j in range(ocp.d):
                for m in range(n_x):
                    i = ocp.v_cx[k][j] + m
                    kind[i], idx[i], fac[i] = self.KIND_Q, m, quad[j] / N
                for m in range(ocp.n_z):
                    i = ocp.v_cz[k][j] + m
                    kind[i], idx[i], fac[i] = self.KIND_ONE, m, quad[j] / N
            # the same u_k enters the d terms of the interval one after the other (pmpc.py:338-342)
            q_sum = 0.0
            for j in range(ocp.d):
                q_sum += quad[j] / N
            for m in range(n_u):
                i = ocp.v_u[k] + m
                kind[i], idx[i], fac[i] = self.KIND_R, m, q_sum
        for m in range(n_x):
            i = ocp.v_x[ocp.n_k] + m
            kind[i], idx[i], fac[i] = self.KIND_P, m, 1.0
        self.kind, self.idx, self.fac = kind, idx, fac

    def weights(self, p):
        """-> per-entry weight [n_v] for a parameter vector."""
        table = [np.zeros(1), p[self.p_Q:self.p_Q + self.ocp.n_x], p[self.p_R:self.p_R + self.ocp.n_u],
                 p[self.p_P:self.p_P + self.ocp.n_x], np.ones(max(1, self.ocp.n_z))]
        w = np.zeros(self.ocp.n_v)
        for kd in (self.KIND_Q, self.KIND_R, self.KIND_P, self.KIND_ONE):
            sel = self.kind == kd
            w[sel] = table[kd][self.idx[sel]]
        return w * self.fac

    def evaluate(self, w, p):
        """-> f, grad [n_v] for one instance (host restatement: test infrastructure)."""
        wt = self.weights(np.asarray(p, dtype=float))
        diff = np.asarray(w, dtype=float) - p[self.p_ref:self.p_ref + self.ocp.n_v]
        return float(np.sum(wt * diff * diff)), 2.0 * wt * diff
