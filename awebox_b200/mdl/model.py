"""Node model: the implicit index-1 DAE residual and the path inequalities of a tree-structured
tethered-kite system, re-derived on the product's own expression graph (``awebox_b200.sym``).

This is the from-scratch restatement of the reference's per-node ``casadi.Function``s

    mdl_eq_fun   = Model.dynamics                  awebox/mdl/model.py:125,  ocp/constraints.py:239
    mdl_ineq_fun                                    awebox/ocp/constraints.py:261

with the input column ``[x; xdot; u; z; theta]`` in scaled units and the parameter column
``[theta0; phi]`` (awebox/mdl/system.py:224-228, 427-430).  Each builder below cites the reference
lines whose behaviour it reproduces.  Supported: 6-DOF and 3-DOF kites on any tree architecture,
'com' / 'stick' tether attachment, multi-element tether drag with constant cd, ISA / uniform
atmosphere, power / log / uniform wind, zero-induction aerodynamics, the inequality families of
dynamics.py (tether force / stress, airspeed, aero validity, anticollision, acceleration, yaw).
Everything else (induction models, cross tethers, drag mode, lifted forces) raises.
"""
import math
import numpy as np

from .. import sym as S
from ..sym import Expr


# ----------------------------------------------------------------------------- architecture
class Architecture(object):
    """tree bookkeeping (awebox/mdl/architecture.py:38-170)."""

    def __init__(self, parent_map):
        self.parent_map = dict(sorted(parent_map.items()))
        self.number_of_nodes = len(self.parent_map) + 1
        parents = set(self.parent_map.values())
        self.kite_nodes = [n for n in self.parent_map if n not in parents]
        self.number_of_kites = len(self.kite_nodes)
        if self.number_of_nodes == 2:
            self.layer_nodes = [0]
            self.children_map = {0: [1]}
        else:
            self.layer_nodes = sorted(parents - {0})
            self.children_map = {p: [k for k, v in self.parent_map.items() if v == p] for p in sorted(parents)}

    def node_label(self, node):
        return '%d%d' % (node, self.parent_map[node])

    def parent_label(self, node):
        parent = self.parent_map[node]
        grand = self.parent_map[parent] if node > 1 else 0
        return '%d%d' % (parent, grand)


def generate_structure(spec, arch):
    """variable lists per type (awebox/mdl/system.py:42-230), restricted to the supported options."""
    dof = int(spec.opt('kite_dof'))
    surface_control = int(spec.opt('surface_control', 1))
    tether_states = [('q', 3), ('dq', 3)]
    kite_states = [('q', 3), ('dq', 3)]
    kite_controls = [('f_fict', 3)]
    if dof == 3:
        kite_states += [('coeff', 2)]
        kite_controls += [('dcoeff', 2)]
    elif dof == 6:
        kite_states += [('omega', 3), ('r', 9)]
        kite_controls += [('m_fict', 3)]
        if surface_control == 0:
            kite_controls += [('delta', 3)]
        else:
            kite_states += [('delta', 3)]
            kite_controls += [('ddelta', 3)]
    else:
        raise ValueError('kite dof %r not supported' % dof)
    if spec.opt('trajectory.system_type') != 'lift_mode':
        raise NotImplementedError('only lift_mode systems are supported')
    x, u, z, gc = [], [], [], []
    for n in range(1, arch.number_of_nodes):
        lab = arch.node_label(n)
        if n in arch.kite_nodes:
            x += [(nm + lab, sz) for nm, sz in kite_states]
            u += [(nm + lab, sz) for nm, sz in kite_controls]
        else:
            x += [(nm + lab, sz) for nm, sz in tether_states]
        z += [('lambda' + lab, 1)]
        gc += ['q' + lab]
    if spec.opt('cross_tether') and len(arch.kite_nodes) > 1:
        raise NotImplementedError('cross tethers are not supported')
    x += [('l_t', 1), ('dl_t', 1)]
    cv = spec.opt('tether.control_var')
    if cv == 'ddl_t':
        u += [('ddl_t', 1)]
    elif cv == 'dddl_t':
        x += [('ddl_t', 1)]
        u += [('dddl_t', 1)]
    else:
        raise ValueError('invalid tether control variable')
    if not spec.opt('integral_outputs', True):
        x += [('e', 1)]
    if spec.opt('tether.lift_tether_force') or spec.opt('aero.lift_aero_force'):
        raise NotImplementedError('lifted force variables are not supported')
    if spec.opt('induction_model', 'not_in_use') != 'not_in_use':
        raise NotImplementedError('induction models are not supported')
    xdot = [('d' + nm, sz) for nm, sz in x]
    theta = [('diam_t', 1), ('t_f', 1)]
    if arch.number_of_nodes - arch.number_of_kites > 1:
        theta += [('l_s', 1), ('diam_s', 1)]
    if len(arch.layer_nodes) > 1:
        theta += [('l_i', 1), ('diam_i', 1)]
    return {'x': x, 'xdot': xdot, 'u': u, 'z': z, 'theta': theta}, gc


VAR_TYPES = ['x', 'xdot', 'u', 'z', 'theta']
SIMPLIFY_REELOUT = True


class NodeModel(object):
    """symbols + residual expressions of one model node evaluation."""

    def __init__(self, spec):
        self.spec = spec
        self.arch = Architecture(spec.parent_map)
        self.var_list, self.gc = generate_structure(spec, self.arch)
        for vt in VAR_TYPES:
            ref = spec.variables.get(vt)
            if ref is not None and [tuple(v) for v in ref] != self.var_list[vt]:
                raise ValueError('variable list mismatch for %s: spec %r, derived %r' % (vt, ref, self.var_list[vt]))
        # symbols (scaled variables, parameters)
        self.offset = {}
        self.var_labels = []
        k = 0
        for vt in VAR_TYPES:
            for nm, sz in self.var_list[vt]:
                self.offset[(vt, nm)] = (k, sz)
                self.var_labels += ['[%s,%s,%d]' % (vt, nm, i) for i in range(sz)]
                k += sz
        self.n_var = k
        if len(spec.scaling) != k:
            raise ValueError('scaling vector has %d entries, model has %d variables' % (len(spec.scaling), k))
        self.vars = np.array([S.sym(l) for l in self.var_labels], dtype=object)
        self.n_theta0 = len(spec.theta0_labels)
        self.param_labels = list(spec.theta0_labels) + ['[phi,%s,0]' % n for n in spec.phi_names]
        self.params = np.array([S.sym(l) for l in self.param_labels], dtype=object)
        self.n_param = len(self.param_labels)
        self._pidx = {}
        for i, l in enumerate(self.param_labels):
            key = tuple(l[1:-1].split(',')[:-1])
            self._pidx.setdefault(key, []).append(i)
        self.n_x = sum(sz for _, sz in self.var_list['x'])
        self.n_u = sum(sz for _, sz in self.var_list['u'])
        self.n_z = sum(sz for _, sz in self.var_list['z'])
        self.n_theta = sum(sz for _, sz in self.var_list['theta'])
        self._seg_cache = {}
        self.eq = []        # [(name, [Expr])]
        self.ineq = []
        self._build()

    # ---- accessors
    def names(self, vt):
        return [nm for nm, _ in self.var_list[vt]]

    def sorted_names(self, vt):
        return sorted(self.names(vt))

    def has(self, vt, nm):
        return (vt, nm) in self.offset

    def scaled(self, vt, nm):
        k, sz = self.offset[(vt, nm)]
        return self.vars[k:k + sz]

    def scale(self, vt, nm):
        k, sz = self.offset[(vt, nm)]
        return self.spec.scaling[k:k + sz]

    def si(self, vt, nm):
        """scaled -> SI (awebox/tools/struct_operations.py:617-631; identity scaling is skipped)."""
        v, s = self.scaled(vt, nm), self.scale(vt, nm)
        if np.all(s == 1.0):
            return v
        return S.vec([S.const(s[i]) * v[i] for i in range(len(v))])

    def par(self, *path):
        idx = self._pidx[('theta0',) + tuple(path)]
        return self.params[idx[0]] if len(idx) == 1 else self.params[idx]

    def phi(self, name):
        return self.params[self._pidx[('phi', name)][0]]

    def var_type_of(self, name):
        """awebox/tools/struct_operations.py:737-762 — every type but xdot first."""
        for vt in ['x', 'u', 'z', 'theta']:
            if self.has(vt, name):
                return vt
        if self.has('xdot', name):
            return 'xdot'
        raise KeyError(name)

    # ---- primitives (awebox/tools/vector_operations.py:48-131)
    @staticmethod
    def smooth_sqrt(arg, eps=1e-8):
        return S.sqrt(arg + eps)

    @classmethod
    def smooth_norm(cls, a, eps=1e-8):
        return cls.smooth_sqrt(S.dot(a, a), eps ** 2.)

    @classmethod
    def norm(cls, a):
        return cls.smooth_norm(a, 0.)

    @classmethod
    def smooth_abs(cls, a, eps=1e-8):
        a = S._e(a)
        return cls.smooth_sqrt(a * a, eps ** 2.)

    # ---- environment (awebox/mdl/atmosphere.py:42-123, awebox/mdl/wind.py:50-86, 184-208)
    def density(self, zz):
        model = self.spec.opt('atmosphere.model', 'isa')
        if model == 'isa':
            t = self.par('atmosphere', 't_ref') - self.par('atmosphere', 'gamma_air') * zz
            expo = self.par('atmosphere', 'g') / self.par('atmosphere', 'gamma_air') / self.par('atmosphere', 'r') - 1.0
            return self.par('atmosphere', 'rho_ref') * S.power(t / self.par('atmosphere', 't_ref'), expo)
        if model in ('uniform', 'log_wind'):
            return self.par('atmosphere', 'rho_ref')
        raise NotImplementedError('atmosphere model %s' % model)

    def wind_velocity(self, zz):
        model = self.spec.opt('wind.model')
        u_ref = self.par('wind', 'u_ref')
        z_cropped = self.smooth_abs(zz, 1.)
        if model == 'power':
            u = u_ref * S.power(z_cropped / self.par('wind', 'z_ref'), self.par('wind', 'power_wind', 'exp_ref'))
        elif model == 'log_wind':
            k = 1.0 / math.log(10.0)
            z0 = self.par('wind', 'log_wind', 'z0_air')
            u = u_ref * (S.log(z_cropped / z0) * k) / (S.log(self.par('wind', 'z_ref') / z0) * k)
        elif model == 'uniform':
            u = u_ref
        else:
            raise NotImplementedError('wind model %s' % model)
        return S.vec([u, S.ZERO, S.ZERO])

    # ---- tether segments (awebox/mdl/aero/tether_dir/tether_aero.py:178-267)
    def segment(self, upper):
        if upper in self._seg_cache:
            return self._seg_cache[upper]
        arch = self.arch
        lower = arch.parent_map[upper]
        main = lower == 0
        secondary = upper in arch.kite_nodes
        if main:
            length_sym, diam_sym = 'l_t', 'diam_t'
        elif secondary:
            length_sym, diam_sym = 'l_s', 'diam_s'
        else:
            length_sym, diam_sym = 'l_i', 'diam_t'
        lt_type = self.var_type_of(length_sym)
        d_type = self.var_type_of(diam_sym)
        q_node = self.si('x', 'q' + arch.node_label(upper))
        q_parent = S.zeros(3) if main else self.si('x', 'q' + arch.node_label(lower))
        p = {}
        p['seg_length'] = self.norm(q_node - q_parent)
        p['scaling_length'] = float(self.scale(lt_type, length_sym)[0])
        p['seg_diam'] = self.si(d_type, diam_sym)[0]
        p['scaling_diam'] = float(self.scale(d_type, diam_sym)[0])
        p['cross_section_area'] = np.pi * S.power(p['seg_diam'] / 2., 2.)
        p['scaling_area'] = np.pi * (p['scaling_diam'] / 2.) ** 2.
        density = self.par('tether', 'rho')
        p['seg_mass'] = p['cross_section_area'] * density * p['seg_length']
        p['scaling_mass'] = p['scaling_area'] * density * p['scaling_length']
        if self.has(lt_type, 'd' + length_sym):
            p['scaling_speed'] = S.const(self.scale(lt_type, 'd' + length_sym)[0])
        else:
            p['scaling_speed'] = (1. / 3.) * self.par('wind', 'u_ref')
        acc_type = None
        for vt in ['x', 'u']:
            if self.has(vt, 'ddl_t'):
                acc_type = vt
        if acc_type is None:
            p['scaling_acc'] = float(np.max(self.spec.opt('system_bounds.x.ddl_t'))) / 2.
        else:
            p['scaling_acc'] = float(self.scale(acc_type, 'ddl_t')[0])
        self._seg_cache[upper] = p
        return p

    # ---- total time derivative (awebox/mdl/lagr_dyn_dir/tools.py:13-73)
    def time_derivative(self, expr):
        arch = self.arch
        expr = S._e(expr)
        deriv = S.ZERO
        rdot = set('dr' + arch.node_label(k) for k in arch.kite_nodes)
        for deriv_name in self.sorted_names('xdot'):
            if deriv_name in rdot:
                continue
            var_name = deriv_name[1:]
            var_type = self.var_type_of(var_name)
            deriv_type = self.var_type_of(deriv_name)
            q_sym = self.scaled(var_type, var_name)
            dq_sym = self.scaled(deriv_type, deriv_name)
            s_q, s_dq = self.scale(var_type, var_name), self.scale(deriv_type, deriv_name)
            row = S.jacobian([expr], q_sym)[0]
            rescale = not np.all(s_q - s_dq == 0.0)
            for j, dfdq in row.items():
                if rescale:
                    deriv = deriv + (dfdq * (1.0 / s_q[j])) * (S.const(s_dq[j]) * dq_sym[j])
                else:
                    deriv = deriv + dfdq * dq_sym[j]
        for kite in arch.kite_nodes:
            lab = arch.node_label(kite)
            if not self.has('x', 'r' + lab):
                continue
            r_kite = self.scaled('x', 'r' + lab)
            row = S.jacobian([expr], r_kite)[0]
            if not row:
                continue
            dcm = r_kite.reshape(3, 3).T            # column-major reshape of the 9-vector
            omega = self.scaled('x', 'omega' + lab)   # (sic) the reference uses the scaled omega here
            dr_dt = S.matmul(S.skew(omega), inv3(dcm.T)).T.reshape(-1)
            for j, dfdr in row.items():
                deriv = deriv + dfdr * dr_dt[j]
        return deriv

    # ------------------------------------------------------------------ assembly
    def _build(self):
        spec, arch = self.spec, self.arch
        dof = int(spec.opt('kite_dof'))
        nn = arch.number_of_nodes
        g_acc = self.par('atmosphere', 'g')

        # --- energies (awebox/mdl/lagr_dyn_dir/energy.py:59-144)
        e_kin, e_pot = S.ZERO, S.ZERO
        for node in range(1, nn):
            lab = arch.node_label(node)
            seg = self.segment(node)
            q_node, dq_node = self.si('x', 'q' + lab), self.si('x', 'dq' + lab)
            if node == 1:
                ehat = q_node / self.norm(q_node)
                reelout = S.dot(self.si('x', 'dq10'), q_node / self.norm(q_node))
                dq_parent = reelout * ehat
                q_parent = S.zeros(3)
            else:
                dq_parent = self.si('x', 'dq' + arch.parent_label(node))
                q_parent = self.si('x', 'q' + arch.parent_label(node))
            if node == 1 and SIMPLIFY_REELOUT:
                # dq_parent = reelout * ehat with |ehat| = 1:  dq_parent.dq_parent = dq_node.dq_parent = reelout^2
                e_kin = e_kin + 0.5 * seg['seg_mass'] / 3 * (S.dot(dq_node, dq_node) + 2. * (reelout * reelout))
            else:
                e_kin = e_kin + 0.5 * seg['seg_mass'] / 3 * (S.dot(dq_node, dq_node) + S.dot(dq_parent, dq_parent)
                                                           + S.dot(dq_node, dq_parent))
            if node in arch.kite_nodes:
                m_k = self.par('geometry', 'm_k')
                e_kin = e_kin + 0.5 * m_k * S.dot(dq_node, dq_node)
                if dof == 6:
                    omega = self.si('x', 'omega' + lab)
                    J = self.par('geometry', 'j').reshape(3, 3).T
                    e_kin = e_kin + 0.5 * S.dot(S.matvec(J.T, omega), omega)
                e_pot = e_pot + g_acc * m_k * q_node[2]
            q_mean = (q_node + q_parent) / 2.
            e_pot = e_pot + g_acc * seg['seg_mass'] * q_mean[2]

        # --- holonomic constraints (awebox/mdl/lagr_dyn_dir/holonomics.py:17-122, 204-273)
        attachment = spec.opt('tether.attachment', 'com')
        if attachment not in ('com', 'stick'):
            raise ValueError('unknown tether attachment %s' % attachment)
        g_list, gdot_list, gddot_list = [], [], []
        work = S.ZERO
        for node in range(1, nn):
            lab = arch.node_label(node)
            q_si = self.si('x', 'q' + lab)
            arm = (node in arch.kite_nodes) and attachment == 'stick'
            if arm:
                if dof != 6:
                    raise ValueError('stick attachment needs 6-DOF kites')
                dcm = self.si('x', 'r' + lab).reshape(3, 3).T
                current = q_si + S.matvec(dcm, self.par('geometry', 'r_tether'))
            else:
                current = q_si
            if node == 1:
                previous = S.zeros(3)
                length = self.si(self.var_type_of('l_t'), 'l_t')[0]
            elif node in arch.kite_nodes:
                previous = self.si('x', 'q' + arch.parent_label(node))
                length = self.si('theta', 'l_s')[0]
            else:
                previous = self.si('x', 'q' + arch.parent_label(node))
                length = self.si('theta', 'l_i')[0]
            seg_vec = current - previous
            c = 0.5 * (S.dot(seg_vec, seg_vec) - S.power(length, 2.0))
            dc = self.time_derivative(c)
            ddc = self.time_derivative(dc)
            g_list.append(c), gdot_list.append(dc), gddot_list.append(ddc)
            work = work + self.si('z', 'lambda' + lab)[0] * c
        self.holonomic = (g_list, gdot_list, gddot_list)
        self.work_holonomic = work

        lag = e_kin - e_pot - work

        # --- generalized forces (awebox/mdl/lagr_dyn_dir/forces.py:47-189)
        forces = {arch.node_label(n): S.zeros(3) for n in range(1, nn)}
        moments = {}
        drag_model = spec.opt('tether.tether_drag.model_type', 'multi')
        for node in range(1, nn):
            lab = arch.node_label(node)
            if drag_model == 'multi':
                lower, upper = self.segment_drag(node, int(spec.opt('tether.aero_elements')))
            elif drag_model == 'split':
                lower, upper = self.segment_drag(node, 1)
            elif drag_model == 'not_in_use':
                lower, upper = S.zeros(3), S.zeros(3)
            else:
                raise NotImplementedError('tether drag model %s' % drag_model)
            if node > 1:
                forces[arch.parent_label(node)] = forces[arch.parent_label(node)] + lower
            forces[lab] = forces[lab] + upper
        self.aero = {}
        gamma = self.phi('gamma')
        for kite in arch.kite_nodes:
            lab = arch.node_label(kite)
            if dof == 6:
                with S.region('aero%d' % kite):
                    f_earth, m_body = self.six_dof_aero(kite)
                moments[lab] = gamma * self.si('u', 'm_fict' + lab) + m_body
            else:
                f_earth = self.three_dof_aero(kite)
            forces[lab] = forces[lab] + (gamma * self.si('u', 'f_fict' + lab) + f_earth)

        # --- translational dynamics (awebox/mdl/lagr_dyn_dir/lagr_dyn.py:62-109)
        gc_q = np.concatenate([self.scaled('x', n) for n in self.gc])
        gc_dq = np.concatenate([self.scaled('x', 'd' + n) for n in self.gc])
        s_q = np.concatenate([self.scale('x', n) for n in self.gc])
        s_dq = np.concatenate([self.scale('x', 'd' + n) for n in self.gc])
        dl_dqdot = S.jacobian([lag], gc_dq)[0]
        dl_dq = S.jacobian([lag], gc_q)[0]
        lhs = []
        for i in range(len(gc_q)):
            t1 = self.time_derivative(dl_dqdot.get(i, S.ZERO)) * (1.0 / s_dq[i])
            t2 = dl_dq.get(i, S.ZERO) * (1.0 / s_q[i])
            lhs.append(t1 - t2)
        rhs = np.concatenate([forces[arch.node_label(n)] for n in range(1, nn)])
        # momentum correction of the open system (lagr_dyn.py:174-204)
        mass_flow = self.time_derivative(self.segment(1)['seg_mass'])
        velocity = self.si('x', 'dq10')
        first = self.gc.index('q10') * 3
        rhs = rhs.copy()
        for i in range(3):
            rhs[first + i] = rhs[first + i] + mass_flow * velocity[i]
        # characteristic forces (awebox/mdl/lagr_dyn_dir/mass.py:52-82)
        trans = []
        k = 0
        for node in range(1, nn):
            node_mass = self.segment(node)['scaling_mass'] / 2.
            for child in arch.children_map.get(node, []):
                node_mass = node_mass + self.segment(child)['scaling_mass'] / 2.
            if node in arch.kite_nodes:
                node_mass = node_mass + self.par('geometry', 'm_k')
            force_scaling = node_mass * float(spec.opt('scaling.other.g')) * 10.
            for i in range(3):
                trans.append((lhs[k] - rhs[k]) / force_scaling)
                k += 1
        self.eq.append(('dynamics_translation', trans))

        # --- holonomic rows (lagr_dyn.py:116-125, holonomics.py:267-300)
        kappa = self.par('tether', 'kappa')
        rows = []
        for i, node in enumerate(range(1, nn)):
            seg = self.segment(node)
            q_scaling_mean = float(np.sum(self.scale('x', 'q' + arch.node_label(node)))) / 3.
            g_loc = seg['scaling_length'] * q_scaling_mean
            scaling = S.ZERO + 2. * kappa * 0. + S.power(kappa, 2.) * g_loc
            lhs_c = gddot_list[i] + 2. * kappa * gdot_list[i] + S.power(kappa, 2.) * g_list[i]
            rows.append(lhs_c / scaling)
        self.eq.append(('dynamics_constraint', rows))

        # --- rotational dynamics (lagr_dyn.py:207-254)
        if dof == 6:
            J = self.par('geometry', 'j').reshape(3, 3).T
            m_scale = float(spec.opt('scaling.z.m_aero'))
            kappa_r = self.par('kappa_r')
            for kite in arch.kite_nodes:
                lab = arch.node_label(kite)
                R = self.si('x', 'r' + lab).reshape(3, 3).T
                dR = self.si('xdot', 'dr' + lab).reshape(3, 3).T
                omega = self.si('x', 'omega' + lab)
                domega = self.si('xdot', 'domega' + lab)
                tether_moment = self.tether_moment(work, kite)
                J_domega = S.matvec(J, domega)
                w_x_Jw = S.cross(omega, S.matvec(J, omega))
                rot = [(moments[lab][i] - (J_domega[i] + w_x_Jw[i] + tether_moment[i])) / m_scale for i in range(3)]
                self.eq.append(('rotation_dynamics%d' % kite, rot))
                RtR = S.matmul(R.T, R)
                ortho = S.zeros(3, 3)
                for i in range(3):
                    for j in range(3):
                        ortho[i, j] = kappa_r / 2. * ((1.0 if i == j else 0.0) - RtR[i, j])
                M = dR - S.matmul(R, ortho + S.skew(omega))
                self.eq.append(('ref_frame_dynamics%d' % kite, list(M.T.reshape(-1))))

        # --- trivial kinematics (lagr_dyn.py:141-169)
        for name in self.sorted_names('xdot'):
            if self.has('x', name):
                ut = 'x'
            elif self.has('u', name):
                ut = 'u'
            else:
                continue
            diff = self.si('xdot', name) - self.si(ut, name)
            mean = np.sqrt(self.scale(ut, name) * self.scale('xdot', name))
            self.eq.append(('trivial_' + name, [diff[i] / mean[i] for i in range(len(diff))]))

        if not spec.opt('integral_outputs', True):
            # alongside integration of the energy state (awebox/mdl/dynamics.py:233-262)
            power = self.si('z', 'lambda10')[0] * self.si('x', 'l_t')[0] * self.si('x', 'dl_t')[0]
            e_scaling = float(spec.opt('scaling.x.e'))
            self.eq.append(('integral_e', [self.scaled('xdot', 'de')[0] - power / e_scaling]))

        self._build_inequalities()

    # ---- tether moment (forces.py:174-189, vector_operations.py:236-262)
    def tether_moment(self, work, kite):
        lab = self.arch.node_label(kite)
        R = self.si('x', 'r' + lab).reshape(3, 3).T
        row = S.jacobian([work], self.scaled('x', 'r' + lab))[0]
        if not row:
            return S.zeros(3)
        A = S.zeros(9)
        for j, v in row.items():
            A[j] = v
        A = A.reshape(3, 3).T
        RtA = S.matmul(R.T, A)
        unskew = S.vec([0.5 * (RtA[2, 1] - RtA[1, 2]), 0.5 * (RtA[0, 2] - RtA[2, 0]), 0.5 * (RtA[1, 0] - RtA[0, 1])])
        return S.vec([2 * unskew[i] for i in range(3)])

    # ---- tether drag (tether_aero.py:121-174, segment.py:38-66, element.py:38-146)
    def segment_drag(self, upper, n_elements):
        arch = self.arch
        lower = arch.parent_map[upper]
        q_top, dq_top = self.si('x', 'q' + arch.node_label(upper)), self.si('x', 'dq' + arch.node_label(upper))
        if lower == 0:
            q_bot, dq_bot = S.zeros(3), S.zeros(3)
        else:
            q_bot, dq_bot = self.si('x', 'q' + arch.node_label(lower)), self.si('x', 'dq' + arch.node_label(lower))
        if lower == 0 or upper not in arch.kite_nodes:
            diam = self.si('theta', 'diam_t')[0]
        else:
            diam = self.si('theta', 'diam_s')[0]
        if self.spec.opt('tether.cd_model', 'constant') != 'constant':
            raise NotImplementedError('only the constant tether cd model is supported')
        cd = self.par('tether', 'cd')
        ds = 1.0 / n_elements
        s_grid = np.linspace(0.5 * ds, 1 - 0.5 * ds, n_elements)
        f_up, f_low = S.zeros(3), S.zeros(3)
        for e in range(n_elements):
            with S.region('tether%d_elem%d' % (upper, e)):
                lo_phi, up_phi = float(e) / float(n_elements), float(e + 1) / float(n_elements)
                q_l = q_bot + (q_top - q_bot) * lo_phi
                q_u = q_bot + (q_top - q_bot) * up_phi
                dq_l = dq_bot + (dq_top - dq_bot) * lo_phi
                dq_u = dq_bot + (dq_top - dq_bot) * up_phi
                zz = ((q_u + q_l) / 2.)[2]
                ua = self.wind_velocity(zz) - (dq_u + dq_l) / 2.
                eps = 1.e-6
                ua_norm = self.smooth_norm(ua, eps)
                ehat = ua / ua_norm
                tether = q_u - q_l
                length_sq = S.dot(tether, tether)
                length_par = S.dot(tether, ehat)
                length_perp = self.smooth_sqrt(length_sq - S.power(length_par, 2.), eps ** 2.)
                drag_mag = cd * 0.5 * self.density(zz) * ua_norm * diam * length_perp
                drag = S.vec([drag_mag * ua[i] for i in range(3)])
                f_up = f_up + float(s_grid[e]) * drag
                f_low = f_low + float(1 - s_grid[e]) * drag
        return f_low, f_up

    # ---- 6-DOF kite aerodynamics (six_dof_kite.py:52-84,165-202; stability_derivatives.py:41-228;
    #      indicators.py:435-463; frames.py:59-71; kite_dir/tools.py:121-230)
    def six_dof_aero(self, kite):
        spec, lab = self.spec, self.arch.node_label(kite)
        if spec.opt('aero.stab_derivs.force_frame') != 'control' or spec.opt('aero.stab_derivs.moment_frame') != 'control':
            raise NotImplementedError('stability derivatives are supported in the control frame only')
        if spec.opt('aero.aero_coeff_ref_velocity', 'eff') != 'eff':
            raise NotImplementedError("aero_coeff_ref_velocity must be 'eff' without induction")
        q, dq = self.si('x', 'q' + lab), self.si('x', 'dq' + lab)
        R = self.si('x', 'r' + lab).reshape(3, 3).T
        vec_u = self.wind_velocity(q[2]) - dq
        if int(spec.opt('surface_control', 1)) == 0:
            delta = self.si('u', 'delta' + lab)
        else:
            delta = self.si('x', 'delta' + lab)
        omega = self.si('x', 'omega' + lab)
        rho = self.density(q[2])
        e1, e2, e3 = R[:, 0], R[:, 1], R[:, 2]
        x_comp = self.smooth_abs(S.dot(vec_u, e1))
        alpha = S.dot(vec_u, e3) / x_comp
        beta = S.dot(vec_u, e2) / x_comp
        airspeed = self.norm(vec_u)
        b_ref, c_ref, s_ref = self.par('geometry', 'b_ref'), self.par('geometry', 'c_ref'), self.par('geometry', 's_ref')
        om_c = [-omega[0], omega[1], -omega[2]]
        om_hat = [o / (2. * airspeed) for o in om_c]
        inputs = {'0': S.ONE, 'alpha': alpha, 'beta': -beta, 'p': om_hat[0] * b_ref, 'q': om_hat[1] * c_ref,
                  'r': om_hat[2] * b_ref, 'deltaa': delta[0], 'deltae': delta[1], 'deltar': delta[2]}
        coeffs = {}
        for cname in ['CX', 'CY', 'CZ', 'Cl', 'Cm', 'Cn']:
            total = S.ZERO
            for iname in ['0', 'alpha', 'beta', 'p', 'q', 'r', 'deltaa', 'deltae', 'deltar']:
                key = ('theta0', 'aero', cname, iname)
                if key not in self._pidx:
                    continue
                stack = self.params[self._pidx[key]]
                # sum_l stack[l] * (input * alpha^l), evaluated as input * Horner(alpha)
                poly = stack[len(stack) - 1]
                for l in range(len(stack) - 2, -1, -1):
                    poly = poly * alpha + stack[l]
                contrib = poly if iname == '0' else inputs[iname] * poly
                if cname in ('Cl', 'Cm', 'Cn') and iname in ('deltaa', 'deltae', 'deltar'):
                    contrib = self.par('aero', 'moment_factor') * contrib
                total = total + contrib
            coeffs[cname] = total
        dyn_p = 1. / 2. * rho * S.dot(vec_u, vec_u)
        f_ctrl = [coeffs[c] * dyn_p * s_ref for c in ('CX', 'CY', 'CZ')]
        ref_len = [b_ref, c_ref, b_ref]
        m_ctrl = [dyn_p * s_ref * (ref_len[i] * coeffs[c]) for i, c in enumerate(('Cl', 'Cm', 'Cn'))]
        f_body = S.vec([-f_ctrl[0], f_ctrl[1], -f_ctrl[2]])
        m_body = S.vec([-m_ctrl[0], m_ctrl[1], -m_ctrl[2]])
        f_earth = S.matvec(R, f_body)
        self.aero[kite] = {'vec_u': vec_u, 'dcm': R, 'alpha': alpha, 'beta': beta, 'f_body': f_body, 'm_body': m_body,
                           'f_earth': f_earth, 'f_control': S.vec(f_ctrl), 'rho': rho}
        return f_earth, m_body

    # ---- 3-DOF (point-mass) kite aerodynamics (three_dof_kite.py:47-64, 100-137, 171-218;
    #      kite_dir/tools.py:142-177): lift along the roll-rotated normal of the (tether, apparent wind)
    #      plane, drag along the apparent wind, CL and roll angle psi are the states coeff[0], coeff[1]
    def three_dof_aero(self, kite):
        spec, arch, lab = self.spec, self.arch, self.arch.node_label(kite)
        if spec.opt('aero.aero_coeff_ref_velocity', 'eff') != 'eff':
            raise NotImplementedError("aero_coeff_ref_velocity must be 'eff' without induction")
        q, dq = self.si('x', 'q' + lab), self.si('x', 'dq' + lab)
        coeff = self.si('x', 'coeff' + lab)
        vec_u = self.wind_velocity(q[2]) - dq
        parent = arch.parent_map[kite]
        vec_t = q - (S.zeros(3) if parent == 0 else self.si('x', 'q' + arch.node_label(parent)))
        vec_v = S.cross(vec_t, vec_u)
        vec_w = S.cross(vec_u, vec_v)
        uhat = vec_u / self.smooth_norm(vec_u)
        vhat = vec_v / self.smooth_norm(vec_v)
        what = vec_w / self.smooth_norm(vec_w)
        psi = coeff[1]
        cpsi, spsi = S.cos(psi), S.sin(psi)
        e1 = uhat
        e2 = cpsi * vhat + spsi * what
        e3 = cpsi * what - spsi * vhat
        dcm = S.zeros(3, 3)
        for i in range(3):
            dcm[i, 0], dcm[i, 1], dcm[i, 2] = e1[i], e2[i], e3[i]
        rho = self.density(q[2])
        cl = coeff[0]
        cd0 = S.ZERO
        for label in ('CX', 'CA', 'CD'):
            key = ('theta0', 'aero', label, '0')
            if key in self._pidx:
                cd0 = self.smooth_abs(self.params[self._pidx[key][0]], 0.)
        cd = cd0 + S.power(cl, 2.) / (np.pi * self.par('geometry', 'ar'))
        s_ref = self.par('geometry', 's_ref')
        lift_mag = cl * 1. / 2. * rho * S.dot(vec_u, vec_u) * s_ref
        drag_mag = cd * 1. / 2. * rho * self.norm(vec_u) * s_ref
        f_earth = S.vec([lift_mag * e3[i] + drag_mag * vec_u[i] for i in range(3)])
        self.aero[kite] = {'vec_u': vec_u, 'dcm': dcm, 'f_earth': f_earth, 'rho': rho}
        return f_earth

    # ---- path inequalities (awebox/mdl/dynamics.py:457-484, 627-821, 1022-1117)
    def _build_inequalities(self):
        spec, arch = self.spec, self.arch
        nn = arch.number_of_nodes
        def node_list(v):
            # a one-element list comes out of the spec file as a bare number
            return [int(n) for n in (v if isinstance(v, (list, tuple)) else ([] if v is None else [v]))]

        includes_force = node_list(spec.opt('model_bounds.tether.tether_constraint_includes.force'))
        includes_stress = node_list(spec.opt('model_bounds.tether.tether_constraint_includes.stress'))
        tightness = float(spec.opt('model_bounds.tether_stress.scaling', 1.0))
        for node in range(1, nn):
            lab = arch.node_label(node)
            seg = self.segment(node)
            tension = self.si('z', 'lambda' + lab)[0] * seg['seg_length']
            lam_scaling = float(self.scale('z', 'lambda' + lab)[0])
            if node in includes_stress:
                max_stress = self.par('tether', 'max_stress') / self.par('tether', 'stress_safety_factor')
                char_tension = math.sqrt((lam_scaling * seg['scaling_length']) ** 2. + 1e-16)
                expr = (tension - seg['cross_section_area'] * max_stress) / char_tension * tightness
                self.ineq.append(('tether_stress' + lab, [expr]))
            elif node in includes_force:
                lims = self.par('model_bounds', 'tether_force_limits')
                fs = lam_scaling * seg['scaling_length']
                self.ineq.append(('tether_force_max' + lab, [(tension - lims[1]) / fs]))
                self.ineq.append(('tether_force_min' + lab, [(lims[0] - tension) / fs]))
        if spec.opt('model_bounds.airspeed.include'):
            lims = self.par('model_bounds', 'airspeed_limits')
            u_ref = self.par('wind', 'u_ref')
            for kite in arch.kite_nodes:
                lab = arch.node_label(kite)
                airspeed = self.smooth_norm(self.aero[kite]['vec_u'])
                self.ineq.append(('airspeed_max' + lab, [(airspeed - lims[1]) / u_ref]))
                self.ineq.append(('airspeed_min' + lab, [(lims[0] - airspeed) / u_ref]))
        if spec.opt('model_bounds.aero_validity.include'):
            tight = float(spec.opt('model_bounds.aero_validity.scaling', 1.0))
            v_ref = float(spec.opt('model_bounds.aero_validity.airspeed_ref'))
            rad = np.pi / 180.0
            a_min, a_max = spec.opt('aero.alpha_min_deg') * rad, spec.opt('aero.alpha_max_deg') * rad
            b_min, b_max = spec.opt('aero.beta_min_deg') * rad, spec.opt('aero.beta_max_deg') * rad
            sabs = lambda v: math.sqrt(v ** 2. + 1e-16)
            for kite in arch.kite_nodes:
                ua, R = self.aero[kite]['vec_u'], self.aero[kite]['dcm']
                u1, u2, u3 = S.dot(ua, R[:, 0]), S.dot(ua, R[:, 1]), S.dot(ua, R[:, 2])
                self.ineq.append(('alpha_ub%d' % kite, [(u3 - u1 * a_max) * tight / v_ref / sabs(a_max)]))
                self.ineq.append(('alpha_lb%d' % kite, [(-u3 + u1 * a_min) * tight / v_ref / sabs(a_min)]))
                self.ineq.append(('beta_ub%d' % kite, [(u2 - u1 * b_max) * tight / v_ref / sabs(b_max)]))
                self.ineq.append(('beta_lb%d' % kite, [(-u2 + u1 * b_min) * tight / v_ref / sabs(b_min)]))
        if spec.opt('model_bounds.anticollision.include'):
            import itertools
            dist_min = float(spec.opt('model_bounds.anticollision.safety_factor')) * self.par('geometry', 'b_ref')
            for ka, kb in itertools.combinations(arch.kite_nodes, 2):
                dist = self.si('x', 'q' + arch.node_label(ka)) - self.si('x', 'q' + arch.node_label(kb))
                self.ineq.append(('anticollision%d%d' % (ka, kb), [1 - (S.dot(dist, dist) / S.power(dist_min, 2.))]))
        if spec.opt('model_bounds.acceleration.include'):
            acc_max = float(spec.opt('model_bounds.acceleration.acc_max')) * float(spec.opt('scaling.other.g'))
            for name in self.sorted_names('xdot'):
                if name.startswith('ddq') and name[3:].isdigit():
                    acc = self.si('xdot', name)
                    self.ineq.append(('acceleration' + name[3:], [S.dot(acc, acc) / acc_max ** 2. - 1.]))
        if int(spec.opt('kite_dof')) == 6 and spec.opt('model_bounds.rotation.include'):
            if spec.opt('model_bounds.rotation.type') != 'yaw':
                raise NotImplementedError('only the yaw rotation bound is supported')
            gamma_max = self.par('model_bounds', 'rot_angles')[2]
            for kite in arch.kite_nodes:
                lab = arch.node_label(kite)
                parent = arch.parent_map[kite]
                q0 = self.si('x', 'q' + lab)
                q1 = S.zeros(3) if parent == 0 else self.si('x', 'q' + arch.node_label(parent))
                q_hat = q0 - q1
                R = self.si('x', 'r' + lab).reshape(3, 3).T
                yaw = S.dot(q_hat, R[:, 2]) - S.cos(gamma_max) * self.norm(q_hat)
                scale = float(spec.opt('scaling.x.l_t')) if kite == 1 else float(spec.opt('scaling.theta.l_s'))
                self.ineq.append(('rotation_max' + lab, [-1. * (yaw / scale)]))
        if spec.opt('model_bounds.ellipsoidal_flight_region.include'):
            raise NotImplementedError('ellipsoidal flight region')

    # ---- flat views
    def eq_exprs(self):
        return [e for _, rows in self.eq for e in rows]

    def ineq_exprs(self):
        return [e for _, rows in self.ineq for e in rows]

    def eq_names(self):
        return [n for n, rows in self.eq for _ in rows]

    def ineq_names(self):
        return [n for n, rows in self.ineq for _ in rows]


def inv3(A):
    """inverse of a 3x3 expression matrix by the adjugate (what ``cas.inv`` yields for a dense
    symbolic 3x3; awebox/mdl/lagr_dyn_dir/tools.py:71, awebox/mdl/aero/kite_dir/frames.py:55)."""
    c = S.zeros(3, 3)
    for i in range(3):
        for j in range(3):
            r = [k for k in range(3) if k != i]
            s = [k for k in range(3) if k != j]
            minor = A[r[0], s[0]] * A[r[1], s[1]] - A[r[0], s[1]] * A[r[1], s[0]]
            c[j, i] = minor if (i + j) % 2 == 0 else -minor
    det = A[0, 0] * c[0, 0] + A[0, 1] * c[1, 0] + A[0, 2] * c[2, 0]
    out = S.zeros(3, 3)
    for i in range(3):
        for j in range(3):
            out[i, j] = c[i, j] / det
    return out
