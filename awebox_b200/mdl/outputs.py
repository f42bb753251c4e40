"""Model outputs: the diagnostic expressions awebox collects next to the dynamics.

Restates, on the expression graph of ``NodeModel``, the entries of ``model.outputs`` (awebox/mdl/dynamics.py:40-207):

    masses              lagr_dyn_dir/mass.py:45-56
    e_kinetic/potential lagr_dyn_dir/energy.py:42-144
    invariants          lagr_dyn_dir/holonomics.py:17-68
    rotation            aero/geometry_dir/unit_normal.py:44-71, 171-187 (normal vector models tether_parallel, xhat)
    aerodynamics        aero/kite_dir/kite_aero.py:125-218, aero/indicators.py:44-118, 147-236, 278-336
    environment         aero/indicators.py:368-381
    aero_validity       aero/indicators.py:278-336
    local_performance   aero/indicators.py:338-366, 384-404, dynamics.py:706-790 (tether stress / force), 1084-1117
    power_balance       aero/indicators.py:238-276, dynamics.py:341-425
    geometry            aero/geometry_dir/geometry.py:106-207 ('averaged' and 'parent' centres; the 'frenet' centre of a
                        single kite at the main tether is not restated: those entries are absent)
    performance         aero/indicators.py:117-152, 406-433, tools/performance_operations.py:38-74
    tether_aero         aero/tether_dir/tether_aero.py:121-174, segment.py:38-100
    tether_moments      lagr_dyn_dir/forces.py:174-189
    model_inequalities / model_equalities / xdot_from_var   dynamics.py:158-170, 427-429

The reference reads outputs BY LABEL ('[group,name,i]'); the index order of ``model.outputs`` follows dictionary and
*set* iteration (indicators.py:171 iterates ``set(...)``), i.e. it is not fixed across Python processes.  The order
here is the reference's insertion order with the set sorted; the label table travels with the values.
"""
import math

import numpy as np

from .. import sym as S
from .model import inv3


def _vec3(v):
    return [v[0], v[1], v[2]]


def _normalize(m, a, eps=0.):
    n = m.smooth_norm(a, eps)
    return S.vec([a[i] / n for i in range(3)])


def _normed_cross(m, a, b, eps=0.):
    return _normalize(m, S.cross(a, b), eps)


def _wind_dcm(m, vec_u, dcm):
    """frames.py:38-47: columns (Dhat, Shat, Lhat)"""
    span = dcm[:, 1]
    dhat = _normalize(m, vec_u, 1e-8)
    lhat = _normed_cross(m, vec_u, span, 1e-8)
    shat = _normed_cross(m, lhat, dhat, 1e-8)
    W = S.zeros(3, 3)
    for i in range(3):
        W[i, 0], W[i, 1], W[i, 2] = dhat[i], shat[i], lhat[i]
    return W


def _power_density(m, zz):
    u = m.wind_velocity(zz)
    return .5 * m.density(zz) * S.power(m.norm(u), 3.)


def _temperature(m, zz):
    model = m.spec.opt('atmosphere.model', 'isa')
    if model == 'isa':
        return m.par('atmosphere', 't_ref') - m.par('atmosphere', 'gamma_air') * zz
    return m.par('atmosphere', 't_ref')


def _pressure(m, zz):
    model = m.spec.opt('atmosphere.model', 'isa')
    if model == 'isa':
        return m.density(zz) * m.par('atmosphere', 'r') * _temperature(m, zz)
    return m.par('atmosphere', 'p_ref')


def _viscosity(m, zz):
    model = m.spec.opt('atmosphere.model', 'isa')
    if model == 'isa':
        t = _temperature(m, zz)
        t_ref, c_s = m.par('atmosphere', 't_ref'), m.par('atmosphere', 'c_sutherland')
        return m.par('atmosphere', 'mu_ref') * (t_ref + c_s) / (t + c_s) * S.power(t / t_ref, 3.0 / 2.0)
    return m.par('atmosphere', 'mu_ref')


def _elevation(q10):
    ground = S.power(S.power(q10[0], 2.) + S.power(q10[1], 2.), 0.5)
    return S.atan2(q10[2], ground)


def _loyd_phf(cl, cd, elevation):
    eps = 1.e-6
    interior = S.power(cd, 2.) / (S.power(cl, 2) + eps ** 2.)
    cr = cl * S.power(1. + interior, 0.5)
    return 4. / 27. * cr * S.power(cr / cd, 2.) * S.power(S.cos(elevation), 3.)


class _Outputs(object):
    def __init__(self):
        self.groups = {}          # group -> {name: [Expr]} in insertion order
        self.order = []

    def put(self, group, name, value):
        if group not in self.groups:
            self.groups[group] = {}
            self.order.append(group)
        if isinstance(value, np.ndarray):
            value = list(value.reshape(-1))
        elif not isinstance(value, (list, tuple)):
            value = [value]
        self.groups[group][name] = [S._e(v) for v in value]

    def get(self, group, name):
        return self.groups[group][name]

    def touch(self, group):
        if group not in self.groups:
            self.groups[group] = {}
            self.order.append(group)

    def flat(self):
        out = []
        for g in self.order:
            for name, vals in self.groups[g].items():
                for i, v in enumerate(vals):
                    out.append(('[%s,%s,%d]' % (g, name, i), v))
        return out


def model_outputs(m):
    """-> [(label, Expr)] over the node variables ``m.vars`` and parameters ``m.params``"""
    spec, arch = m.spec, m.arch
    dof = int(spec.opt('kite_dof'))
    nn = arch.number_of_nodes
    o = _Outputs()
    g_acc = m.par('atmosphere', 'g')
    layer_nodes = sorted(set(arch.parent_map[k] for k in arch.kite_nodes))
    kites_of = {p: [k for k in arch.kite_nodes if arch.parent_map[k] == p] for p in layer_nodes}

    def q_of(node):
        return S.zeros(3) if node == 0 else m.si('x', 'q' + arch.node_label(node))

    def dq_of(node):
        return S.zeros(3) if node == 0 else m.si('x', 'dq' + arch.node_label(node))

    # ---- masses, energies
    for node in range(1, nn):
        o.put('masses', 'm_tether%d' % node, m.segment(node)['seg_mass'])
    for node in range(1, nn):
        lab = arch.node_label(node)
        seg_mass = m.segment(node)['seg_mass']
        q_node, dq_node = q_of(node), dq_of(node)
        if node == 1:
            ehat = _normalize(m, q_node)
            reelout = S.dot(m.si('x', 'dq10'), _normalize(m, m.si('x', 'q10')))
            dq_parent = S.vec([reelout * ehat[i] for i in range(3)])
        else:
            dq_parent = dq_of(arch.parent_map[node])
        o.put('e_kinetic', 'tether' + lab, 0.5 * seg_mass / 3 * (S.dot(dq_node, dq_node) + S.dot(dq_parent, dq_parent)
                                                                 + S.dot(dq_node, dq_parent)))
        is_kite = node in arch.kite_nodes
        o.put('e_kinetic', 'kite_trans' + lab,
              0.5 * m.par('geometry', 'm_k') * S.dot(dq_node, dq_node) if is_kite else S.ZERO)
        if is_kite and dof == 6:
            omega = m.si('x', 'omega' + lab)
            J = m.par('geometry', 'j').reshape(3, 3).T
            o.put('e_kinetic', 'kite_rot' + lab, 0.5 * S.dot(S.matvec(J.T, omega), omega))
        else:
            o.put('e_kinetic', 'kite_rot' + lab, S.ZERO)
        q_mean = (q_node + q_of(arch.parent_map[node])) / 2.
        o.put('e_potential', 'tether' + lab, g_acc * seg_mass * q_mean[2])
        o.put('e_potential', 'kite' + lab, g_acc * m.par('geometry', 'm_k') * q_node[2] if is_kite else S.ZERO)

    # ---- invariants
    g_list, gdot_list, gddot_list = m.holonomic
    for i, node in enumerate(range(1, nn)):
        o.put('invariants', 'c' + arch.node_label(node), g_list[i])
    for i, node in enumerate(range(1, nn)):
        lab = arch.node_label(node)
        o.put('invariants', 'dc' + lab, gdot_list[i])
        o.put('invariants', 'ddc' + lab, gddot_list[i])
        if node in arch.kite_nodes and m.has('x', 'r' + lab):
            r = m.scaled('x', 'r' + lab).reshape(3, 3).T          # (sic) the scaled variable
            rtr = S.matmul(r.T, r)
            ortho = S.zeros(3, 3)
            for a in range(3):
                for b in range(3):
                    ortho[a, b] = rtr[a, b] - (1.0 if a == b else 0.0)
            o.put('invariants', 'orthonormality' + lab, list(ortho.T.reshape(-1)))
            dr_dt = m.si('xdot', 'dr' + lab).reshape(3, 3).T
            dr = S.matmul(r, S.skew(m.si('x', 'omega' + lab)))
            o.put('invariants', 'rot_kinematics10', list((dr_dt - dr).T.reshape(-1)))     # (sic) one fixed name

    # ---- rotation axes
    nvm = spec.opt('aero.actuator.normal_vector_model', 'tether_parallel')

    def n_hat(parent):
        if nvm == 'xhat':
            return S.vec([S.ONE, S.ZERO, S.ZERO])
        if nvm != 'tether_parallel':
            raise NotImplementedError('normal vector model %s' % nvm)
        node = parent if len(kites_of[parent]) > 1 else 1
        n_vec = q_of(node) - q_of(arch.parent_map[node])
        return _normalize(m, n_vec)

    def rotation_axes():
        rot = {}
        for parent in layer_nodes:
            rot['ehat_normal%d' % parent] = n_hat(parent)
        for kite in arch.kite_nodes:
            parent = arch.parent_map[kite]
            nrm = rot['ehat_normal%d' % parent]
            radial = _normed_cross(m, dq_of(kite), nrm)
            rot['ehat_radial%d' % kite] = radial
            rot['ehat_tangential%d' % kite] = _normed_cross(m, nrm, radial)
        return rot

    rot = rotation_axes()
    for name, v in rot.items():
        o.put('rotation', name, _vec3(v))

    # ---- per-kite aerodynamics, environment, validity, local performance, power balance
    b_ref, c_ref, s_ref = m.par('geometry', 'b_ref'), m.par('geometry', 'c_ref'), m.par('geometry', 's_ref')
    elevation = _elevation(m.si('x', 'q10'))
    framed = {}
    for kite in arch.kite_nodes:
        lab = arch.node_label(kite)
        a = m.aero[kite]
        vec_u, dcm = a['vec_u'], a['dcm']
        q, dq = q_of(kite), dq_of(kite)
        rho = m.density(q[2])
        q_eff = 0.5 * rho * S.dot(vec_u, vec_u)
        W = _wind_dcm(m, vec_u, dcm)
        Winv = inv3(W)
        f_earth = a['f_earth']
        if dof == 6:
            f_body, m_body = a['f_body'], a['m_body']
            f_control = S.vec([-f_body[0], f_body[1], -f_body[2]])
        else:
            f_body = S.matvec(inv3(dcm), f_earth)
            f_control = S.vec([-f_body[0], f_body[1], -f_body[2]])
            m_body = S.zeros(3)
        f_wind = S.matvec(Winv, S.matvec(dcm, f_body)) if dof == 6 else S.matvec(Winv, f_earth)
        framed[kite] = dict(f_wind=f_wind, f_earth=f_earth)
        u_eff = m.smooth_norm(vec_u)
        coeff_body = [f_body[i] / q_eff / s_ref for i in range(3)]
        coeff_wind = [f_wind[i] / q_eff / s_ref for i in range(3)]
        ref_len = [b_ref, c_ref, b_ref]
        cm = [(1. / ref_len[i]) * m_body[i] / q_eff / s_ref for i in range(3)]
        f_drag_wind = [f_wind[0], S.ZERO, S.ZERO]
        f_side_wind = [S.ZERO, f_wind[1], S.ZERO]
        f_lift_wind = [S.ZERO, S.ZERO, f_wind[2]]
        f_drag_earth = S.vec([W[i, 0] * f_wind[0] for i in range(3)])
        f_side_earth = S.vec([W[i, 1] * f_wind[1] for i in range(3)])
        f_lift_earth = S.vec([W[i, 2] * f_wind[2] for i in range(3)])
        CD, CS, CL = coeff_wind
        coeffs = {'CD_var': f_drag_wind, 'CS_var': f_side_wind, 'CL_var': f_lift_wind, 'CD': CD, 'CS': CS, 'CL': CL,
                  'CA': coeff_body[0], 'CY': coeff_body[1], 'CN': coeff_body[2], 'Cl': cm[0], 'Cm': cm[1], 'Cn': cm[2],
                  'LoverD': CL / CD}
        ka = str(kite)
        o.put('aerodynamics', 'air_velocity' + ka, _vec3(vec_u))
        o.put('aerodynamics', 'airspeed' + ka, u_eff)                 # overwritten below, position kept
        for cname in sorted(coeffs):
            o.put('aerodynamics', cname + ka, coeffs[cname])
        o.put('aerodynamics', 'f_aero_earth' + ka, _vec3(f_earth))
        o.put('aerodynamics', 'f_aero_body' + ka, _vec3(f_body))
        o.put('aerodynamics', 'f_aero_control' + ka, _vec3(f_control))
        o.put('aerodynamics', 'f_aero_wind' + ka, _vec3(f_wind))
        o.put('aerodynamics', 'f_lift_earth' + ka, _vec3(f_lift_earth))
        o.put('aerodynamics', 'f_drag_earth' + ka, _vec3(f_drag_earth))
        o.put('aerodynamics', 'f_side_earth' + ka, _vec3(f_side_earth))
        o.put('aerodynamics', 'f_lift_norm' + ka, f_wind[2])
        o.put('aerodynamics', 'm_aero_body' + ka, _vec3(m_body))
        o.put('aerodynamics', 'q' + ka, _vec3(q))
        o.put('aerodynamics', 'dq' + ka, _vec3(dq))
        o.put('aerodynamics', 'airspeed' + ka, m.norm(vec_u))
        o.put('aerodynamics', 'u_infty' + ka, _vec3(m.wind_velocity(q[2])))
        o.put('aerodynamics', 'u_app' + ka, _vec3(m.wind_velocity(q[2]) - dq))
        o.put('aerodynamics', 'air_density' + ka, rho)
        o.put('aerodynamics', 'dyn_pressure' + ka, 0.5 * rho * S.dot(vec_u, vec_u))
        e_chord, e_span, e_up = dcm[:, 0], dcm[:, 1], dcm[:, 2]
        o.put('aerodynamics', 'ehat_chord' + ka, _vec3(e_chord))
        o.put('aerodynamics', 'ehat_span' + ka, _vec3(e_span))
        o.put('aerodynamics', 'ehat_up' + ka, _vec3(e_up))
        dtd = S.matmul(dcm.T, dcm)
        resi = S.ZERO
        for aa in range(3):
            for bb in range(3):
                e_ = dtd[aa, bb] - (1.0 if aa == bb else 0.0)
                resi = resi + e_ * e_
        o.put('aerodynamics', 'ortho_resi' + ka, resi)
        omega = m.si('x', 'omega' + lab) if dof == 6 else S.zeros(3)
        for tip, sign in (('ext', +1.), ('int', -1.)):
            shift = S.vec([sign * e_span[i] * b_ref / 2. for i in range(3)])
            x_tip = q + shift
            u_tip = m.wind_velocity(x_tip[2]) - (dq + S.cross(omega, shift))
            o.put('aerodynamics', 'wingtip_' + tip + ka, _vec3(x_tip))
            o.put('aerodynamics', 'u_app_' + tip + ka, _vec3(u_tip))
        o.put('aerodynamics', 'fstar_aero' + ka, S.dot(vec_u, e_chord) / c_ref)
        o.put('aerodynamics', 'r' + ka, list(dcm.T.reshape(-1)))
        if dof == 6:
            o.put('aerodynamics', 'm_aero_body' + ka, _vec3(m_body))
        sound = S.power(m.par('atmosphere', 'gamma') * m.par('atmosphere', 'r') * _temperature(m, q[2]), 0.5)
        o.put('aerodynamics', 'mach' + ka, m.smooth_norm(vec_u) / sound)
        o.put('aerodynamics', 'reynolds' + ka,
              rho * S.power(S.dot(vec_u, vec_u), 0.5) * c_ref / _viscosity(m, q[2]))
        # environment
        wind = m.wind_velocity(q[2])
        o.put('environment', 'windspeed' + ka, m.norm(wind))
        o.put('environment', 'pressure' + ka, _pressure(m, q[2]))
        o.put('environment', 'temperature' + ka, _temperature(m, q[2]))
        o.put('environment', 'density' + ka, rho)
        # aero validity (same expressions as the inequalities) and angles
        u1 = m.smooth_abs(S.dot(vec_u, e_chord))
        alpha, beta = S.dot(vec_u, e_up) / u1, S.dot(vec_u, e_span) / u1
        tight = float(spec.opt('model_bounds.aero_validity.scaling', 1.0))
        v_ref = float(spec.opt('model_bounds.aero_validity.airspeed_ref'))
        rad = np.pi / 180.0
        a_min, a_max = spec.opt('aero.alpha_min_deg') * rad, spec.opt('aero.alpha_max_deg') * rad
        b_min, b_max = spec.opt('aero.beta_min_deg') * rad, spec.opt('aero.beta_max_deg') * rad
        sabs = lambda v: math.sqrt(v ** 2. + 1e-16)
        w1, w2, w3 = S.dot(vec_u, e_chord), S.dot(vec_u, e_span), S.dot(vec_u, e_up)
        o.put('aero_validity', 'alpha_ub' + ka, (w3 - w1 * a_max) * tight / v_ref / sabs(a_max))
        o.put('aero_validity', 'alpha_lb' + ka, (-w3 + w1 * a_min) * tight / v_ref / sabs(a_min))
        o.put('aero_validity', 'beta_ub' + ka, (w2 - w1 * b_max) * tight / v_ref / sabs(b_max))
        o.put('aero_validity', 'beta_lb' + ka, (-w2 + w1 * b_min) * tight / v_ref / sabs(b_min))
        o.put('aerodynamics', 'alpha' + ka, alpha)
        o.put('aerodynamics', 'beta' + ka, beta)
        o.put('aerodynamics', 'alpha_deg' + ka, alpha * 180. / np.pi)
        o.put('aerodynamics', 'beta_deg' + ka, beta * 180. / np.pi)
        # local performance (Loyd comparison)
        eps = 1.e-8
        CR = CL * S.power(1. + S.power(CD / (CL + eps), 2.), 0.5)
        windspeed = m.norm(wind)
        phf_loyd = _loyd_phf(CL, CD, elevation)
        p_loyd = _power_density(m, q[2]) * s_ref * phf_loyd
        speed_loyd = 2. * CR / 3. / CD * windspeed * S.cos(elevation)
        o.put('local_performance', 'CR' + ka, CR)
        o.put('local_performance', 'p_loyd' + ka, p_loyd)
        o.put('local_performance', 'speed_loyd' + ka, speed_loyd)
        o.put('local_performance', 'phf_loyd' + ka, phf_loyd)
        o.put('local_performance', 'speed_ratio' + ka, u_eff / windspeed)
        o.put('local_performance', 'speed_ratio_loyd' + ka, speed_loyd / windspeed)
        # power balance
        o.put('power_balance', 'P_lift' + ka, S.dot(f_lift_earth, dq))
        o.put('power_balance', 'P_drag' + ka, S.dot(f_drag_earth, dq))
        o.put('power_balance', 'P_side' + ka, S.dot(f_side_earth, dq))
        if dof == 6:
            o.put('power_balance', 'P_moment' + ka, S.dot(m_body, omega))

    # ---- geometry (centre of rotation: 'averaged' for several kites, 'parent' for a kite on a secondary tether; the
    #      'frenet' centre of a single kite on the main tether is not restated)
    n_kites = len(arch.kite_nodes)
    if n_kites > 1:
        gtype = 'averaged'
    elif arch.parent_map[arch.kite_nodes[0]] == 0:
        gtype = 'frenet'
    else:
        gtype = 'parent'
    gtype = spec.opt('aero.geometry_type', gtype)
    have_geometry = gtype in ('averaged', 'parent', 'frenet')

    def frenet(kite):
        """frenet_geometry.py:49-225: osculating circle of the kite trajectory; the third derivative of the position is
        not a model variable, so the reference takes it as zero (its own warning at :41-47)"""
        lab = arch.node_label(kite)
        g1, g2 = dq_of(kite), m.si('xdot', 'ddq' + lab)
        g3 = m.si('xdot', 'dddq' + lab) if m.has('xdot', 'dddq' + lab) else S.zeros(3)
        cr = S.cross(g1, g2)
        num_k, den_k = m.smooth_norm(cr), S.power(m.smooth_norm(g1), 3.)
        radius = den_k / num_k
        e1 = _normalize(m, g1, 1e-8)
        inner = S.cross(g2, g1)
        e2n = S.cross(g1, inner)
        e2d = m.smooth_norm(g1) * m.smooth_norm(inner)
        e2 = S.vec([e2n[i] / e2d for i in range(3)])
        e3 = _normed_cross(m, e1, e2, 1e-8)
        # d radius / dt
        d_cr = S.cross(g1, g3)
        cr2 = S.dot(cr, cr)
        d_num_k = 0.5 * (2. * S.dot(cr, d_cr)) / m.smooth_sqrt(cr2)
        d_den_k = (3. / 2.) * m.smooth_sqrt(S.dot(g1, g1)) * (2. * S.dot(g1, g2))
        d_radius = (num_k * d_den_k - den_k * d_num_k) / S.power(num_k, 2.)
        torsion = S.dot(g3, cr) / (cr2 + 1.e-8)
        curvature = 1. / radius
        d_e2_ds = [-1. * curvature * e1[i] + torsion * e3[i] for i in range(3)]
        speed = m.smooth_norm(g1)
        d_e2_dt = [speed * d_e2_ds[i] for i in range(3)]
        x_c = S.vec([q_of(kite)[i] + radius * e2[i] for i in range(3)])
        dx_c = S.vec([g1[i] + (d_radius * e2[i] + radius * d_e2_dt[i]) for i in range(3)])
        return x_c, dx_c

    def center(parent):
        ks = kites_of[parent]
        if gtype == 'averaged':
            x = S.zeros(3)
            dx = S.zeros(3)
            for k in ks:
                x = x + q_of(k) / float(len(ks))
                dx = dx + dq_of(k) / float(len(ks))
            return x, dx
        if gtype == 'frenet':
            children = [n for n in range(1, nn) if arch.parent_map[n] == parent]
            x = S.zeros(3)
            dx = S.zeros(3)
            for k in children:
                xc, dxc = frenet(k)
                x = x + xc / float(len(children))
                dx = dx + dxc / float(len(children))
            return x, dx
        return q_of(parent), dq_of(parent)

    if have_geometry:
        for kite in arch.kite_nodes:
            parent = arch.parent_map[kite]
            ka = str(kite)
            x_c, dx_c = center(parent)
            vec_v = dq_of(kite) - dx_c
            to_kite = q_of(kite) - x_c
            radial = rot['ehat_radial%d' % kite]
            radius = m.smooth_abs(S.dot(radial, to_kite), 0.)
            proj = S.dot(radial, vec_v)
            v_perp = S.vec([vec_v[i] - proj * radial[i] for i in range(3)])
            om = m.norm(v_perp) / radius
            o.put('geometry', 'local_period_of_rotation' + ka, 2. * np.pi / om)
            o.put('geometry', 'vector_from_center_to_kite' + ka, _vec3(to_kite))
            o.put('geometry', 'radius' + ka, m.smooth_abs(S.dot(to_kite, radial), 0.))
            theta_vec = S.cross(S.vec([S.ONE, S.ZERO, S.ZERO]), to_kite)
            comp = S.dot(vec_v, theta_vec)
            o.put('geometry', 'kite_motion_is_right_hand_rule_positive_around_wind_direction' + ka,
                  S.atan(comp / 1e-8) / np.pi + 0.5)
            g1, g2 = dq_of(kite), m.si('xdot', 'ddq' + arch.node_label(kite))
            roc = S.power(m.smooth_norm(g1), 3.) / m.smooth_norm(S.cross(g1, g2))
            o.put('geometry', 'radius_of_curvature' + ka, roc)
        for parent in layer_nodes:
            pa = str(parent)
            x_c, dx_c = center(parent)
            w_c = m.wind_velocity(x_c[2])
            u_zero = w_c - dx_c
            o.put('geometry', 'x_center' + pa, _vec3(x_c))
            o.put('geometry', 'dx_center' + pa, _vec3(dx_c))
            o.put('geometry', 'wind_velocity_at_center' + pa, _vec3(w_c))
            o.put('geometry', 'vec_u_zero' + pa, _vec3(u_zero))
            o.put('geometry', 'u_zero' + pa, m.norm(u_zero))
            roc_avg, rad_avg, per_avg = S.ZERO, S.ZERO, S.ZERO
            ks = kites_of[parent]
            for k in ks:
                nsib = float(len(ks))
                roc_avg = roc_avg + o.get('geometry', 'radius_of_curvature%d' % k)[0] / nsib
                rad_avg = rad_avg + o.get('geometry', 'radius%d' % k)[0] / nsib
                per_avg = per_avg + o.get('geometry', 'local_period_of_rotation%d' % k)[0] / nsib
            o.put('geometry', 'average_radius_of_curvature' + pa, roc_avg)
            o.put('geometry', 'average_relative_radius_of_curvature' + pa, roc_avg / b_ref)
            o.put('geometry', 'average_curvature' + pa, 1. / roc_avg)
            o.put('geometry', 'average_period_of_rotation' + pa, per_avg)
            o.put('geometry', 'average_radius' + pa, rad_avg)
            o.put('geometry', 'average_relative_radius' + pa, rad_avg / b_ref)

    # ---- circulation (indicators.py:62-118)
    for kite in arch.kite_nodes:
        ka = str(kite)
        a = m.aero[kite]
        vec_u, dcm = a['vec_u'], a['dcm']
        rho = o.get('aerodynamics', 'air_density' + ka)[0]
        airspeed = o.get('aerodynamics', 'airspeed' + ka)[0]
        CL = o.get('aerodynamics', 'CL' + ka)[0]
        f_lift_norm = framed[kite]['f_wind'][2]
        W = _wind_dcm(m, vec_u, dcm)
        lift_hat = W[:, 2]
        cr = S.cross(vec_u, dcm[:, 1])
        circ_dot = f_lift_norm / (rho * b_ref) / S.dot(lift_hat, cr)
        o.put('aerodynamics', 'circulation_dot' + ka, circ_dot)
        o.put('aerodynamics', 'circulation_cross' + ka, f_lift_norm / b_ref / rho / m.smooth_norm(cr))
        o.put('aerodynamics', 'circulation_cl' + ka, 0.5 * S.power(airspeed, 2.) * CL * c_ref / m.smooth_norm(cr))
        o.put('aerodynamics', 'circulation' + ka, circ_dot)
    for parent in layer_nodes:
        pa = str(parent)
        total = S.ZERO
        for k in kites_of[parent]:
            total = total + o.get('aerodynamics', 'circulation%d' % k)[0]
        o.put('aerodynamics', 'total_circulation' + pa, total)
        if have_geometry:
            u_zero = S.vec(o.get('geometry', 'vec_u_zero' + pa))
            period = o.get('geometry', 'average_period_of_rotation' + pa)[0]
            nb = period * u_zero[0]
            qc = period * total
            pitch = (nb + period * S.power(m.smooth_abs(S.power(nb, 2.) - 4. * qc), 0.5)) / 2.
            o.put('aerodynamics', 'far_wake_cylinder_pitch' + pa, pitch)
        thrust = S.ZERO
        nrm = rot['ehat_normal%d' % parent]
        for k in kites_of[parent]:
            thrust = thrust + S.dot(m.aero[k]['f_earth'], nrm)
        o.put('aerodynamics', 'thrust' + pa, thrust)

    # ---- performance (indicators.py:117-152)
    dl_t = m.si(m.var_type_of('dl_t'), 'dl_t')[0] if m.has('x', 'dl_t') else S.ZERO

    def reelout_factor(pos):
        return dl_t / m.smooth_norm(m.wind_velocity(pos[2]))

    o.put('performance', 'f', reelout_factor(m.si('x', 'q10')))
    o.put('performance', 'elevation', elevation)
    if have_geometry:
        for parent in layer_nodes:
            x_c, _ = center(parent)
            o.put('performance', 'trajectory_center%d' % parent, _vec3(x_c))
            o.put('performance', 'f%d' % parent, reelout_factor(x_c))
    p_loyd_total = S.ZERO
    for kite in arch.kite_nodes:
        p_loyd_total = p_loyd_total + o.get('local_performance', 'p_loyd%d' % kite)[0]
    o.put('performance', 'p_loyd_total', p_loyd_total)
    avail = S.ZERO
    for kite in arch.kite_nodes:
        avail = avail + _power_density(m, q_of(kite)[2]) * s_ref
    lam10 = m.si('z', 'lambda10')[0]
    l_t = m.si(m.var_type_of('l_t'), 'l_t')[0]
    if m.has('x', 'l_t'):
        current_power = lam10 * l_t * m.si('x', 'dl_t')[0]
    else:
        current_power = lam10 * l_t
    avail_hub = _power_density(m, m.si('x', 'q10')[2]) * s_ref * float(n_kites)
    o.put('performance', 'phf', current_power / avail)
    o.put('performance', 'phf_hubheight', current_power / avail_hub)
    o.put('performance', 'hubheight_power_availability', avail_hub)
    o.put('performance', 'phf_loyd_total', p_loyd_total / avail_hub)
    o.put('performance', 'loyd_factor', current_power / m.smooth_abs(p_loyd_total, 1.e-4))
    o.put('performance', 'power_density', current_power / float(n_kites) / s_ref)

    # ---- tether aerodynamics (tether_aero.py:121-174)
    n_el = int(spec.opt('tether.aero_elements'))
    drag_model = spec.opt('tether.tether_drag.model_type', 'multi')
    drag_on = {}
    for node in range(1, nn):
        ns = str(node)
        multi_low, multi_up = m.segment_drag(node, n_el)
        split_low, split_up = m.segment_drag(node, 1)
        parent = arch.parent_map[node]
        q_up, q_lo, dq_up, dq_lo = q_of(node), q_of(parent), dq_of(node), dq_of(parent)
        diam = m.si('theta', 'diam_t')[0] if (parent == 0 or node not in arch.kite_nodes) else m.si('theta', 'diam_s')[0]
        ko_low, ko_up = S.zeros(3), S.zeros(3)
        if node in arch.kite_nodes:
            e1 = o.get('aerodynamics', 'ehat_chord' + ns)
            e3 = o.get('aerodynamics', 'ehat_span' + ns)                  # (sic) segment.py:73
            alpha = o.get('aerodynamics', 'alpha' + ns)[0]
            d_hat = [S.cos(alpha) * e1[i] + S.sin(alpha) * e3[i] for i in range(3)]
            dyn_p = o.get('aerodynamics', 'dyn_pressure' + ns)[0]
            length = m.norm(q_up - q_lo)
            d_mag = (1. / 4.) * m.par('tether', 'cd') * dyn_p * diam * length
            ko_up = S.vec([d_mag * d_hat[i] for i in range(3)])
        o.put('tether_aero', 'multi_upper' + ns, _vec3(multi_up))
        o.put('tether_aero', 'multi_lower' + ns, _vec3(multi_low))
        o.put('tether_aero', 'split_upper' + ns, _vec3(split_up))
        o.put('tether_aero', 'split_lower' + ns, _vec3(split_low))
        o.put('tether_aero', 'kite_only_upper' + ns, _vec3(ko_up))
        o.put('tether_aero', 'kite_only_lower' + ns, _vec3(ko_low))
        if drag_model == 'multi':
            d_node, d_parent = multi_up, multi_low
        elif drag_model == 'split':
            d_node, d_parent = split_up, split_low
        elif drag_model == 'kite_only':
            d_node, d_parent = ko_up, ko_low
        else:
            d_node, d_parent = S.zeros(3), S.zeros(3)
        o.put('tether_aero', 'homotopy_upper' + ns, _vec3(d_node))
        o.put('tether_aero', 'homotopy_lower' + ns, _vec3(d_parent))
        # Reynolds number of the segment (reynolds.py, element.py:get_uapp)
        q_avg = (q_up + q_lo) / 2.
        ua = m.wind_velocity(q_avg[2]) - (dq_up + dq_lo) / 2.
        o.put('tether_aero', 'reynolds' + ns,
              m.density(q_avg[2]) * m.norm(ua) * diam / _viscosity(m, q_avg[2]))
        drag_on[node] = (d_parent, d_node)
    # distribution to the nodes and the drag losses (tether_aero.py:92-118, indicators.py:265-276)
    node_drag = {n: S.zeros(3) for n in range(1, nn)}
    for node in range(1, nn):
        d_parent, d_node = drag_on[node]
        node_drag[node] = node_drag[node] + d_node
        if node > 1:
            par = arch.parent_map[node]
            node_drag[par] = node_drag[par] + d_parent
    for node in range(1, nn):
        o.put('power_balance', 'P_tetherdrag%d' % node, S.dot(node_drag[node], dq_of(node)))

    # ---- tether moments
    if dof == 6:
        for kite in arch.kite_nodes:
            o.put('tether_moments', 'n' + arch.node_label(kite), _vec3(m.tether_moment(m.work_holonomic, kite)))

    # ---- tether stress / force (dynamics.py:706-790), yaw angle (dynamics.py:1062-1117)
    for node in range(1, nn):
        lab = arch.node_label(node)
        seg = m.segment(node)
        tension = m.si('z', 'lambda' + lab)[0] * seg['seg_length']
        o.put('local_performance', 'tether_stress' + lab, tension / seg['cross_section_area'])
        o.put('local_performance', 'tether_force' + lab, tension)
    if dof == 6 and spec.opt('model_bounds.rotation.include') and spec.opt('model_bounds.rotation.type') == 'yaw':
        for kite in arch.kite_nodes:
            lab = arch.node_label(kite)
            q_hat = q_of(kite) - q_of(arch.parent_map[kite])
            R = m.si('x', 'r' + lab).reshape(3, 3).T
            o.put('local_performance', 'rot_angles' + lab, S.acos(S.dot(q_hat, R[:, 2]) / m.norm(q_hat)))

    o.put('performance', 'p_current', current_power)

    # ---- constraints, xdot
    o.touch('model_inequalities')
    for name, rows in m.ineq:
        o.put('model_inequalities', name, rows)
    o.touch('model_equalities')
    for name, rows in m.eq:
        o.put('model_equalities', name, rows)
    for name in m.names('xdot'):
        o.put('xdot_from_var', name, list(m.si('xdot', name)))

    # ---- power balance, second part (dynamics.py:341-425)
    for node in range(1, nn):
        lab = arch.node_label(node)
        tdir = _normalize(m, q_of(node) - q_of(arch.parent_map[node]))
        force = o.get('local_performance', 'tether_force' + lab)[0]
        o.put('power_balance', 'P_tether%d' % node, -S.dot(S.vec([force * tdir[i] for i in range(3)]), dq_of(node)))
    for source in list(o.groups['e_kinetic']):
        o.put('power_balance', 'P_kin_' + source, -1. * m.time_derivative(o.get('e_kinetic', source)[0]))
    for source in list(o.groups['e_potential']):
        o.put('power_balance', 'P_pot_' + source, -1. * m.time_derivative(o.get('e_potential', source)[0]))
    return o.flat()
