"""Synthetic primal points of the named OCP shapes.

Base point ``w0``: a circular-flight guess in scaled units — the kite(s) fly a circle on a cone around
the tether axis (inclination / cone angle / ground speed of ampyx_ap2_settings.py:65-68), the DCM has
its chord axis along the apparent wind and its up axis along the tether, omega is the circular rate,
collocation states are that circle sampled at the Radau times, xdot its time derivative.  Batch
points are ``w_b = w0 * (1 + 1e-2 * U(-1, 1))`` (no re-orthonormalisation: the residual must cope
with non-orthonormal R).
"""
import math

import numpy as np


def _unit(v):
    return v / np.linalg.norm(v)


def base_point(ocp, l_t=200.0, inclination_deg=45.0, cone_deg=15.0, groundspeed=15.0, l_s=50.0):
    m = ocp.model
    arch = m.arch
    spec = ocp.spec
    incl, cone = math.radians(inclination_deg), math.radians(cone_deg)
    e_c = np.array([math.cos(incl), 0.0, math.sin(incl)])
    e_y = np.array([0.0, 1.0, 0.0])
    e_zp = np.array([-math.sin(incl), 0.0, math.cos(incl)])
    single = arch.number_of_nodes == 2
    seg = l_t if single else l_s
    radius = seg * math.sin(cone)
    t_f = 2.0 * math.pi * radius / groundspeed
    omega_c = 2.0 * math.pi / t_f
    u_ref = 10.0
    for lab, val in zip(spec.theta0_labels, spec.theta0_values):
        if lab == '[theta0,wind,u_ref,0]':
            u_ref = float(val)

    def state_at(t):
        """dict name -> SI value of every state, and its time derivative."""
        x, dx = {}, {}
        for node in range(1, arch.number_of_nodes):
            lab = arch.node_label(node)
            if node in arch.kite_nodes:
                sib = arch.children_map[arch.parent_map[node]]
                phase = omega_c * t + 2.0 * math.pi * sib.index(node) / max(1, len(sib))
                root = np.zeros(3) if single else l_t * e_c
                rad = math.cos(phase) * e_y + math.sin(phase) * e_zp
                drad = omega_c * (-math.sin(phase) * e_y + math.cos(phase) * e_zp)
                ddrad = -omega_c ** 2 * rad
                q = root + seg * (math.cos(cone) * e_c + math.sin(cone) * rad)
                dq = seg * math.sin(cone) * drad
                ddq = seg * math.sin(cone) * ddrad
            else:
                q, dq, ddq = l_t * e_c, np.zeros(3), np.zeros(3)
            x['q' + lab], dx['q' + lab] = q, dq
            x['dq' + lab], dx['dq' + lab] = dq, ddq
            if node in arch.kite_nodes and int(spec.opt('kite_dof')) == 6:
                ua = np.array([u_ref, 0.0, 0.0]) - dq
                e1 = _unit(ua)
                e3 = _unit((q - (np.zeros(3) if single else l_t * e_c)))
                e2 = _unit(np.cross(e3, e1))
                e3 = np.cross(e1, e2)
                R = np.stack([e1, e2, e3], axis=1)
                w_earth = omega_c * e_c
                w_body = R.T @ w_earth
                x['omega' + lab], dx['omega' + lab] = w_body, np.zeros(3)
                x['r' + lab] = R.T.reshape(-1)                       # column-major 9-vector
                skew = np.array([[0, -w_body[2], w_body[1]], [w_body[2], 0, -w_body[0]], [-w_body[1], w_body[0], 0]])
                dx['r' + lab] = (R @ skew).T.reshape(-1)
                x['delta' + lab], dx['delta' + lab] = np.zeros(3), np.zeros(3)
            elif node in arch.kite_nodes:
                x['coeff' + lab], dx['coeff' + lab] = np.array([1.0, 0.0]), np.zeros(2)
        x['l_t'], dx['l_t'] = np.array([l_t]), np.zeros(1)
        x['dl_t'], dx['dl_t'] = np.zeros(1), np.zeros(1)
        if m.has('x', 'ddl_t'):
            x['ddl_t'], dx['ddl_t'] = np.zeros(1), np.zeros(1)
        if m.has('x', 'e'):
            x['e'], dx['e'] = np.zeros(1), np.zeros(1)
        return x, dx

    def pack(vt, values, scaled_like=None):
        out = np.zeros(sum(sz for _, sz in m.var_list[vt]))
        k = 0
        for nm, sz in m.var_list[vt]:
            key = nm if vt != 'xdot' else nm[1:]
            src = values.get(key)
            if src is not None:
                out[k:k + sz] = np.asarray(src, dtype=float) / m.scale(vt, nm)
            k += sz
        return out

    w = np.zeros(ocp.n_v)
    for nm in ocp.theta_names:
        o, sz = ocp.v_theta[nm]
        si = {'t_f': t_f, 'diam_t': 2e-3, 'l_s': l_s, 'diam_s': 2e-3, 'l_i': 100.0, 'diam_i': 4e-3}.get(nm, 1.0)
        if nm == 't_f' and sz == 2:
            w[o:o + 2] = t_f / m.scale('theta', nm)[0]
        else:
            w[o] = si / m.scale('theta', nm)[0]
    w[ocp.v_phi:ocp.v_phi + ocp.n_phi] = 1.0
    h = t_f / ocp.n_k
    tau = ocp.coll.tau_root
    u0 = np.zeros(ocp.n_u)
    z0 = np.ones(ocp.n_z)
    for k in range(ocp.n_k + 1):
        x, dx = state_at(k * h)
        w[ocp.v_x[k]:ocp.v_x[k] + ocp.n_x] = pack('x', x)
        if k < ocp.n_k:
            w[ocp.v_u[k]:ocp.v_u[k] + ocp.n_u] = u0
            w[ocp.v_xdot[k]:ocp.v_xdot[k] + ocp.n_x] = pack('xdot', dx)
            w[ocp.v_z[k]:ocp.v_z[k] + ocp.n_z] = z0
            for j in range(ocp.d):
                xj, _ = state_at((k + tau[j + 1]) * h)
                w[ocp.v_cx[k][j]:ocp.v_cx[k][j] + ocp.n_x] = pack('x', xj)
                w[ocp.v_cz[k][j]:ocp.v_cz[k][j] + ocp.n_z] = z0
    return w


def batch(ocp, B, seed, w0=None, rel=1e-2):
    w0 = base_point(ocp) if w0 is None else w0
    rng = np.random.default_rng(seed)
    return w0[None, :] * (1.0 + rel * rng.uniform(-1.0, 1.0, (B, ocp.n_v)))


def scenario_params(ocp, W, seed, rel=1e-2):
    """per-instance parameter vectors of an MPC scenario batch (BASELINE.json configs[4]): the tracking
    reference ``p.ref`` (a V-shaped block at the start of P, awebox/ocp/discretization.py:157-179) is the
    instance's own primal point perturbed by +-rel, so that the initial-condition rows x_0 - p.ref[x_0] differ
    per instance, and the reference wind speed is perturbed by +-rel."""
    rng = np.random.default_rng(seed)
    B = W.shape[0]
    P = np.tile(ocp.default_p(), (B, 1))
    P[:, :ocp.n_v] = W * (1.0 + rel * rng.uniform(-1.0, 1.0, W.shape))
    if '[theta0,wind,u_ref,0]' in ocp.spec.theta0_labels:
        iu = ocp.p_theta0 + ocp.spec.theta0_labels.index('[theta0,wind,u_ref,0]')
        P[:, iu] *= 1.0 + rel * rng.uniform(-1.0, 1.0, B)
    return P


def power_curve_params(ocp, B, n_speeds=64, lo=5.0, hi=8.0, first=0):
    """per-instance theta0 vectors of a power-curve batch (BASELINE.json configs[3]; the sweep of
    examples/dual_kites_power_curve.py:44-52 over ``params.wind.u_ref``): instance ``first + i`` of the global batch
    gets wind speed number ``(first + i) mod n_speeds`` of ``linspace(lo, hi, n_speeds)``.  -> [B, n_theta0]"""
    P = np.tile(np.asarray(ocp.spec.theta0_values, dtype=float), (B, 1))
    iu = ocp.spec.theta0_labels.index('[theta0,wind,u_ref,0]')
    speeds = np.linspace(lo, hi, n_speeds)
    P[:, iu] = speeds[(first + np.arange(B)) % n_speeds]
    return P
