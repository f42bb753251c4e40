"""Generate the model specs and the golden vectors by running the REFERENCE here.  TEST INFRASTRUCTURE.

    python oracle/make_golden.py [config ...]

For every configuration (oracle/ref_build.py) this script imports the reference's own Python from
/root/reference over the CasADi restatement (oracle/casadi_shim), builds the reference's model and
NLP, takes the reference's own initial guess as base point and writes

  awebox_b200/specs/<config>.json     numeric ModelSpec read off the reference's Options / Model
  tests/golden/<config>_nk<N>.npz     w [n, n_v], p [n_p], g [n, n_g] = NLP.g_fun(w, p),
                                      colind/row = structural CCS pattern of d g / d V,
                                      jv [n, nnz] = its values (compressed forward sweeps),
                                      node_vars/node_eq/node_ineq = Model.dynamics / inequality
                                      function values at random node inputs, and the V / g labels' hash

These files are what the oracle (oracle/awebox_oracle.c) and the CUDA path are pinned against.  The
script cannot run on the GPU box (no /root/reference there); its outputs are committed.

Live-CasADi export (for a machine that has ``casadi==3.6.4`` and the reference; NOT runnable in this
container, where the package cannot be installed):

    AWEBOX_REAL_CASADI=1 AWEBOX_REFERENCE=/path/to/awebox python oracle/make_golden.py

then imports the real package, takes the CCS pattern from ``Function.sparsity_out`` of the Jacobian function
CasADi derives (``jacobian(g, V)``, what ``nlpsol`` uses: awebox/opti/preparation.py:373-379) and its non-zeros
from a numeric call, and writes ``tests/golden_casadi/<config>_nk<N>.npz`` (same keys).  tests/test_parity_casadi.py
consumes those files and is skipped, with that reason, while they do not exist.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_build                                   # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
JOBS = {
    # config: [(n_k, number of points with Jacobian)]
    'ampyx_ap2': [(4, 3), (40, 8)],
    'dual_kites': [(3, 2), (20, 8)],
    'mpc_3dof': [(3, 3), (10, 3)],
}


def perturbed_points(w0, n, seed):
    rng = np.random.default_rng(seed)
    pts = [w0]
    for _ in range(n - 1):
        pts.append(w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, w0.shape)))
    return np.array(pts)


def run(config, n_k, n_pts, write_spec):
    from awebox_b200 import spec as bspec
    t0 = time.time()
    T = ref_build.build(config, n_k=n_k)
    import casadi as cas            # the restatement under oracle/casadi_shim (activated by ref_build)
    theta0 = np.array(T.p_fix_num['theta0'].full()).ravel()
    sp = bspec.from_awebox(config, T.options, T.architecture, T.model, theta0_values=theta0)
    n_v = T.nlp.V.cat.shape[0]
    real = 'shim' not in cas.__version__
    if real:
        # live CasADi: the Jacobian function nlpsol derives from g_fun
        Vs, Ps = cas.MX.sym('V', n_v), cas.MX.sym('P', T.nlp.P.cat.shape[0])
        jac_fun = cas.Function('jac_g', [Vs, Ps], [cas.jacobian(T.nlp.g_fun(Vs, Ps), Vs)])
        try:
            jac_fun = jac_fun.expand()      # the reference solves with expand=True (awebox/opts/funcs.py:268-286)
        except Exception:
            pass
        sp_j = jac_fun.sparsity_out(0)
        colind, row = np.array(sp_j.colind(), dtype=np.int64), np.array(sp_j.row(), dtype=np.int64)
        tape = None
    else:
        T.nlp.g_fun(cas.DM(T.w0), cas.DM(T.p0))            # the first numeric call records the instruction tape
        tape = T.nlp.g_fun._tape
        colind, row = tape.jac_ccs(n_v)
    g_labels = T.nlp.g.labels()
    n_init = sum(1 for l in g_labels if l.startswith('[initial,'))
    if n_init:
        # the reference builds the initial-condition rows by iterating over a set of state names
        # (awebox/ocp/operation.py:312-315): record the order this run produced (run with PYTHONHASHSEED=0)
        v_labels = T.nlp.V.labels()
        order = []
        for r_ in range(n_init):
            cols = [c for c in range(n_v) if r_ in row[colind[c]:colind[c + 1]]]
            names = [v_labels[c][1:-1].split(',') for c in cols if v_labels[c].startswith('[x,0,')]
            assert len(names) == 1, (r_, names)
            if not order or order[-1] != names[0][2]:
                order.append(names[0][2])
        sp.d['nlp']['initial_order'] = order
        sp.nlp = sp.d['nlp']
    if write_spec and not real:
        os.makedirs(bspec.SPEC_DIR, exist_ok=True)
        sp.to_json(os.path.join(bspec.SPEC_DIR, config + '.json'))
    W = perturbed_points(T.w0, n_pts, 20261018 + n_k)
    p = T.p0
    G = []
    for w in W:
        G.append(np.array(T.nlp.g_fun(cas.DM(w), cas.DM(p)).full()).ravel())
    if real:
        JV = [np.array(jac_fun(cas.DM(w), cas.DM(p)).nonzeros(), dtype=float) for w in W]
    else:
        JV = [tape.jac_values(np.concatenate([w, p]), colind, row) for w in W]
    # node-level vectors
    M = T.model
    mdl_eq = M.constraints_list.get_function(T.options['nlp'], M.variables, M.parameters, 'eq')
    mdl_ineq = M.constraints_list.get_function(T.options['nlp'], M.variables, M.parameters, 'ineq')
    rng = np.random.default_rng(7 + n_k)
    n_var = M.variables.cat.shape[0]
    node_vars = rng.uniform(0.5, 1.5, (8, n_var))
    node_p = np.concatenate([theta0, np.ones(M.parameters.cat.shape[0] - len(theta0))])
    node_eq = np.array([np.array(mdl_eq(cas.DM(v), cas.DM(node_p)).full()).ravel() for v in node_vars])
    node_ineq = np.array([np.array(mdl_ineq(cas.DM(v), cas.DM(node_p)).full()).ravel() for v in node_vars])
    labels = '\n'.join(T.nlp.V.labels()) + '\n--\n' + '\n'.join(T.nlp.g.labels())
    gdir = GOLDEN + '_casadi' if real else GOLDEN
    os.makedirs(gdir, exist_ok=True)
    out = os.path.join(gdir, '%s_nk%d.npz' % (config, n_k))
    np.savez_compressed(out, w=W, p=p, g=np.array(G), colind=colind, row=row, jv=np.array(JV),
                        node_vars=node_vars, node_p=node_p, node_eq=node_eq, node_ineq=node_ineq,
                        eq_names=np.array(M.constraints_list.get_name_list('eq')),
                        ineq_names=np.array(M.constraints_list.get_name_list('ineq')),
                        var_labels=np.array(M.variables.labels()),
                        labels_sha256=np.array(hashlib.sha256(labels.encode()).hexdigest()),
                        casadi_version=np.array(str(cas.__version__)))
    print('%s n_k=%d: n_v=%d n_g=%d nnz=%d  %d points  (%.1f s)  -> %s'
          % (config, n_k, n_v, len(G[0]), len(row), n_pts, time.time() - t0, os.path.relpath(out, ROOT)))


def run_objective(config, n_k, n_pts):
    """golden vectors of the objective path: nlp.f_fun and its gradient with respect to V
    (awebox/ocp/objective.py:422-520), the cost components, and the node-level integrands (power integral output,
    side-slip angle) -> tests/golden/<config>_nk<N>_obj.npz.  P is filled the way the reference does before a solve
    (awebox/opti/preparation.py:111-147: weights and references) with a perturbed copy of the initial guess as
    reference and distinct non-zero cost factors, so that every term of the objective is exercised."""
    t0 = time.time()
    T = ref_build.build(config, n_k=n_k)
    import casadi as cas
    import awebox.opti.preparation as prep
    nlp = T.nlp
    rng = np.random.default_rng(20261019 + n_k)
    n_v = len(T.w0)
    V_ref = nlp.V(T.w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, n_v)) + 1e-3 * rng.uniform(-1.0, 1.0, n_v))
    pfix = prep.add_weights_and_refs_to_opti_parameters(T.p_fix_num, V_ref, nlp, T.model, T.V_init, T.options['solver'])
    p = np.array(pfix.cat.full()).ravel()
    labs = nlp.P.labels()
    cost_idx = [i for i, l in enumerate(labs) if l.startswith('[cost,')]
    cost_names = [labs[i][1:-1].split(',')[1] for i in cost_idx]
    for n, i in enumerate(cost_idx):
        p[i] = 0.5 + 0.1 * n
    W = np.array([T.w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, n_v)) + 1e-3 * rng.uniform(-1.0, 1.0, n_v)
                  for _ in range(n_pts)])
    # homotopy parameters inside (0, 1): psi weighs tracking against power
    phi_idx = [i for i, l in enumerate(nlp.V.labels()) if l.startswith('[phi,')]
    W[:, phi_idx] = rng.uniform(0.2, 0.8, (n_pts, len(phi_idx)))
    F, GR, COMP = [], [], []
    comp_funs, _ = nlp.cost_components
    comp_names = [k[:-4] for k in comp_funs.keys()]
    nlp.f_fun(cas.DM(W[0]), cas.DM(p))
    tape = nlp.f_fun._tape
    colind, row = tape.jac_ccs(n_v)
    for w in W:
        F.append(float(nlp.f_fun(cas.DM(w), cas.DM(p))))
        vals = tape.jac_values(np.concatenate([w, p]), colind, row)
        g = np.zeros(n_v)
        for c in range(n_v):
            if colind[c + 1] > colind[c]:
                g[c] = vals[colind[c]]
        GR.append(g)
        COMP.append([float(fn(cas.DM(w), cas.DM(p))) for fn in comp_funs.values()])
    # node-level integrands at random node inputs
    M = T.model
    theta0 = np.array(T.p_fix_num['theta0'].full()).ravel()
    n_var = M.variables.cat.shape[0]
    node_vars = rng.uniform(0.5, 1.5, (8, n_var))
    node_p = np.concatenate([theta0, np.ones(M.parameters.cat.shape[0] - len(theta0))])
    node_int = np.array([np.array(M.integral_outputs_fun(cas.DM(v), cas.DM(node_p)).full()).ravel() for v in node_vars])
    out_labels = M.outputs.labels()
    beta_idx = [i for i, l in enumerate(out_labels) if l.startswith('[aerodynamics,beta')]
    node_out = np.array([np.array(M.outputs_fun(cas.DM(v), cas.DM(node_p)).full()).ravel() for v in node_vars])
    out = os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (config, n_k))
    np.savez_compressed(out, w=W, p=p, f=np.array(F), grad=np.array(GR), components=np.array(COMP),
                        component_names=np.array(comp_names), cost_names=np.array(cost_names),
                        node_vars=node_vars, node_p=node_p, node_integral=node_int,
                        integral_names=np.array(list(M.integral_outputs.keys())),
                        node_beta=node_out[:, beta_idx], beta_labels=np.array([out_labels[i] for i in beta_idx]),
                        quad_weights=np.array(nlp.Collocation.quad_weights, dtype=float).ravel()
                        if hasattr(nlp, 'Collocation') else np.zeros(0))
    print('%s n_k=%d objective: f %s  (%.1f s) -> %s' % (config, n_k, F[:2], time.time() - t0, os.path.relpath(out, ROOT)))


def run_mpc_objective(config, n_k, n_pts):
    """golden vectors of the tracking cost of the periodic MPC (BASELINE configs[4]): the reference's own
    Pmpc.__generate_objective (awebox/pmpc.py:304-358) run on the MPC trial, with the parameter struct of
    Pmpc.__construct_solver (pmpc.py:152-162).  Pmpc's constructor needs a solved periodic trial and IPOPT; here the
    object is created without it and given exactly the attributes the method reads."""
    t0 = time.time()
    T = ref_build.build(config, n_k=n_k)
    import casadi as ct
    import casadi.tools as ctt
    import awebox.pmpc as pmpc
    nlp = T.nlp
    nx, nu, nz = (T.model.variables_dict[k].cat.shape[0] for k in ('x', 'u', 'z'))
    p_struct = ctt.struct_symMX([
        ctt.entry('x0', shape=(nx,)),
        ctt.entry('ref', struct=nlp.V),
        ctt.entry('u_ref', shape=(1, 1)),
        ctt.entry('Q', shape=(nx, 1)),
        ctt.entry('R', shape=(nu, 1)),
        ctt.entry('P', shape=(nx, 1)),
    ])

    class _Trial(object):
        pass
    trial = _Trial()
    trial.nlp = nlp
    obj = pmpc.Pmpc.__new__(pmpc.Pmpc)
    for name, val in (('trial', trial), ('mpc_options', {'u_param': 'zoh'}), ('N', n_k), ('d', nlp.d), ('nx', nx),
                      ('nu', nu), ('nz', nz), ('p', p_struct), ('cost_type', 'tracking')):
        setattr(obj, '_Pmpc__' + name, val)
    f = obj._Pmpc__generate_objective()
    f_fun = ct.Function('f_mpc', [nlp.V, p_struct], [f])
    rng = np.random.default_rng(20261020 + n_k)
    n_v, n_p = len(T.w0), p_struct.cat.shape[0]
    W = np.array([T.w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, n_v)) + 1e-3 * rng.uniform(-1.0, 1.0, n_v)
                  for _ in range(n_pts)])
    P = []
    for _ in range(n_pts):
        pn = p_struct(0.0)
        pn['x0'] = rng.uniform(-1.0, 1.0, nx)
        pn['ref'] = T.w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, n_v)) + 1e-3 * rng.uniform(-1.0, 1.0, n_v)
        pn['u_ref'] = rng.uniform(0.5, 1.5)
        pn['Q'], pn['R'], pn['P'] = rng.uniform(0.5, 2.0, nx), rng.uniform(0.5, 2.0, nu), rng.uniform(0.5, 2.0, nx)
        P.append(np.array(pn.cat.full()).ravel())
    P = np.array(P)
    f_fun(ct.DM(W[0]), ct.DM(P[0]))
    tape = f_fun._tape
    colind, row = tape.jac_ccs(n_v)
    F, GR = [], []
    for w, pv in zip(W, P):
        F.append(float(f_fun(ct.DM(w), ct.DM(pv))))
        vals = tape.jac_values(np.concatenate([w, pv]), colind, row)
        g = np.zeros(n_v)
        for c in range(n_v):
            if colind[c + 1] > colind[c]:
                g[c] = vals[colind[c]]
        GR.append(g)
    out = os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (config, n_k))
    np.savez_compressed(out, w=W, p=P, f=np.array(F), grad=np.array(GR), p_labels=np.array(p_struct.labels()))
    print('%s n_k=%d MPC tracking cost: f %s  (%.1f s) -> %s' % (config, n_k, F[:2], time.time() - t0,
                                                                 os.path.relpath(out, ROOT)))


def run_hessian(config, n_k, n_pts):
    """golden vectors of the second-order hook: the upper triangle of the Hessian of
    lam_g' g(V, P) with respect to V, pattern and values, formed the way casadi's nlpsol forms ``nlp_hess_l`` from the
    reference's own ``nlp.g_fun`` (awebox/opti/preparation.py:373-379 with hessian_approximation = False,
    awebox/opts/default.py:323; the Lagrangian the reference writes down itself: awebox/tools/debug_operations.py:567-596)
    -> tests/golden/<config>_nk<N>_hess.npz.  The compressed sweeps are cross-checked here against uncompressed
    Hessian-times-direction products of the same tape."""
    t0 = time.time()
    T = ref_build.build(config, n_k=n_k)
    import casadi as cas
    from casadi import _eval
    nlp = T.nlp
    n_v = len(T.w0)
    g_fun = nlp.g_fun
    ins = [v for s in g_fun._ins for v in s.e]
    outs = [v for o in g_fun._outs for v in o.e]
    LH = _eval.LagrangianHessian(ins, outs, n_v)
    rng = np.random.default_rng(20261021 + n_k)
    W = perturbed_points(T.w0, n_pts, 20261018 + n_k)
    p = T.p0
    LAM = rng.uniform(-1.0, 1.0, (n_pts, LH.n_g))
    HV = np.array([LH.values(w, p, lam) for w, lam in zip(W, LAM)])
    # self-check of the compression: H . D for three dense directions
    D = rng.uniform(-1.0, 1.0, (n_v, 3))
    HD = LH.times(W[-1], p, LAM[-1], D)
    import scipy.sparse as sps
    U = sps.csc_matrix((HV[-1], LH.row, LH.colind), shape=(n_v, n_v))
    full = U + sps.triu(U, 1).T
    err = np.abs(full @ D - HD).max() / max(1.0, np.abs(HD).max())
    assert err < 1e-12, err
    out = os.path.join(GOLDEN, '%s_nk%d_hess.npz' % (config, n_k))
    np.savez_compressed(out, w=W, p=p, lam=LAM, colind=LH.colind, row=LH.row, hv=HV, dirs=D, hd=HD)
    print('%s n_k=%d hess_lag (constraint part): n_v=%d nnz(triu)=%d colours=%d dense columns=%d  check %.1e  (%.1f s) -> %s'
          % (config, n_k, n_v, len(LH.row), LH.n_colours, len(LH.head), err, time.time() - t0, os.path.relpath(out, ROOT)))


def run_hessian_full(config, n_k, n_pts):
    """golden vectors of the complete ``hess_lag`` of nlpsol: triu of the Hessian of gamma = lam_f f + lam_g' g with
    respect to V, from the reference's own ``nlp.f_fun`` and ``nlp.g_fun`` (awebox/opti/preparation.py:373-379), with P
    filled as before a solve (references, weights, distinct cost factors; as run_objective)
    -> tests/golden/<config>_nk<N>_hessf.npz"""
    t0 = time.time()
    T = ref_build.build(config, n_k=n_k)
    import casadi as cas
    from casadi import _eval
    import awebox.opti.preparation as prep
    nlp = T.nlp
    n_v = len(T.w0)
    rng = np.random.default_rng(20261022 + n_k)
    V_ref = nlp.V(T.w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, n_v)) + 1e-3 * rng.uniform(-1.0, 1.0, n_v))
    pfix = prep.add_weights_and_refs_to_opti_parameters(T.p_fix_num, V_ref, nlp, T.model, T.V_init, T.options['solver'])
    p = np.array(pfix.cat.full()).ravel()
    labs = nlp.P.labels()
    cost_idx = [i for i, l in enumerate(labs) if l.startswith('[cost,')]
    for n, i in enumerate(cost_idx):
        p[i] = 0.5 + 0.1 * n
    W = np.array([T.w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, n_v)) + 1e-3 * rng.uniform(-1.0, 1.0, n_v)
                  for _ in range(n_pts)])
    phi_idx = [i for i, l in enumerate(nlp.V.labels()) if l.startswith('[phi,')]
    W[:, phi_idx] = rng.uniform(0.2, 0.8, (n_pts, len(phi_idx)))
    g_fun, f_fun = nlp.g_fun, nlp.f_fun
    ins = [v for s in g_fun._ins for v in s.e]
    assert all(a is b for sa, sb in zip(g_fun._ins, f_fun._ins) for a, b in zip(sa.e, sb.e))
    outs = [v for o in g_fun._outs for v in o.e]
    LH = _eval.LagrangianHessian(ins, outs, n_v, f_output=f_fun._outs[0].e[0])
    LAM = rng.uniform(-1.0, 1.0, (n_pts, LH.n_g))
    SIG = rng.uniform(0.5, 1.5, n_pts)
    HV = np.array([LH.values(w, p, lam, sg) for w, lam, sg in zip(W, LAM, SIG)])
    D = rng.uniform(-1.0, 1.0, (n_v, 3))
    HD = LH.times(W[-1], p, LAM[-1], D, SIG[-1])
    import scipy.sparse as sps
    U = sps.csc_matrix((HV[-1], LH.row, LH.colind), shape=(n_v, n_v))
    full = U + sps.triu(U, 1).T
    err = np.abs(full @ D - HD).max() / max(1.0, np.abs(HD).max())
    assert err < 1e-12, err
    out = os.path.join(GOLDEN, '%s_nk%d_hessf.npz' % (config, n_k))
    np.savez_compressed(out, w=W, p=p, lam=LAM, lam_f=SIG, colind=LH.colind, row=LH.row, hv=HV)
    print('%s n_k=%d hess_lag (objective + constraints): n_v=%d nnz(triu)=%d colours=%d dense columns=%d  check %.1e  (%.1f s) -> %s'
          % (config, n_k, n_v, len(LH.row), LH.n_colours, len(LH.head), err, time.time() - t0, os.path.relpath(out, ROOT)))


HESSF_JOBS = {'ampyx_ap2': [(4, 3), (40, 2)], 'dual_kites': [(3, 2), (20, 2)]}

HESS_JOBS = {'ampyx_ap2': [(4, 3), (40, 2)], 'dual_kites': [(3, 2), (20, 2)], 'mpc_3dof': [(3, 3), (10, 3)]}

OUT_JOBS = {'ampyx_ap2': [(4, 3), (40, 5)], 'dual_kites': [(3, 2), (20, 5)], 'mpc_3dof': [(3, 3), (10, 5)]}


def run_outputs(config, n_k, n_pts):
    """golden vectors of the outputs map: nlp.Outputs_fun(V, P), the horizontal
    concatenation of model.outputs_fun over the shooting and collocation nodes of every interval
    (awebox/ocp/collocation.py:366-381, awebox/ocp/discretization.py:316-334), with the labels of
    model.outputs -> tests/golden/<config>_nk<N>_out.npz: w [n, n_v], p [n_p], out [n, n_nodes, n_out]."""
    t0 = time.time()
    T = ref_build.build(config, n_k=n_k)
    import casadi as cas
    nlp = T.nlp
    rng = np.random.default_rng(20261020 + n_k)
    n_v = len(T.w0)
    W = np.array([T.w0] + [T.w0 * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, n_v)) + 1e-3 * rng.uniform(-1.0, 1.0, n_v)
                           for _ in range(n_pts - 1)])
    phi_idx = [i for i, l in enumerate(nlp.V.labels()) if l.startswith('[phi,')]
    W[1:, phi_idx] = rng.uniform(0.2, 0.8, (n_pts - 1, len(phi_idx)))
    p = T.p0
    OUT = []
    for w in W:
        o = np.array(nlp.Outputs_fun(cas.DM(w), cas.DM(p)).full())
        OUT.append(o.T.copy())                      # [n_nodes, n_out]
    labels = T.model.outputs.labels()
    out = os.path.join(GOLDEN, '%s_nk%d_out.npz' % (config, n_k))
    np.savez_compressed(out, w=W, p=p, out=np.array(OUT), labels=np.array(labels))
    print('%s n_k=%d outputs: %s  (%.1f s) -> %s' % (config, n_k, np.array(OUT).shape, time.time() - t0,
                                                   os.path.relpath(out, ROOT)))


OBJ_JOBS = {'ampyx_ap2': [(4, 3), (40, 3)], 'dual_kites': [(3, 2), (20, 2)], 'mpc_3dof': [(3, 3), (10, 3)]}


if __name__ == '__main__':
    import logging
    logging.disable(logging.CRITICAL)
    args = [a for a in sys.argv[1:] if a not in ('--objective', '--hessian', '--outputs')]
    if '--outputs' in sys.argv[1:]:
        small = '--small' in args
        args = [a for a in args if a != '--small']
        for config in (args or list(OUT_JOBS)):
            for n_k, n_pts in (OUT_JOBS[config][:1] if small else OUT_JOBS[config]):
                run_outputs(config, n_k, n_pts)
        sys.exit(0)
    if '--hessian' in sys.argv[1:]:
        small = '--small' in args
        args = [a for a in args if a != '--small']
        full = '--full' in args
        args = [a for a in args if a != '--full']
        jobs = HESSF_JOBS if full else HESS_JOBS
        for config in (args or list(jobs)):
            for n_k, n_pts in (jobs[config][:1] if small else jobs[config]):
                (run_hessian_full if full else run_hessian)(config, n_k, n_pts)
        sys.exit(0)
    if '--objective' in sys.argv[1:]:
        for config in (args or list(OBJ_JOBS)):
            for n_k, n_pts in OBJ_JOBS[config]:
                (run_mpc_objective if config == 'mpc_3dof' else run_objective)(config, n_k, n_pts)
        sys.exit(0)
    for config in (args or list(JOBS)):
        for i, (n_k, n_pts) in enumerate(JOBS[config]):
            run(config, n_k, n_pts, write_spec=(i == len(JOBS[config]) - 1))
