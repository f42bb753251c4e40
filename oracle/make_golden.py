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


if __name__ == '__main__':
    import logging
    logging.disable(logging.CRITICAL)
    for config in (sys.argv[1:] or list(JOBS)):
        for i, (n_k, n_pts) in enumerate(JOBS[config]):
            run(config, n_k, n_pts, write_spec=(i == len(JOBS[config]) - 1))
