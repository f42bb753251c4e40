"""Outputs map: nlp.Outputs_fun(V, P) against the reference run.

Golden vectors: tests/golden/*_out.npz (oracle/make_golden.py --outputs): the reference's own ``nlp.Outputs_fun`` at its
initial guess and at perturbed points, with the labels of ``model.outputs``.  The reference addresses outputs by label
(its index order follows Python set iteration), so entries are matched by label; every label of the reference must be
present.

CPU: the expressions of awebox_b200/mdl/outputs.py evaluated by the host restatement (OutputsMap.evaluate).
GPU: awegpu_eval_outputs through the C ABI against the golden vectors, against the host restatement on seeded batches
(ragged sizes, parameter modes, device entry) and, at BASELINE sizes, against properties: the 'model_equalities' rows of
the collocation nodes are the collocation rows of g, the inequalities of the shooting nodes its path rows, airspeed = |air_velocity|, xdot of the shooting nodes is V[xdot].

Tolerance: |a - b| <= 1e-10 max(|ref|, S_name, 1e-6 S_group), S_name = largest |ref| over the nodes and components of the
same name, S_group the same over the group: relative to the entry, with a floor for entries that are cancellation
noise at the test points (d c / d t at a consistent point, the side force of the 3-DOF kite, the Frenet centre
velocity = difference of 20 m/s terms); at the unperturbed initial guess (point 0) an absolute 1e-10 is added."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, get_ocp

CASES = [('ampyx_ap2', 4), ('ampyx_ap2', 40), ('dual_kites', 3), ('dual_kites', 20), ('mpc_3dof', 3), ('mpc_3dof', 10)]
SMALL = [('ampyx_ap2', 4), ('dual_kites', 3), ('mpc_3dof', 3), ('mpc_3dof', 10)]
_MAP = {}


def outputs_map(name, n_k):
    from awebox_b200.ocp.outputs import OutputsMap
    if (name, n_k) not in _MAP:
        _MAP[(name, n_k)] = OutputsMap(get_ocp(name, n_k))
    return _MAP[(name, n_k)]


def tolerance(labels, ref, consistent_point=False, rtol=1e-10):
    """ref [n_out, n_nodes] in the order of ``labels`` -> tolerance of the same shape.  ``consistent_point``: the
    unperturbed initial guess, where whole outputs (d c / d t, xdot - x') vanish up to the rounding of their operands:
    an absolute 1e-10 is added there (xdot of the collocation polynomial is a sum of terms 100 times larger than its value,
    and it is multiplied by masses and speeds); the perturbed points are judged without it"""
    names = [l[1:-1].rsplit(',', 1)[0] for l in labels]
    groups = [n.split(',')[0] for n in names]
    mag = np.max(np.abs(ref), axis=1)
    s_name, s_group = {}, {}
    for n, g, v in zip(names, groups, mag):
        s_name[n] = max(s_name.get(n, 0.0), v)
        s_group[g] = max(s_group.get(g, 0.0), v)
    floor = np.array([max(s_name[n], 1e-6 * s_group[g]) for n, g in zip(names, groups)])
    return rtol * np.maximum(np.abs(ref), floor[:, None]) + (1e-10 if consistent_point else 0.0)


def by_label(gold, labels, pt):
    """golden values of point ``pt`` reordered to ``labels``: [n_out, n_nodes]"""
    gi = {str(l): i for i, l in enumerate(gold['labels'])}
    return gold['out'][pt][:, [gi[l] for l in labels]].T


@pytest.mark.parametrize('name,n_k', CASES)
def test_every_reference_label_is_present(name, n_k):
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_out.npz' % (name, n_k)))
    om = outputs_map(name, n_k)
    assert sorted(om.labels) == sorted(str(l) for l in gold['labels'])
    assert len(set(om.labels)) == om.n_out
    assert gold['out'].shape[1:] == (om.n_nodes, om.n_out)


@pytest.mark.parametrize('name,n_k', SMALL)
def test_host_restatement_against_reference(name, n_k):
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_out.npz' % (name, n_k)))
    om = outputs_map(name, n_k)
    for pt in range(gold['w'].shape[0]):
        out = om.evaluate(gold['w'][pt], gold['p']).T
        ref = by_label(gold, om.labels, pt)
        assert np.all(np.isfinite(ref))
        bad = np.abs(out - ref) > tolerance(om.labels, ref, pt == 0)
        assert not bad.any(), [om.labels[i] for i in np.unique(np.nonzero(bad)[0])][:10]


def test_library_labels_match_the_expressions():
    """the compiled label table is the one of awebox_b200/mdl/outputs.py (no GPU needed: table lookups only)"""
    from awebox_b200.evaluator import load_library
    for name, n_k in (('ampyx_ap2', 4), ('dual_kites', 3), ('mpc_3dof', 3)):
        lib = load_library(name)
        om = outputs_map(name, n_k)
        assert int(lib.awegpu_outputs_n()) == om.n_out
        assert [lib.awegpu_outputs_label(i).decode() for i in range(om.n_out)] == om.labels
        assert lib.awegpu_outputs_label(om.n_out) is None


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k', CASES)
def test_gpu_outputs_against_reference(name, n_k):
    from awebox_b200.evaluator import Evaluator
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_out.npz' % (name, n_k)))
    ev = Evaluator(name, n_k=n_k)
    labels = ev.outputs_labels()
    assert sorted(labels) == sorted(str(l) for l in gold['labels'])
    out = ev.eval_outputs(gold['w'], gold['p'])
    assert out.shape == (gold['w'].shape[0], len(labels), n_k * (ev.ocp.d + 1))
    worst = {False: 0.0, True: 0.0}
    for pt in range(gold['w'].shape[0]):
        ref = by_label(gold, labels, pt)
        err = np.abs(out[pt] - ref)
        tol = tolerance(labels, ref, pt == 0)
        bad = err > tol
        assert not bad.any(), [labels[i] for i in np.unique(np.nonzero(bad)[0])][:10]
        worst[pt == 0] = max(worst[pt == 0], float(np.max(err[tol > 0] / tol[tol > 0])))
        # single-instance entry, the way the reference calls Outputs_fun(V, P)
        one = ev.outputs_fun(gold['w'][pt], gold['p'])
        assert np.array_equal(one, out[pt])
    print('%s n_k=%d: %d outputs x %d nodes, %d points: largest error / tolerance %.2e at the perturbed points '
          '(tolerance = 1e-10 of the scale), %.2e at the unperturbed initial guess (with the absolute 1e-10)'
          % (name, n_k, len(labels), out.shape[2], gold['w'].shape[0], worst[False], worst[True]))


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k,B', [('ampyx_ap2', 4, 37), ('dual_kites', 3, 5), ('mpc_3dof', 3, 33)])
def test_gpu_outputs_batch_against_host_restatement(name, n_k, B):
    """ragged batch, per-instance theta0, theta0-only and full parameter vectors, device entry"""
    import torch
    from awebox_b200.evaluator import Evaluator
    from awebox_b200 import synthetic
    ev = Evaluator(name, n_k=n_k)
    ocp = ev.ocp
    om = outputs_map(name, n_k)
    W = synthetic.batch(ocp, B, seed=11)
    rng = np.random.default_rng(5)
    p0 = ocp.default_p()
    P = np.tile(p0, (B, 1))
    P[:, ocp.p_theta0:] *= 1.0 + 1e-3 * rng.uniform(-1.0, 1.0, (B, ocp.n_theta0))
    out_full = ev.eval_outputs(W, P)
    out_th0 = ev.eval_outputs(W, np.ascontiguousarray(P[:, ocp.p_theta0:]))
    assert np.array_equal(out_full, out_th0)
    out_shared = ev.eval_outputs(W, p0)
    for b in (0, B // 2, B - 1):
        # synthetic points are not trajectories: outputs that vanish on one (lift of the 3-DOF kite at zero roll rate ...)
        # are differences of large terms here, hence 1e-9 of the scale
        ref = om.evaluate(W[b], P[b]).T
        assert np.all(np.abs(out_full[b] - ref) <= tolerance(om.labels, ref, rtol=1e-9))
        ref = om.evaluate(W[b], p0).T
        assert np.all(np.abs(out_shared[b] - ref) <= tolerance(om.labels, ref, rtol=1e-9))
    dev = torch.device('cuda:0')
    Wd, Pd = torch.from_numpy(W).to(dev), torch.from_numpy(P).to(dev)
    Od = torch.empty(out_full.shape, dtype=torch.float64, device=dev)
    ev.eval_outputs_device(Wd, Pd, Od)
    torch.cuda.synchronize()
    assert np.array_equal(Od.cpu().numpy(), out_full)
    assert ev.last_outputs_ms() > 0.0
    assert ev.eval_outputs(W[:0], p0).shape[0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k,B', [('ampyx_ap2', 40, 512), ('dual_kites', 20, 256), ('mpc_3dof', 10, 2048)])
def test_gpu_outputs_properties_at_baseline_sizes(name, n_k, B):
    """size-independent properties: the equality rows of the collocation nodes are the collocation rows of g (two
    different kernels and programs), airspeed = |air_velocity|, xdot of a shooting node is V[xdot, k], instances do not
    leak into each other (permutation)"""
    from awebox_b200.evaluator import Evaluator
    from awebox_b200 import synthetic
    ev = Evaluator(name, n_k=n_k)
    ocp = ev.ocp
    d = ocp.d
    W = synthetic.batch(ocp, B, seed=3)
    p = ocp.default_p()
    out = ev.eval_outputs(W, p)
    labels = ev.outputs_labels()
    G, _ = ev.eval(W, p, jac=False)
    eq = ev.outputs_index('model_equalities')
    assert len(eq) == ocp.n_eq
    scale = float(ocp.constraint_scale)
    for k in (0, n_k // 2, n_k - 1):
        for j in range(d):
            node = k * (d + 1) + 1 + j
            r0 = int(ocp.row_coll[k][j])
            g_rows = G[:, r0:r0 + ocp.n_eq]
            o_rows = out[:, eq, node] * scale
            tol = 1e-10 * np.maximum(1.0, np.max(np.abs(g_rows), axis=0, keepdims=True))
            assert np.all(np.abs(g_rows - o_rows) <= tol)
    # ... and the path inequalities of the shooting nodes are the path rows of g (zoh controls: constraints.py:261-268)
    iq = ev.outputs_index('model_inequalities')
    assert len(iq) == ocp.n_ineq
    for k in (0, n_k // 2, n_k - 1):
        r0 = int(ocp.row_path[k])
        g_rows = G[:, r0:r0 + ocp.n_ineq]
        o_rows = out[:, iq, k * (d + 1)] * scale
        tol = 1e-10 * np.maximum(1.0, np.max(np.abs(g_rows), axis=0, keepdims=True))
        assert np.all(np.abs(g_rows - o_rows) <= tol)
    for kite in ocp.model.arch.kite_nodes:
        va = out[:, ev.outputs_index('aerodynamics', 'air_velocity%d' % kite), :]
        sp = out[:, ev.outputs_index('aerodynamics', 'airspeed%d' % kite)[0], :]
        assert np.allclose(np.sqrt(np.sum(va * va, axis=1)), sp, rtol=1e-13, atol=0.0)
    # shooting nodes read the lifted xdot
    i0 = 0
    for nm, sz in ocp.model.var_list['xdot']:
        idx = ev.outputs_index('xdot_from_var', nm)
        k0, _ = ocp.model.offset[('xdot', nm)]
        s = ocp.spec.scaling[k0:k0 + sz]
        for k in (0, n_k - 1):
            lifted = W[:, ocp.v_xdot[k] + i0:ocp.v_xdot[k] + i0 + sz] * s
            assert np.allclose(out[:, idx, k * (d + 1)], lifted, rtol=1e-14, atol=0.0)
        i0 += sz
    perm = np.random.default_rng(0).permutation(B)
    assert np.array_equal(ev.eval_outputs(W[perm], p), out[perm])
    assert np.all(np.isfinite(out))


def test_tether_moment_kat():
    """the reference's own known-answer test of the outputs map (test/units/test_model.py:255-318): with the tether
    attached at r_tether ('stick'), the moment of the holonomic constraint on a 6-DOF kite is
    lambda * r_tether x (R^T q).  Same inputs and the same analytic answer as the reference's test (its model is the
    drag-mode variant of the single 6-DOF kite; the tether moment does not involve the generators)."""
    import copy
    from awebox_b200.spec import ModelSpec
    from awebox_b200.mdl.model import NodeModel
    from awebox_b200.mdl.outputs import model_outputs
    from awebox_b200 import sym as S
    base = ModelSpec.load('ampyx_ap2')
    d = copy.deepcopy(base.d)
    d['options']['tether.attachment'] = 'stick'
    m = NodeModel(ModelSpec(d))
    r_tether = np.array([0.0, 0.0, -0.1])
    si = {('x', 'q10'): [130.644, 24.5223, 74.2863], ('x', 'dq10'): [-16.3061, -27.0514, 37.5959],
          ('x', 'omega10'): [-0.0181481, 0.275884, 1.5743],
          ('x', 'r10'): [0.271805, 0.334641, -0.902295, 0.0595685, 0.929945, 0.362839, 0.960506, -0.15237, 0.23283],
          ('x', 'delta10'): [0.0693676, 0.261684, -0.258425], ('x', 'l_t'): [152.184], ('z', 'lambda10'): [45.024],
          ('theta', 'diam_t'): [0.005], ('theta', 't_f'): [3.93805]}
    v = np.zeros(m.n_var)
    for (vt, nm), val in si.items():
        k, sz = m.offset[(vt, nm)]
        v[k:k + sz] = np.array(val) / m.spec.scaling[k:k + sz]
    params = np.concatenate([np.array(m.spec.theta0_values, dtype=float), np.ones(m.n_param - m.n_theta0)])
    for i in range(3):
        params[m.param_labels.index('[theta0,geometry,r_tether,%d]' % i)] = r_tether[i]
    entries = dict(model_outputs(m))
    exprs = [S._e(entries['[tether_moments,n10,%d]' % i]) for i in range(3)]
    found = np.array(S.evaluate(exprs, list(m.vars) + list(m.params), np.concatenate([v, params])))
    dcm = np.array(si[('x', 'r10')]).reshape(3, 3).T                  # column-major reshape, as cas.reshape
    true = si[('z', 'lambda10')][0] * np.cross(r_tether, dcm.T @ np.array(si[('x', 'q10')]))
    assert np.linalg.norm(true - found) / np.linalg.norm(true) < 1e-8


@pytest.mark.gpu
def test_casadi_externals_for_objective_and_outputs():
    """the symbols casadi.external('awegpu_f', lib) / ('awegpu_outputs', lib) bind, called with CasADi's conventions:
    f(V, P) and jac_awegpu_f(V, P, out_f) -> (jac_f_V dense 1 x n_v, jac_f_P empty) against the golden vectors of the
    reference's f_fun; Outputs_fun(V, P) -> dense n_out x n_nodes in column-major non-zero order against the batched
    entry"""
    import ctypes as C
    from awebox_b200.evaluator import Evaluator
    ev = Evaluator('ampyx_ap2', n_k=4)
    lib, ocp = ev.lib, ev.ocp
    assert lib.awegpu_casadi_bind(ev.handle) == 0
    PD, PLL = C.POINTER(C.c_double), C.POINTER(C.c_longlong)
    for f in (lib.awegpu_f_n_in, lib.awegpu_f_n_out, lib.jac_awegpu_f_n_in, lib.jac_awegpu_f_n_out, lib.awegpu_outputs_n_in,
              lib.awegpu_outputs_n_out):
        f.restype = C.c_longlong
    assert (lib.awegpu_f_n_in(), lib.awegpu_f_n_out(), lib.jac_awegpu_f_n_in(), lib.jac_awegpu_f_n_out()) == (2, 1, 3, 2)
    assert (lib.awegpu_outputs_n_in(), lib.awegpu_outputs_n_out()) == (2, 1)
    for f in (lib.awegpu_f_sparsity_in, lib.awegpu_f_sparsity_out, lib.jac_awegpu_f_sparsity_out, lib.awegpu_outputs_sparsity_out,
              lib.jac_awegpu_f_sparsity_in):
        f.restype, f.argtypes = PLL, [C.c_longlong]
    sp = lib.awegpu_f_sparsity_in(1)
    assert (sp[0], sp[1], sp[2]) == (ocp.n_p, 1, 1)
    sp = lib.awegpu_f_sparsity_out(0)
    assert (sp[0], sp[1], sp[2]) == (1, 1, 1)
    sp = lib.jac_awegpu_f_sparsity_out(0)
    assert (sp[0], sp[1]) == (1, ocp.n_v) and sp[2 + ocp.n_v] == ocp.n_v and sp[2 + ocp.n_v + 1 + 5] == 0
    sp = lib.jac_awegpu_f_sparsity_out(1)
    assert (sp[0], sp[1]) == (1, ocp.n_p) and sp[2 + ocp.n_p] == 0
    sp = lib.jac_awegpu_f_sparsity_in(2)
    assert (sp[0], sp[1], sp[2]) == (1, 1, 1)
    n_out, n_nodes = len(ev.outputs_labels()), ocp.n_k * (ocp.d + 1)
    sp = lib.awegpu_outputs_sparsity_out(0)
    assert (sp[0], sp[1], sp[2]) == (n_out, n_nodes, 1)
    sig = [C.POINTER(PD), C.POINTER(PD), PLL, PD, C.c_int]
    lib.awegpu_f.argtypes = lib.jac_awegpu_f.argtypes = lib.awegpu_outputs.argtypes = sig
    gold = np.load(os.path.join(GOLDEN, 'ampyx_ap2_nk4_obj.npz'))
    w, p = np.ascontiguousarray(gold['w'][1]), np.ascontiguousarray(gold['p'])
    fv = np.empty(1)
    arg = (PD * 3)(w.ctypes.data_as(PD), p.ctypes.data_as(PD), fv.ctypes.data_as(PD))
    res = (PD * 2)(fv.ctypes.data_as(PD), None)
    assert lib.awegpu_f(arg, res, None, None, 0) == 0
    assert abs(fv[0] - gold['f'][1]) <= 1e-11 * abs(gold['f'][1])
    gr = np.empty(ocp.n_v)
    res = (PD * 2)(gr.ctypes.data_as(PD), None)
    assert lib.jac_awegpu_f(arg, res, None, None, 0) == 0
    g_ref = gold['grad'][1]
    assert np.all(np.abs(gr - g_ref) <= np.maximum(1e-11 * np.abs(g_ref), 1e-13 * max(1.0, float(np.max(np.abs(g_ref))))))
    go = np.load(os.path.join(GOLDEN, 'ampyx_ap2_nk4_out.npz'))
    w, p = np.ascontiguousarray(go['w'][1]), np.ascontiguousarray(go['p'])
    nz = np.empty(n_out * n_nodes)
    arg = (PD * 2)(w.ctypes.data_as(PD), p.ctypes.data_as(PD))
    res = (PD * 1)(nz.ctypes.data_as(PD))
    assert lib.awegpu_outputs(arg, res, None, None, 0) == 0
    assert np.array_equal(nz.reshape(n_nodes, n_out).T, ev.outputs_fun(w, p))       # column-major non-zeros
    arg0 = (PD * 2)(None, p.ctypes.data_as(PD))
    assert lib.awegpu_outputs(arg0, res, None, None, 0) != 0
