"""Objective path: f(w, p) and its gradient against the reference run.

CPU: the host restatement of awebox/ocp/objective.py (awebox_b200/ocp/objective.py, which also produces the layout tables
of the CUDA path) against the golden vectors of the reference's own ``nlp.f_fun`` and its gradient
(tests/golden/*_obj.npz, oracle/make_golden.py --objective), component by component.
GPU: awegpu_eval_objective through the C ABI against the same golden vectors and against the host restatement on a seeded
batch."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, get_ocp

CASES = [('ampyx_ap2', 4), ('ampyx_ap2', 40), ('dual_kites', 3), ('dual_kites', 20)]
_OBJ = {}


def objective(name, n_k):
    from awebox_b200.ocp.objective import Objective
    if (name, n_k) not in _OBJ:
        _OBJ[(name, n_k)] = Objective(get_ocp(name, n_k))
    return _OBJ[(name, n_k)]


def grad_tolerance(g_ref):
    # relative to the entry, with a floor at 1e-13 of the largest entry: entries of grad f are sums of terms of the size of
    # the largest entries (the regularisation terms 2 w (v - ref) of one node feed several columns through the
    # collocation polynomial and t_f)
    return np.maximum(1e-11 * np.abs(g_ref), 1e-13 * max(1.0, float(np.max(np.abs(g_ref)))))


@pytest.mark.parametrize('name,n_k', CASES)
def test_node_integrands_against_reference(name, n_k):
    """power integrand (model.integral_outputs_fun) and side-slip angles (model outputs) at random node inputs"""
    from awebox_b200 import sym as S
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (name, n_k)))
    obj = objective(name, n_k)
    m = obj.ocp.model
    flat, layout = obj.integrands.flat_outputs()
    n_beta = len(layout) - 1
    for v, pw, bt in zip(gold['node_vars'], gold['node_integral'], gold['node_beta']):
        vals = S.evaluate(flat, list(m.vars) + list(m.params), np.concatenate([v, gold['node_p']]))
        assert abs(vals[layout[0][0]] - pw[0]) <= 1e-12 * max(1.0, abs(pw[0]))
        ref_beta = [bt[i] for i, lab in enumerate(gold['beta_labels']) if 'beta_deg' not in str(lab)]
        assert len(ref_beta) == n_beta
        for (first, _), rb in zip(layout[1:], ref_beta):
            assert abs(vals[first] - rb) <= 1e-12 * max(1.0, abs(rb))


@pytest.mark.parametrize('name,n_k', CASES)
def test_host_restatement_against_reference(name, n_k):
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (name, n_k)))
    obj = objective(name, n_k)
    names = [str(n) for n in gold['component_names']]
    for i in range(len(gold['f'])):
        f, g, comp = obj.evaluate(gold['w'][i], gold['p'], components=True)
        assert abs(f - gold['f'][i]) <= 1e-12 * max(1.0, abs(gold['f'][i]))
        for n, c in zip(names, gold['components'][i]):
            if n in comp:
                assert abs(comp[n] - c) <= 1e-12 * max(1.0, abs(c)), n
        assert np.all(np.abs(g - gold['grad'][i]) <= grad_tolerance(gold['grad'][i]))


def test_power_cycle_objective_is_refused_for_mpc_trials():
    from awebox_b200.ocp.objective import Objective, MpcObjective
    with pytest.raises(NotImplementedError):
        Objective(get_ocp('mpc_3dof', 3))
    with pytest.raises(NotImplementedError):
        MpcObjective(get_ocp('ampyx_ap2', 4))


MPC_CASES = [('mpc_3dof', 3), ('mpc_3dof', 10)]


@pytest.mark.parametrize('name,n_k', MPC_CASES)
def test_mpc_tracking_cost_restatement_against_reference(name, n_k):
    """awebox_b200/ocp/objective.py: MpcObjective against the reference's own Pmpc.__generate_objective
    (awebox/pmpc.py:304-358) run on the MPC trial: oracle/make_golden.py --objective mpc_3dof"""
    from awebox_b200.ocp.objective import MpcObjective
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (name, n_k)))
    mo = MpcObjective(get_ocp(name, n_k))
    assert mo.n_p == gold['p'].shape[1]
    labels = [str(l) for l in gold['p_labels']]
    assert labels[mo.p_ref].startswith('[ref,') and labels[mo.p_Q] == '[Q,0]' and labels[mo.p_R] == '[R,0]' and labels[mo.p_P] == '[P,0]'
    for w, p, f, g in zip(gold['w'], gold['p'], gold['f'], gold['grad']):
        f0, g0 = mo.evaluate(w, p)
        assert abs(f0 - f) <= 1e-13 * abs(f)
        assert np.array_equal(g0 != 0.0, g != 0.0)
        assert np.all(np.abs(g0 - g) <= 1e-13 * np.abs(g))


# ---------------------------------------------------------------------------------------------- CUDA path
def _evaluator(name, n_k):
    from awebox_b200.evaluator import Evaluator
    key = ('ev', name, n_k)
    if key not in _OBJ:
        _OBJ[key] = Evaluator(name, ocp=get_ocp(name, n_k))
    return _OBJ[key]


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k', CASES)
def test_cuda_objective_against_reference_golden_vectors(name, n_k):
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (name, n_k)))
    ev = _evaluator(name, n_k)
    F, G, comp = ev.eval_objective(gold['w'], gold['p'], components=True)
    names = [str(n) for n in gold['component_names']]
    for i in range(len(gold['f'])):
        assert abs(F[i] - gold['f'][i]) <= 1e-12 * max(1.0, abs(gold['f'][i]))
        assert np.all(np.abs(G[i] - gold['grad'][i]) <= grad_tolerance(gold['grad'][i]))
        for ci, cn in enumerate(ev.COMPONENTS):
            ref = gold['components'][i][names.index(cn + ('_cost' if cn != 'time' else '_cost'))]
            assert abs(comp[i, ci] - ref) <= 1e-12 * max(1.0, abs(ref)), cn


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k,B,shared', [('ampyx_ap2', 40, 70, True), ('ampyx_ap2', 4, 33, False), ('dual_kites', 3, 40, True)])
def test_cuda_objective_against_host_restatement_on_seeded_batch(name, n_k, B, shared):
    from awebox_b200 import synthetic
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (name, n_k)))
    ev = _evaluator(name, n_k)
    ocp = ev.ocp
    obj = objective(name, n_k)
    rng = np.random.default_rng(12)
    W = gold['w'][0][None, :] * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, (B, ocp.n_v)))
    W[:, ocp.v_phi:ocp.v_phi + ocp.n_phi] = rng.uniform(0.1, 0.9, (B, ocp.n_phi))
    if shared:
        P = gold['p']
    else:
        P = np.tile(gold['p'], (B, 1))
        P[:, obj.p_cost:obj.p_cost + obj.n_cost] *= 1.0 + 0.3 * rng.uniform(-1.0, 1.0, (B, obj.n_cost))
        P[:, :ocp.n_v] *= 1.0 + 1e-2 * rng.uniform(-1.0, 1.0, (B, ocp.n_v))
    F, G = ev.eval_objective(W, P)
    for b in range(0, B, 7):
        f0, g0 = obj.evaluate(W[b], P if shared else P[b])
        assert abs(F[b] - f0) <= 1e-12 * max(1.0, abs(f0))
        assert np.all(np.abs(G[b] - g0) <= grad_tolerance(g0))


@pytest.mark.gpu
def test_cuda_objective_single_instance_entry_points():
    gold = np.load(os.path.join(GOLDEN, 'ampyx_ap2_nk4_obj.npz'))
    ev = _evaluator('ampyx_ap2', 4)
    assert abs(ev.f_fun(gold['w'][1], gold['p']) - gold['f'][1]) <= 1e-12 * abs(gold['f'][1])
    assert np.all(np.abs(ev.grad_f_fun(gold['w'][1], gold['p']) - gold['grad'][1]) <= grad_tolerance(gold['grad'][1]))
    with pytest.raises(Exception):      # an MPC trial takes Pmpc's parameter vector, not the NLP's
        _evaluator('mpc_3dof', 3).eval_objective(np.zeros((1, get_ocp('mpc_3dof', 3).n_v)), np.zeros(get_ocp('mpc_3dof', 3).n_p))


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k', MPC_CASES)
def test_cuda_mpc_tracking_cost_against_reference_golden_vectors(name, n_k):
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_obj.npz' % (name, n_k)))
    ev = _evaluator(name, n_k)
    F, G = ev.eval_objective(gold['w'], gold['p'])            # one parameter vector per instance
    assert np.all(np.abs(F - gold['f']) <= 1e-13 * np.abs(gold['f']))
    assert np.all(np.abs(G - gold['grad']) <= 1e-13 * np.abs(gold['grad']))
    f1 = ev.f_fun(gold['w'][2], gold['p'][2])                 # one instance, shared parameter vector
    assert abs(f1 - gold['f'][2]) <= 1e-13 * abs(gold['f'][2])
    with pytest.raises(Exception):
        ev.eval_objective(gold['w'], gold['p'], components=True)


@pytest.mark.gpu
def test_cuda_mpc_tracking_cost_on_seeded_scenario_batch():
    """8192-scenario shape of BASELINE configs[4] (reduced batch): every scenario its own reference and weights"""
    from awebox_b200.ocp.objective import MpcObjective
    gold = np.load(os.path.join(GOLDEN, 'mpc_3dof_nk10_obj.npz'))
    ev = _evaluator('mpc_3dof', 10)
    mo = MpcObjective(ev.ocp)
    rng = np.random.default_rng(5)
    B = 300
    W = gold['w'][0][None, :] * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, (B, ev.n_v)))
    P = np.tile(gold['p'][0], (B, 1)) * (1.0 + 1e-1 * rng.uniform(-1.0, 1.0, (B, mo.n_p)))
    F, G = ev.eval_objective(W, P)
    for b in range(0, B, 13):
        f0, g0 = mo.evaluate(W[b], P[b])
        assert abs(F[b] - f0) <= 1e-13 * abs(f0)
        assert np.all(np.abs(G[b] - g0) <= 1e-13 * np.abs(g0))
    F2, G2 = ev.eval_objective(W, P[0])                       # shared parameters
    f0, g0 = mo.evaluate(W[7], P[0])
    assert abs(F2[7] - f0) <= 1e-13 * abs(f0) and np.all(np.abs(G2[7] - g0) <= 1e-13 * np.abs(g0))
