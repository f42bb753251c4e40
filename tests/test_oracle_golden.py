"""The CPU checker and the host-side assembly against the golden vectors produced by running the
reference's own Python here (oracle/make_golden.py).  Integer structure: bit-exact.  Values: 1e-12."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, get_ocp, get_oracle, jac_tolerance
from awebox_b200 import cabi

CASES = [('ampyx_ap2', 4), ('ampyx_ap2', 40), ('dual_kites', 3), ('dual_kites', 20), ('mpc_3dof', 3), ('mpc_3dof', 10)]


def load(name, n_k):
    path = os.path.join(GOLDEN, '%s_nk%d.npz' % (name, n_k))
    if not os.path.exists(path):
        pytest.skip('golden file %s not generated' % os.path.basename(path))
    return np.load(path)


@pytest.mark.parametrize('name,n_k', CASES)
def test_sizes_and_pattern_bit_exact(name, n_k):
    gold = load(name, n_k)
    ocp = get_ocp(name, n_k)
    assert ocp.n_v == gold['w'].shape[1]
    assert ocp.n_p == gold['p'].shape[0]
    assert ocp.n_g == gold['g'].shape[1]
    assert ocp.nnz == gold['row'].shape[0]
    assert ocp.colind.dtype == np.int64 and ocp.row.dtype == np.int64
    assert np.array_equal(ocp.colind, gold['colind'])
    assert np.array_equal(ocp.row, gold['row'])


def test_ap2_sizes_match_survey():
    ocp = get_ocp('ampyx_ap2', 40)
    assert (ocp.n_v, ocp.n_g) == (6154, 6103)
    assert (ocp.n_var, ocp.n_eq, ocp.n_ineq) == (59, 24, 9)


@pytest.mark.parametrize('name,n_k', CASES)
def test_model_names_and_variables(name, n_k):
    gold = load(name, n_k)
    m = get_ocp(name, n_k).model
    assert list(gold['var_labels']) == m.var_labels
    assert list(gold['eq_names']) == m.eq_names()
    assert list(gold['ineq_names']) == m.ineq_names()


@pytest.mark.parametrize('name,n_k', CASES)
def test_node_functions_against_reference(name, n_k):
    """Model.dynamics / inequality function of the reference at random node inputs."""
    from awebox_b200 import sym as S
    gold = load(name, n_k)
    m = get_ocp(name, n_k).model
    syms = np.concatenate([m.vars, m.params])
    for v, eq, ineq in zip(gold['node_vars'], gold['node_eq'], gold['node_ineq']):
        mine = S.evaluate(m.eq_exprs() + m.ineq_exprs(), syms, np.concatenate([v, gold['node_p']]))
        ref = np.concatenate([eq, ineq])
        assert np.all(np.abs(mine - ref) <= 1e-12 * np.maximum(1.0, np.abs(ref)))


@pytest.mark.parametrize('name,n_k', CASES)
def test_oracle_values_against_reference(name, n_k):
    gold = load(name, n_k)
    ocp = get_ocp(name, n_k)
    G, Jv = get_oracle(name, n_k).eval(gold['w'], gold['p'], flags=cabi.AWEGPU_P_SHARED)
    assert np.all(np.abs(G - gold['g']) <= 1e-12 * np.maximum(1.0, np.abs(gold['g'])))
    assert np.all(np.abs(Jv - gold['jv']) <= jac_tolerance(ocp, gold['jv']))
    # and the strictly relative statement holds for every entry that is not cancellation noise (an entry three
    # orders of magnitude below the largest entry of its row is the difference of terms of that size)
    rowmax = jac_tolerance(ocp, gold['jv'], rtol=1.0) / 1.0
    big = (np.abs(gold['jv']) > 1e-6) & (np.abs(gold['jv']) >= 1e-3 * rowmax)
    assert np.max(np.abs(Jv - gold['jv'])[big] / np.abs(gold['jv'])[big]) < 1e-10


def test_oracle_parameter_modes():
    ocp = get_ocp('ampyx_ap2', 4)
    orc = get_oracle('ampyx_ap2', 4)
    gold = load('ampyx_ap2', 4)
    W, p = gold['w'], gold['p']
    G0, J0 = orc.eval(W, p, flags=cabi.AWEGPU_P_SHARED)
    Pfull = np.tile(p, (W.shape[0], 1))
    G1, J1 = orc.eval(W, Pfull, flags=cabi.AWEGPU_P_FULL)
    G2, J2 = orc.eval(W, Pfull[:, ocp.p_theta0:], flags=cabi.AWEGPU_P_THETA0)
    assert np.array_equal(G0, G1) and np.array_equal(J0, J1)
    assert np.array_equal(G0, G2) and np.array_equal(J0, J2)


def test_mpc_initial_rows_read_the_reference_block_of_p():
    """trajectory type 'mpc': g starts with x_0 - p.ref[x_0] (awebox/ocp/operation.py:303-326) in the order the
    reference run produced; no periodicity rows; theta0-only parameter blocks are refused."""
    from awebox_b200 import synthetic
    ocp = get_ocp('mpc_3dof', 3)
    assert ocp.initial_conditions and not ocp.periodic and len(ocp.periodic_idx) == 0
    assert ocp.spec.nlp['initial_order'] == ['dq10', 'l_t', 'dl_t', 'q10', 'ddl_t', 'coeff10']
    W = synthetic.batch(ocp, 5, seed=2)
    P = synthetic.scenario_params(ocp, W, seed=3)
    G, _ = get_oracle('mpc_3dof', 3).eval(W, P, flags=cabi.AWEGPU_P_FULL)
    idx = np.array(ocp.initial_idx)
    want = W[:, ocp.v_x[0] + idx] - P[:, ocp.v_x[0] + idx]
    assert np.array_equal(G[:, :len(idx)], want)
    with pytest.raises(RuntimeError):
        get_oracle('mpc_3dof', 3).eval(W, P[:, ocp.p_theta0:], flags=cabi.AWEGPU_P_THETA0)
