"""Second-order hook: upper triangle of d^2/dw^2 (lam' g(w, p)).

CPU: the host restatement (awebox_b200/ocp/hessian.py: node Hessians from the product's own expression graph, constant
assembly matrix) against the golden vectors made from the reference's own ``nlp.g_fun`` the way casadi's nlpsol forms
``nlp_hess_l`` (tests/golden/*_hess.npz, oracle/make_golden.py --hessian): pattern bit-exact, values relative to the
largest entry of the row / column they sit in.
GPU: awegpu_eval_hessian through the C ABI against the same golden vectors, against the host restatement on seeded
batches (ragged tiles), symmetry-free properties at the BASELINE sizes (Hessian-times-direction against the golden
products, linearity in lam)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, get_ocp

CASES = [('ampyx_ap2', 4), ('ampyx_ap2', 40), ('dual_kites', 3), ('dual_kites', 20), ('mpc_3dof', 3), ('mpc_3dof', 10)]
_H = {}


def hessian(name, n_k):
    from awebox_b200.ocp.hessian import LagrangianHessian
    if (name, n_k) not in _H:
        _H[(name, n_k)] = LagrangianHessian(get_ocp(name, n_k))
    return _H[(name, n_k)]


def hess_tolerance(H, ref, rtol=1e-12):
    """|a - b| <= rtol * max(|ref|, largest |entry| of the entry's row and of its column in the full symmetric matrix):
    entries are sums over the nodes of an interval (and over all nodes for the theta / phi block) of terms of that
    size."""
    n_v = len(H.colind) - 1
    col = np.repeat(np.arange(n_v), np.diff(H.colind))
    scale = np.zeros(n_v)
    np.maximum.at(scale, H.row, np.abs(ref))
    np.maximum.at(scale, col, np.abs(ref))
    return rtol * np.maximum(np.abs(ref), np.maximum(scale[H.row], scale[col]))


@pytest.mark.parametrize('name,n_k', CASES)
def test_pattern_bit_exact_and_values_against_reference(name, n_k):
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_hess.npz' % (name, n_k)))
    H = hessian(name, n_k)
    assert H.colind.dtype == np.int64 and H.row.dtype == np.int64
    assert np.array_equal(H.colind, gold['colind'])
    assert np.array_equal(H.row, gold['row'])
    for w, lam, ref in zip(gold['w'], gold['lam'], gold['hv']):
        hv = H.evaluate(w, gold['p'], lam)
        assert np.all(np.abs(hv - ref) <= hess_tolerance(H, ref))


@pytest.mark.parametrize('name,n_k', [('ampyx_ap2', 4), ('dual_kites', 3), ('mpc_3dof', 3)])
def test_hessian_is_the_derivative_of_jT_lambda(name, n_k):
    """independent of the golden vectors: central differences of J(w)' lam (checker) along a direction against H . dir"""
    from conftest import get_oracle
    import scipy.sparse as sps
    ocp = get_ocp(name, n_k)
    H = hessian(name, n_k)
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_hess.npz' % (name, n_k)))
    w, p, lam = gold['w'][0], gold['p'], gold['lam'][0]
    orc = get_oracle(name, n_k)
    rng = np.random.default_rng(5)
    dirn = rng.uniform(-1.0, 1.0, ocp.n_v)

    def jtl(wv):
        _, jv = orc.eval(wv[None, :], p)
        J = sps.csc_matrix((jv[0], ocp.row, ocp.colind), shape=(ocp.n_g, ocp.n_v))
        return J.T @ lam
    eps = 1e-6
    fd = (jtl(w + eps * dirn) - jtl(w - eps * dirn)) / (2 * eps)
    U = sps.csc_matrix((H.evaluate(w, p, lam), H.row, H.colind), shape=(ocp.n_v, ocp.n_v))
    hd = (U + sps.triu(U, 1).T) @ dirn
    assert np.abs(hd - fd).max() <= 1e-6 * max(1.0, np.abs(hd).max())
