"""Second-order hook: upper triangle of d^2/dw^2 (lam_f f(w, p) + lam' g(w, p)).

CPU: the host restatement (awebox_b200/ocp/hessian.py: node Hessians from the product's own expression graph, constant
assembly matrix) against the golden vectors made from the reference's own ``nlp.g_fun`` the way casadi's nlpsol forms
``nlp_hess_l`` (tests/golden/*_hess.npz: constraint part, all configurations; *_hessf.npz: objective + constraints, the
power-cycle configurations; oracle/make_golden.py --hessian [--full]): pattern bit-exact, values relative to the largest entry
of the row / column they sit in.
GPU: awegpu_eval_hessian through the C ABI against the same golden vectors, against the host restatement on seeded
batches (ragged tiles), symmetry-free properties at the BASELINE sizes (Hessian-times-direction against the golden
products, linearity in lam)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, get_ocp

CASES = [('ampyx_ap2', 4), ('ampyx_ap2', 40), ('dual_kites', 3), ('dual_kites', 20), ('mpc_3dof', 3), ('mpc_3dof', 10)]
FULL_CASES = [('ampyx_ap2', 4), ('ampyx_ap2', 40), ('dual_kites', 3), ('dual_kites', 20)]
_H = {}


def hessian(name, n_k, objective=False):
    from awebox_b200.ocp.hessian import LagrangianHessian
    if (name, n_k, objective) not in _H:
        _H[(name, n_k, objective)] = LagrangianHessian(get_ocp(name, n_k), objective=objective)
    return _H[(name, n_k, objective)]


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


@pytest.mark.parametrize('name,n_k', FULL_CASES)
def test_complete_hess_lag_against_reference(name, n_k):
    """lam_f f + lam_g' g: pattern of the union bit-exact, values against the reference run"""
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_hessf.npz' % (name, n_k)))
    H = hessian(name, n_k, objective=True)
    assert np.array_equal(H.colind, gold['colind'])
    assert np.array_equal(H.row, gold['row'])
    for w, lam, sig, ref in zip(gold['w'], gold['lam'], gold['lam_f'], gold['hv']):
        hv = H.evaluate(w, gold['p'], lam, sig)
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


# ------------------------------------------------------------------------------------------------ CUDA path
_EV = {}


def evaluator(name, n_k):
    from awebox_b200.evaluator import Evaluator
    if (name, n_k) not in _EV:
        _EV[(name, n_k)] = Evaluator(name, ocp=get_ocp(name, n_k))
        _EV[(name, n_k)].hessian_setup()
    return _EV[(name, n_k)]


def embed(H_sub, H_full):
    """positions of the entries of the pattern H_sub inside the pattern H_full (H_sub must be a subset)"""
    n_v = len(H_full.colind) - 1
    col_f = np.repeat(np.arange(n_v), np.diff(H_full.colind))
    col_s = np.repeat(np.arange(n_v), np.diff(H_sub.colind))
    key_f = col_f * n_v + H_full.row
    key_s = col_s * n_v + H_sub.row
    pos = np.searchsorted(key_f, key_s)
    assert np.array_equal(key_f[pos], key_s)
    return pos


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k', CASES)
def test_gpu_against_reference_golden_vectors(name, n_k):
    """the complete hess_lag where the problem has an objective on this path, else the constraint part"""
    full = (name, n_k) in FULL_CASES
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_%s.npz' % (name, n_k, 'hessf' if full else 'hess')))
    ev = evaluator(name, n_k)
    assert ev.hess_objective == full
    colind, row = ev.hess_sparsity()
    assert colind.dtype == np.int64 and row.dtype == np.int64
    assert np.array_equal(colind, gold['colind']) and np.array_equal(row, gold['row'])
    H = ev.eval_hessian(gold['w'], gold['p'], gold['lam'], sigma=gold['lam_f'] if full else None)
    Hh = hessian(name, n_k, objective=full)
    for hv, ref in zip(H, gold['hv']):
        assert np.all(np.abs(hv - ref) <= hess_tolerance(Hh, ref))
    # one instance per call, the way nlpsol calls hess_lag
    one = ev.hess_lag_fun(gold['w'][0], gold['p'], gold['lam_f'][0] if full else 1.0, gold['lam'][0])
    assert np.array_equal(one, H[0])


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k', FULL_CASES)
def test_gpu_constraint_part_alone(name, n_k):
    """Sigma = NULL: the constraint part on the pattern of the complete Hessian -- the golden vectors of lam' g on their own
    pattern, exact zeros elsewhere; P may then be theta0 only"""
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_hess.npz' % (name, n_k)))
    ev = evaluator(name, n_k)
    Hg, Hf = hessian(name, n_k), hessian(name, n_k, objective=True)
    pos = embed(Hg, Hf)
    ocp = get_ocp(name, n_k)
    for P in (gold['p'], gold['p'][ocp.p_theta0:]):
        H = ev.eval_hessian(gold['w'], P, gold['lam'])
        rest = np.ones(Hf.nnz, dtype=bool)
        rest[pos] = False
        assert not np.any(H[:, rest])
        for hv, ref in zip(H[:, pos], gold['hv']):
            assert np.all(np.abs(hv - ref) <= hess_tolerance(Hg, ref))


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k,B', [('ampyx_ap2', 4, 70), ('dual_kites', 3, 37), ('mpc_3dof', 10, 97)])
def test_gpu_against_host_restatement_on_a_seeded_ragged_batch(name, n_k, B):
    from awebox_b200 import synthetic
    ocp = get_ocp(name, n_k)
    ev = evaluator(name, n_k)
    full = ev.hess_objective
    Hh = hessian(name, n_k, objective=full)
    W = synthetic.batch(ocp, B, seed=11)
    rng = np.random.default_rng(3)
    # parameter modes: one shared vector (AP2), per-instance theta0 (dual kites: a wind-speed sweep; constraint part only),
    # per-instance full vectors (MPC scenarios)
    sigma = None
    if name == 'dual_kites':
        P = synthetic.power_curve_params(ocp, B)
    elif name == 'mpc_3dof':
        P = synthetic.scenario_params(ocp, W, seed=12)
    else:
        P = ocp.default_p()
        P[:ocp.n_v] = W[1] * (1.0 + 1e-2 * rng.uniform(-1.0, 1.0, ocp.n_v))          # p.ref
        P[ocp.n_v:ocp.p_theta0] = rng.uniform(0.5, 1.5, ocp.p_theta0 - ocp.n_v)        # p.weights, cost factors
        sigma = rng.uniform(0.5, 1.5, B)
    Lam = rng.uniform(-1.0, 1.0, (B, ocp.n_g))
    H = ev.eval_hessian(W, P, Lam, sigma=sigma)
    for b in (0, 1, 31, 32, B - 1):
        p = P if P.ndim == 1 else P[b]
        ref = Hh.evaluate(W[b], p if len(p) == ocp.n_p else np.concatenate([np.zeros(ocp.p_theta0), p]), Lam[b],
                          0.0 if sigma is None else sigma[b])
        assert np.all(np.abs(H[b] - ref) <= hess_tolerance(Hh, ref)), b
    # deterministic: a second call gives the same bits
    assert np.array_equal(H, ev.eval_hessian(W, P, Lam, sigma=sigma))


@pytest.mark.gpu
@pytest.mark.parametrize('name,n_k,B', [('ampyx_ap2', 40, 256), ('dual_kites', 20, 128), ('mpc_3dof', 10, 512)])
def test_gpu_properties_at_full_size(name, n_k, B):
    """linear in (lam_f, lam); zero multipliers give a zero Hessian; instances do not depend on their position in the batch"""
    from awebox_b200 import synthetic
    ocp = get_ocp(name, n_k)
    ev = evaluator(name, n_k)
    W, P = synthetic.batch(ocp, B, seed=12), ocp.default_p()
    rng = np.random.default_rng(4)
    full = ev.hess_objective
    if full:
        P[:ocp.n_v] = W[0]
        P[ocp.n_v:ocp.p_theta0] = rng.uniform(0.5, 1.5, ocp.p_theta0 - ocp.n_v)
    s1, s2 = (rng.uniform(0.5, 1.5, B), rng.uniform(0.5, 1.5, B)) if full else (None, None)
    L1, L2 = rng.uniform(-1.0, 1.0, (B, ocp.n_g)), rng.uniform(-1.0, 1.0, (B, ocp.n_g))
    H1, H2 = ev.eval_hessian(W, P, L1, sigma=s1), ev.eval_hessian(W, P, L2, sigma=s2)
    H12 = ev.eval_hessian(W, P, L1 + 2.0 * L2, sigma=s1 + 2.0 * s2 if full else None)
    scale = np.maximum(1.0, np.abs(H1).max() + 2.0 * np.abs(H2).max())
    assert np.abs(H12 - (H1 + 2.0 * H2)).max() <= 1e-12 * scale
    assert not np.any(ev.eval_hessian(W[:3], P, np.zeros((3, ocp.n_g)), sigma=np.zeros(3) if full else None))
    perm = rng.permutation(B)
    Hp = ev.eval_hessian(W[perm], P, L1[perm], sigma=s1[perm] if full else None)
    assert np.array_equal(Hp, H1[perm])


@pytest.mark.gpu
def test_casadi_hess_lag_symbols():
    """awegpu_hess_lag called with CasADi's external-function conventions (arg / res pointer arrays, compact CCS)"""
    import ctypes as C
    name, n_k = 'ampyx_ap2', 4
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d_hessf.npz' % (name, n_k)))
    ev = evaluator(name, n_k)
    lib = ev.lib
    assert lib.awegpu_casadi_bind(ev.handle) == 0
    lib.awegpu_hess_lag_n_in.restype = C.c_longlong
    lib.awegpu_hess_lag_n_out.restype = C.c_longlong
    assert lib.awegpu_hess_lag_n_in() == 4 and lib.awegpu_hess_lag_n_out() == 1
    lib.awegpu_hess_lag_sparsity_out.restype = C.POINTER(C.c_longlong)
    lib.awegpu_hess_lag_sparsity_out.argtypes = [C.c_longlong]
    sp = lib.awegpu_hess_lag_sparsity_out(0)
    n_v = ev.n_v
    assert sp[0] == n_v and sp[1] == n_v
    colind = np.array([sp[2 + i] for i in range(n_v + 1)], dtype=np.int64)
    nnz = int(colind[-1])
    row = np.array([sp[2 + n_v + 1 + i] for i in range(nnz)], dtype=np.int64)
    assert np.array_equal(colind, gold['colind']) and np.array_equal(row, gold['row'])
    lib.awegpu_hess_lag_checkout.restype = C.c_int
    mem = lib.awegpu_hess_lag_checkout()
    w, p, lam = (np.ascontiguousarray(a, dtype=np.float64) for a in (gold['w'][1], gold['p'], gold['lam'][1]))
    lam_f = np.array([gold['lam_f'][1]])
    out = np.full(nnz, np.nan)
    arg = (C.POINTER(C.c_double) * 4)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in (w, p, lam_f, lam)])
    res = (C.POINTER(C.c_double) * 1)(out.ctypes.data_as(C.POINTER(C.c_double)))
    lib.awegpu_hess_lag.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    assert lib.awegpu_hess_lag(arg, res, None, None, mem) == 0
    assert np.all(np.abs(out - gold['hv'][1]) <= hess_tolerance(hessian(name, n_k, objective=True), gold['hv'][1]))
