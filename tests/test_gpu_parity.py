"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU checker on the same
seeded inputs, against the golden fixtures, and — at BASELINE.json's full size — through
size-independent properties.  Integer structure bit-exact; FP64 values within 1e-12 (entry-relative,
row-scale floor: see conftest.jac_tolerance)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, get_ocp, get_oracle, jac_tolerance
from awebox_b200 import cabi, synthetic

pytestmark = pytest.mark.gpu

_EV = {}


def evaluator(name, n_k):
    from awebox_b200.evaluator import Evaluator
    if (name, n_k) not in _EV:
        _EV[(name, n_k)] = Evaluator(name, ocp=get_ocp(name, n_k))
    return _EV[(name, n_k)]


def assert_parity(ocp, G, Jv, G0, J0):
    assert np.all(np.abs(G - G0) <= 1e-12 * np.maximum(1.0, np.abs(G0)))
    assert np.all(np.abs(Jv - J0) <= jac_tolerance(ocp, J0))


def test_native_library_is_loaded():
    ev = evaluator('ampyx_ap2', 4)
    assert ev.lib.awegpu_model_name() == b'ampyx_ap2'
    with open('/proc/self/maps') as f:
        assert 'libawegpu_ampyx_ap2.so' in f.read()


CASES = [('ampyx_ap2', 4), ('ampyx_ap2', 40), ('dual_kites', 3), ('dual_kites', 20), ('mpc_3dof', 3), ('mpc_3dof', 10)]


@pytest.mark.parametrize('name,n_k', CASES)
def test_sparsity_bit_exact(name, n_k):
    ev = evaluator(name, n_k)
    colind, row = ev.sparsity()
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d.npz' % (name, n_k)))
    assert colind.dtype == np.int64 and row.dtype == np.int64
    assert np.array_equal(colind, gold['colind']) and np.array_equal(row, gold['row'])


@pytest.mark.parametrize('name,n_k', CASES)
def test_against_reference_golden_vectors(name, n_k):
    ev = evaluator(name, n_k)
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d.npz' % (name, n_k)))
    G, Jv = ev.eval(gold['w'], gold['p'])
    assert_parity(ev.ocp, G, Jv, gold['g'], gold['jv'])


@pytest.mark.parametrize('name,n_k,B', [('ampyx_ap2', 4, 257), ('ampyx_ap2', 40, 96), ('dual_kites', 3, 97),
                                        ('dual_kites', 20, 40), ('mpc_3dof', 10, 130)])
def test_against_checker_on_seeded_batch(name, n_k, B):
    ev = evaluator(name, n_k)
    ocp = ev.ocp
    W = synthetic.batch(ocp, B, seed=20261018 + n_k)
    p = ocp.default_p()
    G0, J0 = get_oracle(name, n_k).eval(W, p, flags=cabi.AWEGPU_P_SHARED)
    G, Jv = ev.eval(W, p)
    assert_parity(ocp, G, Jv, G0, J0)


@pytest.mark.parametrize('B', [1, 31, 32, 33])
def test_ragged_tiles(B):
    ev = evaluator('ampyx_ap2', 4)
    W = synthetic.batch(ev.ocp, B, seed=5)
    p = ev.ocp.default_p()
    G0, J0 = get_oracle('ampyx_ap2', 4).eval(W, p, flags=cabi.AWEGPU_P_SHARED)
    G, Jv = ev.eval(W, p)
    assert G.shape == (B, ev.n_g) and Jv.shape == (B, ev.nnz)
    assert_parity(ev.ocp, G, Jv, G0, J0)


def test_dual_kite_power_curve_batch():
    """BASELINE configs[3]: one OCP, per-instance wind speed (theta0 differs per instance), ragged batch."""
    ev = evaluator('dual_kites', 3)
    ocp = ev.ocp
    B = 64 + 5
    W = synthetic.batch(ocp, B, seed=3)
    P = np.tile(ocp.default_p(), (B, 1))
    iu = ocp.p_theta0 + ocp.spec.theta0_labels.index('[theta0,wind,u_ref,0]')
    P[:, iu] = np.resize(np.linspace(5.0, 8.0, 64), B)
    G0, J0 = get_oracle('dual_kites', 3).eval(W, P, flags=cabi.AWEGPU_P_FULL)
    G, Jv = ev.eval(W, P)
    assert_parity(ocp, G, Jv, G0, J0)
    # the t_f bound rows (single_reelout) are the last two rows of g and depend on the two t_f entries only
    o_tf = ocp.v_theta['t_f'][0]
    assert list(ocp.aff_row) == [ocp.n_g - 2, ocp.n_g - 1]
    n_ro = round(ocp.n_k * 0.7)
    period = (W[:, o_tf] * n_ro + W[:, o_tf + 1] * (ocp.n_k - n_ro)) / ocp.n_k
    lb, ub = ocp.spec.opt('system_bounds.theta.t_f')
    assert np.allclose(G[:, -2], (period - ub) / 0.7, rtol=1e-13) and np.allclose(G[:, -1], (lb - period) / 0.7, rtol=1e-13)


def test_mpc_scenario_batch():
    """BASELINE configs[4]: tracking-MPC NLP (3-DOF kite, N = 10), a batch of scenarios that differ in the primal
    point, the tracked reference / initial state (p.ref) and the wind speed: per-instance full parameter vectors."""
    from awebox_b200.evaluator import AwegpuError
    ev = evaluator('mpc_3dof', 10)
    ocp = ev.ocp
    B = 257
    W = synthetic.batch(ocp, B, seed=9)
    P = synthetic.scenario_params(ocp, W, seed=10)
    G0, J0 = get_oracle('mpc_3dof', 10).eval(W, P, flags=cabi.AWEGPU_P_FULL)
    G, Jv = ev.eval(W, P)
    assert_parity(ocp, G, Jv, G0, J0)
    idx = np.array(ocp.initial_idx)
    assert np.array_equal(G[:, :len(idx)], W[:, ocp.v_x[0] + idx] - P[:, ocp.v_x[0] + idx])
    G2, J2 = ev.eval(W, P, batch_minor=True)
    assert_parity(ocp, np.ascontiguousarray(G2.T), np.ascontiguousarray(J2.T), G0, J0)
    with pytest.raises(AwegpuError):
        ev.eval(W, P[:, ocp.p_theta0:])          # the initial rows need p.ref: theta0-only blocks are refused


def test_chunked_batches(monkeypatch):
    """batches larger than the side buffers are evaluated in chunks of tiles (forced small here)."""
    from awebox_b200.evaluator import Evaluator
    monkeypatch.setenv('AWEGPU_CHUNK_TILES', '2')
    ocp = get_ocp('ampyx_ap2', 4)
    ev = Evaluator('ampyx_ap2', ocp=ocp)
    try:
        B = 257                                            # 9 tiles -> 5 chunks, the last one ragged
        W = synthetic.batch(ocp, B, seed=21)
        p = ocp.default_p()
        G0, J0 = get_oracle('ampyx_ap2', 4).eval(W, p, flags=cabi.AWEGPU_P_SHARED)
        G, Jv = ev.eval(W, p)
        assert_parity(ocp, G, Jv, G0, J0)
        G2, _ = ev.eval(W, p, jac=False)
        assert np.all(np.abs(G2 - G0) <= 1e-12 * np.maximum(1.0, np.abs(G0)))
    finally:
        ev.close()


def test_pinned_host_pipeline_matches_staged_entry():
    """AWEGPU_HOST_PINNED (direct chunked copies on three streams, what bench.py's e2e leg times) against the
    staged host entry."""
    import torch
    ev = evaluator('ampyx_ap2', 4)
    ocp = ev.ocp
    B = 1200                                               # three 512-instance copy chunks, the last one ragged
    W = synthetic.batch(ocp, B, seed=22)
    p = ocp.default_p()
    G0, J0 = ev.eval(W, p)
    hW, hP = torch.from_numpy(W).pin_memory(), torch.from_numpy(p).pin_memory()
    hG = torch.empty((B, ocp.n_g), dtype=torch.float64).pin_memory()
    hJ = torch.empty((B, ocp.nnz), dtype=torch.float64).pin_memory()
    rc = ev.lib.awegpu_eval_batch_host(ev.handle, B, hW.data_ptr(), hP.data_ptr(), hG.data_ptr(), hJ.data_ptr(),
                                       cabi.AWEGPU_P_SHARED | cabi.AWEGPU_HOST_PINNED)
    assert rc == 0
    assert np.array_equal(hG.numpy(), G0) and np.array_equal(hJ.numpy(), J0)


def test_empty_batch_is_a_no_op():
    ev = evaluator('ampyx_ap2', 4)
    G, Jv = ev.eval(np.zeros((0, ev.n_v)), ev.ocp.default_p())
    assert G.shape == (0, ev.n_g) and Jv.shape == (0, ev.nnz)


def test_layouts_and_parameter_modes_agree():
    ev = evaluator('ampyx_ap2', 4)
    ocp = ev.ocp
    B = 70
    W = synthetic.batch(ocp, B, seed=11)
    rng = np.random.default_rng(2)
    P = np.tile(ocp.default_p(), (B, 1))
    iu = ocp.p_theta0 + ocp.spec.theta0_labels.index('[theta0,wind,u_ref,0]')
    P[:, iu] = rng.uniform(5.0, 8.0, B)                 # per-instance wind speed (power-curve batch)
    G0, J0 = get_oracle('ampyx_ap2', 4).eval(W, P, flags=cabi.AWEGPU_P_FULL)
    G1, J1 = ev.eval(W, P)
    assert_parity(ocp, G1, J1, G0, J0)
    G2, J2 = ev.eval(W, P[:, ocp.p_theta0:])             # theta0-only parameter block
    assert np.array_equal(G1, G2) and np.array_equal(J1, J2)
    G3, J3 = ev.eval(W, P, batch_minor=True)             # [n_g, B] / [nnz, B]
    assert_parity(ocp, np.ascontiguousarray(G3.T), np.ascontiguousarray(J3.T), G0, J0)
    G4, _ = ev.eval(W, P, jac=False)                     # residual-only kernels (their own schedule)
    assert np.all(np.abs(G4 - G0) <= 1e-12 * np.maximum(1.0, np.abs(G0)))
    G5, _ = ev.eval(W, P, jac=False, batch_minor=True)
    assert np.all(np.abs(G5.T - G0) <= 1e-12 * np.maximum(1.0, np.abs(G0)))


def test_reference_style_single_instance_calls():
    ev = evaluator('ampyx_ap2', 4)
    gold = np.load(os.path.join(GOLDEN, 'ampyx_ap2_nk4.npz'))
    g = ev.g_fun(gold['w'][0], gold['p'])
    assert g.shape == (ev.n_g,)
    assert np.all(np.abs(g - gold['g'][0]) <= 1e-12 * np.maximum(1.0, np.abs(gold['g'][0])))
    J = ev.jac_g_scipy(gold['w'][0], gold['p'])
    assert J.shape == (ev.n_g, ev.n_v) and J.nnz == ev.nnz


def test_full_size_properties():
    """BASELINE configs[1]: AP2 n_k=40, B=4096 — properties that do not need the checker at full size."""
    import torch
    ev = evaluator('ampyx_ap2', 40)
    ocp = ev.ocp
    B = 4096
    W = synthetic.batch(ocp, B, seed=20261018 + 2)
    p = ocp.default_p()
    dW, dP = torch.from_numpy(W).cuda(), torch.from_numpy(p).cuda()
    dG = torch.empty((B, ocp.n_g), dtype=torch.float64, device='cuda')
    dJ = torch.empty((B, ocp.nnz), dtype=torch.float64, device='cuda')
    dJ.fill_(float('nan'))
    ev.eval_device(dW, dP, dG, dJ)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(dG).all()) and bool(torch.isfinite(dJ).all())      # every slot was written
    # (1) a sample of instances against the checker
    idx = np.arange(0, B, 257)
    G0, J0 = get_oracle('ampyx_ap2', 40).eval(W[idx], p, flags=cabi.AWEGPU_P_SHARED)
    assert_parity(ocp, dG[idx].cpu().numpy(), dJ[idx].cpu().numpy(), G0, J0)
    # (2) instance independence: evaluating a permuted batch permutes the results bit for bit
    perm = torch.randperm(B, device='cuda', generator=torch.Generator(device='cuda').manual_seed(0))
    dG2, dJ2 = torch.empty_like(dG), torch.empty_like(dJ)
    ev.eval_device(dW[perm].contiguous(), dP, dG2, dJ2)
    torch.cuda.synchronize()
    assert torch.equal(dG2, dG[perm]) and torch.equal(dJ2, dJ[perm])
    # (3) directional-derivative check of the whole sparse Jacobian, all instances at once:
    #     g(w + e v) - g(w - e v) = 2 e J v + O(e^3)
    colind = torch.from_numpy(ocp.colind).cuda()
    col = torch.repeat_interleave(torch.arange(ocp.n_v, device='cuda'), colind[1:] - colind[:-1])
    row = torch.from_numpy(ocp.row).cuda()
    v = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, ocp.n_v))).cuda() * dW.abs()
    eps = 1e-6
    Gp, Gm = torch.empty_like(dG), torch.empty_like(dG)
    ev.eval_device((dW + eps * v).contiguous(), dP, Gp)
    ev.eval_device((dW - eps * v).contiguous(), dP, Gm)
    Jv_prod = torch.zeros_like(dG).index_add_(1, row, dJ * v[:, col])
    fd = (Gp - Gm) / (2 * eps)
    scale = torch.zeros_like(dG).index_add_(1, row, (dJ * v[:, col]).abs()) + 1e-3
    assert float(((fd - Jv_prod).abs() / scale).max()) < 1e-5
    # (4) continuity rows are linear: x_{k+1} - x_{k,d}  (Radau), checked exactly on the inputs
    k = 7
    r0 = ocp.row_cont[k]
    want = dW[:, ocp.v_x[k + 1]:ocp.v_x[k + 1] + ocp.n_x] - dW[:, ocp.v_cx[k][ocp.d - 1]:ocp.v_cx[k][ocp.d - 1] + ocp.n_x]
    assert float((dG[:, r0:r0 + ocp.n_x] - want).abs().max()) < 1e-14


@pytest.mark.parametrize('name,n_k,B', [('dual_kites', 20, 4096), ('mpc_3dof', 10, 8192)])
def test_other_baseline_configs_at_full_batch(name, n_k, B):
    """BASELINE configs[2] (dual kites, batch 4096) and configs[4] (8192 MPC scenarios): every output slot is
    written, a sample of instances agrees with the checker, and instances do not influence one another."""
    import torch
    ev = evaluator(name, n_k)
    ocp = ev.ocp
    W = synthetic.batch(ocp, B, seed=31)
    if ocp.needs_full_p:
        P = synthetic.scenario_params(ocp, W, seed=32)
        flags = cabi.AWEGPU_P_FULL
    else:
        P = ocp.default_p()
        flags = cabi.AWEGPU_P_SHARED
    dW, dP = torch.from_numpy(W).cuda(), torch.from_numpy(P).cuda()
    dG = torch.full((B, ocp.n_g), float('nan'), dtype=torch.float64, device='cuda')
    dJ = torch.full((B, ocp.nnz), float('nan'), dtype=torch.float64, device='cuda')
    ev.eval_device(dW, dP, dG, dJ)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(dG).all()) and bool(torch.isfinite(dJ).all())
    idx = np.arange(0, B, 509)
    G0, J0 = get_oracle(name, n_k).eval(W[idx], P if P.ndim == 1 else P[idx], flags=flags)
    assert_parity(ocp, dG[idx].cpu().numpy(), dJ[idx].cpu().numpy(), G0, J0)
    perm = torch.randperm(B, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))
    dG2, dJ2 = torch.empty_like(dG), torch.empty_like(dJ)
    ev.eval_device(dW[perm].contiguous(), dP if dP.ndim == 1 else dP[perm].contiguous(), dG2, dJ2)
    torch.cuda.synchronize()
    assert torch.equal(dG2, dG[perm]) and torch.equal(dJ2, dJ[perm])


def test_casadi_external_abi():
    """The symbols casadi.external('awegpu_g', lib) binds (casadi/core/external.cpp conventions): g(V, P) and the
    Jacobian function jac_awegpu_g(V, P, out_g) -> (jac_g_V [CCS], jac_g_P [empty]), called the way CasADi calls a
    generated function: arrays of input / output pointers, compact CCS sparsity descriptors."""
    import ctypes as C
    ev = evaluator('ampyx_ap2', 4)
    lib, ocp = ev.lib, ev.ocp
    gold = np.load(os.path.join(GOLDEN, 'ampyx_ap2_nk4.npz'))
    assert lib.awegpu_casadi_bind(ev.handle) == 0
    PD, PLL = C.POINTER(C.c_double), C.POINTER(C.c_longlong)
    for f in (lib.awegpu_g_n_in, lib.awegpu_g_n_out, lib.jac_awegpu_g_n_in, lib.jac_awegpu_g_n_out):
        f.restype = C.c_longlong
    assert (lib.awegpu_g_n_in(), lib.awegpu_g_n_out()) == (2, 1)
    assert (lib.jac_awegpu_g_n_in(), lib.jac_awegpu_g_n_out()) == (3, 2)
    for f in (lib.awegpu_g_name_in, lib.awegpu_g_name_out, lib.jac_awegpu_g_name_out):
        f.restype, f.argtypes = C.c_char_p, [C.c_longlong]
    assert (lib.awegpu_g_name_in(0), lib.awegpu_g_name_in(1), lib.awegpu_g_name_out(0)) == (b'V', b'P', b'g')
    assert (lib.jac_awegpu_g_name_out(0), lib.jac_awegpu_g_name_out(1)) == (b'jac_g_V', b'jac_g_P')
    for f in (lib.awegpu_g_sparsity_in, lib.awegpu_g_sparsity_out, lib.jac_awegpu_g_sparsity_out):
        f.restype, f.argtypes = PLL, [C.c_longlong]
    sp = lib.awegpu_g_sparsity_in(0)
    assert (sp[0], sp[1], sp[2]) == (ocp.n_v, 1, 1)                      # dense column: [nrow, ncol, 1]
    sp = lib.awegpu_g_sparsity_out(0)
    assert (sp[0], sp[1], sp[2]) == (ocp.n_g, 1, 1)
    sp = lib.jac_awegpu_g_sparsity_out(0)                                # [nrow, ncol, colind[ncol+1], row[nnz]]
    assert (sp[0], sp[1]) == (ocp.n_g, ocp.n_v)
    colind = np.ctypeslib.as_array(sp, shape=(2 + ocp.n_v + 1 + ocp.nnz,))
    assert np.array_equal(colind[2:2 + ocp.n_v + 1], gold['colind'])
    assert np.array_equal(colind[2 + ocp.n_v + 1:], gold['row'])
    sp = lib.jac_awegpu_g_sparsity_out(1)
    assert (sp[0], sp[1]) == (ocp.n_g, ocp.n_p) and sp[2 + ocp.n_p] == 0   # structurally empty
    # numeric calls: int f(const double** arg, double** res, casadi_int* iw, double* w, int mem)
    w = np.ascontiguousarray(gold['w'][1])
    p = np.ascontiguousarray(gold['p'])
    g = np.empty(ocp.n_g)
    arg = (PD * 3)(w.ctypes.data_as(PD), p.ctypes.data_as(PD), g.ctypes.data_as(PD))
    res = (PD * 2)(g.ctypes.data_as(PD), None)
    lib.awegpu_g.argtypes = [C.POINTER(PD), C.POINTER(PD), PLL, PD, C.c_int]
    lib.jac_awegpu_g.argtypes = lib.awegpu_g.argtypes
    assert lib.awegpu_g(arg, res, None, None, 0) == 0
    assert np.all(np.abs(g - gold['g'][1]) <= 1e-12 * np.maximum(1.0, np.abs(gold['g'][1])))
    jv = np.empty(ocp.nnz)
    res = (PD * 2)(jv.ctypes.data_as(PD), None)                          # jac_g_P not requested
    assert lib.jac_awegpu_g(arg, res, None, None, 0) == 0
    assert np.all(np.abs(jv - gold['jv'][1]) <= jac_tolerance(ocp, gold['jv'][1])[0])
    # CasADi passes NULL for an all-zero input: refused (g is undefined at t_f = 0), reported as an evaluation error
    arg0 = (PD * 3)(None, p.ctypes.data_as(PD), None)
    res = (PD * 2)(g.ctypes.data_as(PD), None)
    assert lib.awegpu_g(arg0, res, None, None, 0) != 0


def test_casadi_external_abi_two_problems_in_one_process():
    """two OCPs of one model (two handles of one library) behind the CasADi symbols: each Function addresses its own
    handle through the `mem` index its checkout returned (the hippo triple of awebox/opti/preparation.py:285-380 shares
    one g_fun; a sweep over n_k does not)"""
    import ctypes as C
    ev_a, ev_b = evaluator('ampyx_ap2', 4), evaluator('ampyx_ap2', 40)
    lib = ev_a.lib
    PD, PLL = C.POINTER(C.c_double), C.POINTER(C.c_longlong)
    lib.awegpu_g.argtypes = [C.POINTER(PD), C.POINTER(PD), PLL, PD, C.c_int]
    lib.awegpu_g_sparsity_out.restype, lib.awegpu_g_sparsity_out.argtypes = PLL, [C.c_longlong]
    mems = []
    for ev in (ev_a, ev_b):
        assert lib.awegpu_casadi_bind(ev.handle) == 0
        lib.awegpu_g_incref()
        mems.append(lib.awegpu_g_checkout())
        assert lib.awegpu_g_init_mem(mems[-1]) == 0
        assert lib.awegpu_g_sparsity_out(0)[0] == ev.ocp.n_g         # what a Function created now would read
    assert mems[0] != mems[1]
    for ev, mem in ((ev_b, mems[1]), (ev_a, mems[0]), (ev_b, mems[1])):
        gold = np.load(os.path.join(GOLDEN, 'ampyx_ap2_nk%d.npz' % ev.ocp.n_k))
        w, p = np.ascontiguousarray(gold['w'][1]), np.ascontiguousarray(gold['p'])
        g = np.empty(ev.ocp.n_g)
        arg = (PD * 2)(w.ctypes.data_as(PD), p.ctypes.data_as(PD))
        res = (PD * 1)(g.ctypes.data_as(PD))
        assert lib.awegpu_g(arg, res, None, None, mem) == 0
        assert np.all(np.abs(g - gold['g'][1]) <= 1e-12 * np.maximum(1.0, np.abs(gold['g'][1])))
        lib.awegpu_g_release(mem)
    lib.awegpu_g_decref()
    lib.awegpu_g_decref()


# ---------------------------------------------------------------------------------------------- device-resident products
def _csc(ocp, jv):
    import scipy.sparse as sp
    return sp.csc_matrix((jv, ocp.row, ocp.colind), shape=(ocp.n_g, ocp.n_v))


@pytest.mark.parametrize('name,n_k,B,shared', [('ampyx_ap2', 4, 37, False), ('ampyx_ap2', 40, 9, True),
                                               ('dual_kites', 3, 33, False), ('mpc_3dof', 10, 65, True)])
def test_reduced_products_against_the_checker(name, n_k, B, shared):
    """awegpu_eval_batch_reduce_host (g, J^T lambda, J v, row / column max-norms, max |g|; the Jacobian values stay
    on the device) against the same products formed on the host from the checker's Jacobian."""
    ev = evaluator(name, n_k)
    ocp = ev.ocp
    W = synthetic.batch(ocp, B, seed=77)
    p = ocp.default_p()
    if ocp.needs_full_p:
        P = np.tile(p, (B, 1))
        G0, J0 = get_oracle(name, n_k).eval(W, P)
    else:
        P = p
        G0, J0 = get_oracle(name, n_k).eval(W, p, flags=cabi.AWEGPU_P_SHARED)
    rng = np.random.default_rng(5)
    lam = rng.standard_normal(ocp.n_g) if shared else rng.standard_normal((B, ocp.n_g))
    v = rng.standard_normal(ocp.n_v) if shared else rng.standard_normal((B, ocp.n_v))
    out = ev.eval_reduce(W, P, lam=lam, v=v)
    assert np.all(np.abs(out['G'] - G0) <= 1e-12 * np.maximum(1.0, np.abs(G0)))
    for b in range(B):
        J = _csc(ocp, J0[b])
        A = abs(J)
        lb, vb = (lam if shared else lam[b]), (v if shared else v[b])
        # sums of products: tolerance relative to the sum of the magnitudes of the terms
        jtl, jtl_mag = J.T @ lb, A.T @ np.abs(lb)
        assert np.all(np.abs(out['JTl'][b] - jtl) <= 1e-12 * np.maximum(1.0, jtl_mag))
        jv, jv_mag = J @ vb, A @ np.abs(vb)
        assert np.all(np.abs(out['Jv'][b] - jv) <= 1e-12 * np.maximum(1.0, jv_mag))
        cn = np.asarray(A.max(axis=0).todense()).ravel()
        rn = np.asarray(A.max(axis=1).todense()).ravel()
        assert np.all(np.abs(out['cnorm'][b] - cn) <= 1e-12 * np.maximum(1.0, cn))
        assert np.all(np.abs(out['rnorm'][b] - rn) <= 1e-12 * np.maximum(1.0, rn))
        assert abs(out['ginf'][b] - np.max(np.abs(G0[b]))) <= 1e-12 * max(1.0, np.max(np.abs(G0[b])))


def test_reduced_products_device_entry_matches_host_entry():
    import torch
    ev = evaluator('ampyx_ap2', 40)
    ocp = ev.ocp
    B = 70
    W = synthetic.batch(ocp, B, seed=78)
    p = ocp.default_p()
    rng = np.random.default_rng(6)
    lam, v = rng.standard_normal((B, ocp.n_g)), rng.standard_normal(ocp.n_v)
    ref = ev.eval_reduce(W, p, lam=lam, v=v)
    dW, dP = torch.from_numpy(W).cuda(), torch.from_numpy(p).cuda()
    dG = torch.empty((B, ocp.n_g), dtype=torch.float64, device='cuda')
    dJ = torch.empty((B, ocp.nnz), dtype=torch.float64, device='cuda')
    ev.eval_device(dW, dP, dG, dJ)
    out = ev.reduce_device(dG, dJ, lam=torch.from_numpy(lam).cuda(), v=torch.from_numpy(v).cuda())
    torch.cuda.synchronize()
    for key in ('cnorm', 'rnorm', 'ginf'):
        assert np.array_equal(out[key].cpu().numpy(), ref[key]), key
    # sums are accumulated with reductions in L2: the order of the additions is not fixed
    for key in ('Jv', 'JTl'):
        assert np.all(np.abs(out[key].cpu().numpy() - ref[key]) <= 1e-12 * np.maximum(1.0, np.abs(ref[key]))), key


# ---------------------------------------------------------------------------------------------- degenerate inputs, error report
def _strict_report(ocp, G, Jv, G0, J0):
    """largest strictly entry-relative error |a - b| / |b| over the entries that are not below the row-scale floor
    (conftest.jac_tolerance: |b| >= 1e-6 of the largest entry of the row), and the number of entries below it"""
    rowmax = np.zeros((J0.shape[0], ocp.n_g))
    for b in range(J0.shape[0]):
        np.maximum.at(rowmax[b], ocp.row, np.abs(J0[b]))
    big = np.abs(J0) >= 1e-6 * rowmax[:, ocp.row]
    big &= J0 != 0.0
    rel_j = float(np.max(np.abs(Jv - J0)[big] / np.abs(J0)[big]))
    # rows of g are sums of O(1) terms: relative to the entry from 1e-3 upwards, absolute below
    gb = np.abs(G0) >= 1e-3
    rel_g = float(np.max(np.abs(G - G0)[gb] / np.abs(G0)[gb])) if gb.any() else 0.0
    abs_g = float(np.max(np.abs(G - G0)[~gb])) if (~gb).any() else 0.0
    return {'max_rel_g': rel_g, 'max_abs_g_below_1e-3': abs_g, 'max_rel_jac': rel_j, 'jac_entries': int(big.size),
            'jac_entries_below_row_floor': int((~big).sum())}


@pytest.mark.parametrize('name,n_k', [('ampyx_ap2', 40), ('dual_kites', 20), ('mpc_3dof', 10)])
def test_strict_relative_error_against_golden_vectors(name, n_k):
    """what 'within 1e-12' means entry by entry: the strictly relative error of every Jacobian entry that is not pure
    cancellation noise (>= 1e-6 of its row's largest entry) is below 1e-10; printed and written to
    gpurun_out/strict_rel_err.json for DESIGN.md"""
    import json
    ev = evaluator(name, n_k)
    gold = np.load(os.path.join(GOLDEN, '%s_nk%d.npz' % (name, n_k)))
    G, Jv = ev.eval(gold['w'], gold['p'])
    rep = _strict_report(ev.ocp, G, Jv, gold['g'], gold['jv'])
    rep['points'] = int(gold['w'].shape[0])
    print('strict relative error %s n_k=%d: %s' % (name, n_k, rep))
    out = os.path.join(os.path.dirname(GOLDEN), '..', 'gpurun_out')
    if os.path.isdir(out):
        path = os.path.join(out, 'strict_rel_err.json')
        allrep = json.load(open(path)) if os.path.exists(path) else {}
        allrep['%s_nk%d' % (name, n_k)] = rep
        with open(path, 'w') as f:
            json.dump(allrep, f, indent=1, sort_keys=True)
    assert rep['max_rel_jac'] < 1e-10 and rep['max_rel_g'] < 1e-10 and rep['max_abs_g_below_1e-3'] < 1e-12


@pytest.mark.parametrize('name,n_k', [('ampyx_ap2', 4), ('mpc_3dof', 3)])
def test_non_finite_and_degenerate_inputs_match_the_checker(name, n_k):
    """zero position vector (zero norms), t_f -> 0, NaN and Inf in w: the set of non-finite results is the checker's
    (IEEE arithmetic on the CPU), and every finite result agrees as usual"""
    ev = evaluator(name, n_k)
    ocp = ev.ocp
    W = synthetic.batch(ocp, 8, seed=31)
    p = ocp.default_p()
    P = np.tile(p, (W.shape[0], 1)) if ocp.needs_full_p else p
    oq, _ = ocp.model.offset[('x', 'q10')]
    W[1, ocp.v_cx[1][2] + oq:ocp.v_cx[1][2] + oq + 3] = 0.0            # a collocation node at the origin
    W[2, ocp.v_x[0] + oq:ocp.v_x[0] + oq + 3] = 0.0                    # a shooting node at the origin
    W[3, int(ocp.tf_idx[0])] = 0.0                                    # t_f = 0
    W[4, ocp.v_cx[0][0] + 1] = np.nan
    W[5, ocp.v_u[1]] = np.inf
    W[6, ocp.v_z[0]] = -np.inf                                        # (denormal arguments are read as zero on the GPU: not tested)
    W[7, ocp.v_cx[2][1]:ocp.v_cx[2][1] + ocp.n_x] = 0.0               # a whole state block zero
    flags = 0 if ocp.needs_full_p else cabi.AWEGPU_P_SHARED
    with np.errstate(all='ignore'):
        G0, J0 = get_oracle(name, n_k).eval(W, P, flags=flags)
    G, Jv = ev.eval(W, P)
    assert np.array_equal(np.isfinite(G), np.isfinite(G0))
    assert np.array_equal(np.isfinite(Jv), np.isfinite(J0))
    fg, fj = np.isfinite(G0), np.isfinite(J0)
    assert np.all(np.abs(G - G0)[fg] <= 1e-11 * np.maximum(1.0, np.abs(G0))[fg])
    # finite entries of otherwise degenerate rows are judged against the entry itself
    assert np.all(np.abs(Jv - J0)[fj] <= 1e-9 * np.maximum(1.0, np.abs(J0))[fj])
    assert fg.sum() > 0.5 * fg.size and (~fg).sum() > 0
