"""The scheduled / register-allocated node programs (k_*.inc, what the CUDA kernels run) against the
plain emission (node_*.inc, what the CPU checker runs), compiled for the host from the same text."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, get_ocp

PD = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(PD)


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    from awebox_b200 import build as pbuild
    g = pbuild.gen_dir('ampyx_ap2')
    if not os.path.exists(os.path.join(g, 'k_coll.inc')):
        pbuild.generate('ampyx_ap2')
    out = str(tmp_path_factory.mktemp('harness') / 'libsched_harness.so')
    cmd = ['gcc', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-I', g, '-o', out,
           os.path.join(ROOT, 'oracle', 'sched_harness.c'), '-lm']
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(out)
    lib.sched_coll.argtypes = [PD] * 7 + [C.c_double, C.c_int, PD, C.c_int, PD, PD, PD, C.c_double, PD]
    lib.sched_shoot.argtypes = [PD] * 6 + [PD, C.c_int]
    lib.plain_coll.argtypes = [PD] * 4 + [C.c_double, PD]
    lib.plain_shoot.argtypes = [PD, PD, PD]
    return lib


def _inputs(ocp, seed):
    rng = np.random.default_rng(seed)
    d, n_x = ocp.d, ocp.n_x
    xr = rng.uniform(0.5, 1.5, (d + 1, n_x))
    u = rng.uniform(0.5, 1.5, ocp.n_u)
    z = rng.uniform(0.5, 1.5, ocp.n_z)
    th = rng.uniform(0.5, 1.5, ocp.n_theta)
    p = np.concatenate([ocp.spec.theta0_values, np.ones(ocp.n_phi)])
    return xr, u, z, th, p


@pytest.mark.parametrize('j', [0, 1, 3])
def test_scheduled_collocation_program_matches_plain_emission(harness, j):
    """The kernel programs of node j (head outputs + the node's own CCS range, which holds rows of all d nodes of the
    interval) against the plain emission evaluated at every node of the interval."""
    ocp = get_ocp('ampyx_ap2', 4)
    d = ocp.d
    xr, u, _, th, p = _inputs(ocp, 3 + j)
    zs_all = np.random.default_rng(17 + j).uniform(0.5, 1.5, (d, ocp.n_z))
    tf = th[ocp.tf_local - ocp.rng['theta'][0]]
    C = ocp.coll.coeff_collocation
    plain, xdots = [], []
    for jj in range(d):
        cc = C[:, jj + 1] * ocp.n_k / tf
        own = np.zeros(d + 1)
        own[jj + 1] = 1.0
        xdot = np.ascontiguousarray(cc @ xr)
        xdots.append(xdot)
        v = np.concatenate([xr[jj + 1], xdot, u, zs_all[jj], th])
        a = np.zeros(len(ocp.coll_out))
        harness.plain_coll(_ptr(v), _ptr(p), _ptr(np.ascontiguousarray(cc)), _ptr(own), 1.0 / tf, _ptr(a))
        plain.append(dict(zip(ocp.coll_out, a)))
    # expected values in kernel output order
    exp = np.zeros(len(ocp.kcoll_out))
    for n, item in enumerate(ocp.kcoll_out):
        if item[0] == 'O':
            _, i, mm = item
            exp[n] = plain[j][('Xc', i, mm, j + 1)] if mm in ocp.eq_split[i]['xdot'] else plain[j][('Xo', i, mm)]
        elif item[0] == 'S':
            _, dl, i, mm = item
            exp[n] = plain[(j + dl) % d][('Xc', i, mm, j + 1)]
        else:
            exp[n] = plain[j][item]
    sib = [(j + dl) % d for dl in range(1, d)]
    xs = np.ascontiguousarray(np.concatenate([xr[s_ + 1] for s_ in sib]))
    xds = np.ascontiguousarray(np.concatenate([xdots[s_] for s_ in sib]))
    zs = np.ascontiguousarray(np.concatenate([zs_all[s_] for s_ in sib]))
    cc = np.ascontiguousarray(C[:, j + 1] * ocp.n_k / tf)
    cco = float(C[j + 1, j + 1] * ocp.n_k / tf)
    ccs = np.ascontiguousarray([0.0] + [C[j + 1, s_ + 1] * ocp.n_k / tf for s_ in sib])
    xo, xd, z = np.ascontiguousarray(xr[j + 1]), xdots[j], np.ascontiguousarray(zs_all[j])
    b, c = np.zeros(len(exp)), np.zeros(len(exp))
    args = (_ptr(xo), _ptr(xd), _ptr(u), _ptr(z), _ptr(th), _ptr(p), _ptr(cc), 1.0 / tf, j + 1)
    harness.sched_coll(*args, _ptr(b), 1, _ptr(xs), _ptr(xds), _ptr(zs), cco, _ptr(ccs))
    harness.sched_coll(*args, _ptr(c), 0, _ptr(xs), _ptr(xds), _ptr(zs), cco, _ptr(ccs))
    scale = np.maximum(1.0, np.abs(exp))
    assert np.max(np.abs(exp - b) / scale) < 1e-11
    assert np.max(np.abs(exp[:ocp.n_eq] - c[:ocp.n_eq]) / scale[:ocp.n_eq]) < 1e-11


def test_scheduled_shooting_program_matches_plain_emission(harness):
    ocp = get_ocp('ampyx_ap2', 4)
    xr, u, z, th, p = _inputs(ocp, 11)
    xo, xd = np.ascontiguousarray(xr[0]), np.ascontiguousarray(xr[1])
    v = np.concatenate([xo, xd, u, z, th])
    n_out = len(ocp.shoot_out)
    a, b, c = np.zeros(n_out), np.zeros(n_out), np.zeros(n_out)
    harness.plain_shoot(_ptr(v), _ptr(p), _ptr(a))
    harness.sched_shoot(_ptr(xo), _ptr(xd), _ptr(u), _ptr(z), _ptr(th), _ptr(p), _ptr(b), 1)
    harness.sched_shoot(_ptr(xo), _ptr(xd), _ptr(u), _ptr(z), _ptr(th), _ptr(p), _ptr(c), 0)
    n_rows = len(ocp.shoot_rows)
    scale = np.maximum(1.0, np.abs(a))
    assert np.max(np.abs(a - b) / scale) < 1e-11
    assert np.max(np.abs(a[:n_rows] - c[:n_rows]) / scale[:n_rows]) < 1e-11
