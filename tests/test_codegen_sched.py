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
    lib.sched_coll.argtypes = [PD] * 7 + [C.c_double, C.c_int, PD, C.c_int]
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
    ocp = get_ocp('ampyx_ap2', 4)
    xr, u, z, th, p = _inputs(ocp, 3 + j)
    tf = th[ocp.tf_local - ocp.rng['theta'][0]]
    cc = ocp.coll.coeff_collocation[:, j + 1] * ocp.n_k / tf
    own = np.zeros(ocp.d + 1)
    own[j + 1] = 1.0
    xo = np.ascontiguousarray(xr[j + 1])
    xdot = cc @ xr
    v = np.concatenate([xo, xdot, u, z, th])
    n_out = len(ocp.coll_out)
    a, b, c = np.zeros(n_out), np.zeros(n_out), np.zeros(n_out)
    harness.plain_coll(_ptr(v), _ptr(p), _ptr(cc), _ptr(own), 1.0 / tf, _ptr(a))
    xdot = np.ascontiguousarray(xdot)
    harness.sched_coll(_ptr(xo), _ptr(xdot), _ptr(u), _ptr(z), _ptr(th), _ptr(p), _ptr(cc), 1.0 / tf, j + 1, _ptr(b), 1)
    harness.sched_coll(_ptr(xo), _ptr(xdot), _ptr(u), _ptr(z), _ptr(th), _ptr(p), _ptr(cc), 1.0 / tf, j + 1, _ptr(c), 0)
    scale = np.maximum(1.0, np.abs(a))
    assert np.max(np.abs(a - b) / scale) < 1e-11
    assert np.max(np.abs(a[:ocp.n_eq] - c[:ocp.n_eq]) / scale[:ocp.n_eq]) < 1e-11


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
