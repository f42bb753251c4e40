import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def ap2_spec():
    from awebox_b200.spec import ModelSpec
    return ModelSpec.load('ampyx_ap2')


_OCP_CACHE = {}


def get_ocp(name, n_k):
    from awebox_b200.spec import ModelSpec
    from awebox_b200.ocp.nlp import OcpProblem
    key = (name, n_k)
    if key not in _OCP_CACHE:
        _OCP_CACHE[key] = OcpProblem(ModelSpec.load(name), n_k=n_k)
    return _OCP_CACHE[key]


_ORACLE_CACHE = {}


def get_oracle(name, n_k):
    """the plain-C checker (oracle/awebox_oracle.c) for one problem; builds it on first use."""
    import build_oracle
    from awebox_b200 import build as pbuild
    key = (name, n_k)
    if key not in _ORACLE_CACHE:
        if not os.path.exists(os.path.join(pbuild.gen_dir(name), 'node_coll.inc')):
            pbuild.generate(name)
        _ORACLE_CACHE[key] = build_oracle.Oracle(get_ocp(name, n_k))
    return _ORACLE_CACHE[key]


def jac_tolerance(ocp, ref_jv, rtol=1e-12):
    """|a-b| <= rtol * max(|ref|, largest |entry| of the same row): relative to the entry, with a
    floor at the row scale so that entries that are pure cancellation noise (e.g. d g/d t_f at a
    kinematically consistent point) are judged against the terms they are the difference of."""
    import numpy as np
    ref_jv = np.atleast_2d(ref_jv)
    rowmax = np.zeros((ref_jv.shape[0], ocp.n_g))
    for b in range(ref_jv.shape[0]):
        np.maximum.at(rowmax[b], ocp.row, np.abs(ref_jv[b]))
    return rtol * np.maximum(np.abs(ref_jv), rowmax[:, ocp.row])
