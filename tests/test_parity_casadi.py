"""Parity against a LIVE casadi==3.6.4 evaluation of the reference.

The fixtures ``tests/golden_casadi/<config>_nk<N>.npz`` are produced on a machine that has the real package:

    AWEBOX_REAL_CASADI=1 AWEBOX_REFERENCE=/path/to/awebox python oracle/make_golden.py

(``cas.jacobian(g_fun(V, P), V)`` -> ``sparsity_out(0).colind() / row()`` and the non-zeros of a numeric call: what
``cas.nlpsol`` hands to IPOPT, awebox/opti/preparation.py:373-379).  CasADi cannot be installed in the build container
(no network, no wheel), so the files do not exist there and these tests are skipped with that reason; until they run,
parity is pinned against the reference's own Python executed over the CasADi restatement (tests/golden/), not against
CasADi's own instruction tape."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, get_ocp, get_oracle, jac_tolerance

GOLDEN_CASADI = os.path.join(ROOT, 'tests', 'golden_casadi')
FILES = sorted(glob.glob(os.path.join(GOLDEN_CASADI, '*_nk*.npz')))
REASON = ('no live-CasADi fixtures under tests/golden_casadi/: casadi==3.6.4 is not installable in this environment; '
          'generate them with AWEBOX_REAL_CASADI=1 python oracle/make_golden.py where the package exists')

if not FILES:
    pytest.skip(REASON, allow_module_level=True)


def _case(path):
    name, nk = os.path.basename(path)[:-4].rsplit('_nk', 1)
    return name, int(nk)


@pytest.mark.parametrize('path', FILES)
def test_pattern_bit_exact_against_live_casadi(path):
    name, n_k = _case(path)
    gold = np.load(path)
    assert 'shim' not in str(gold['casadi_version'])
    ocp = get_ocp(name, n_k)
    assert np.array_equal(ocp.colind, gold['colind']) and np.array_equal(ocp.row, gold['row'])


@pytest.mark.parametrize('path', FILES)
def test_checker_values_against_live_casadi(path):
    name, n_k = _case(path)
    gold = np.load(path)
    ocp = get_ocp(name, n_k)
    G, Jv = get_oracle(name, n_k).eval(gold['w'], gold['p'], flags=8)
    assert np.all(np.abs(G - gold['g']) <= 1e-12 * np.maximum(1.0, np.abs(gold['g'])))
    assert np.all(np.abs(Jv - gold['jv']) <= jac_tolerance(ocp, gold['jv']))


@pytest.mark.gpu
@pytest.mark.parametrize('path', FILES)
def test_cuda_values_against_live_casadi(path):
    from awebox_b200.evaluator import Evaluator
    name, n_k = _case(path)
    gold = np.load(path)
    ocp = get_ocp(name, n_k)
    G, Jv = Evaluator(name, ocp=ocp).eval(gold['w'], gold['p'])
    assert np.all(np.abs(G - gold['g']) <= 1e-12 * np.maximum(1.0, np.abs(gold['g'])))
    assert np.all(np.abs(Jv - gold['jv']) <= jac_tolerance(ocp, gold['jv']))
