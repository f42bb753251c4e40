"""The C-ABI library loads and exports every symbol include/awegpu.h declares; without a CUDA device
every entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, get_ocp


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'awegpu.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'\b(awegpu_\w+|jac_awegpu_\w+)\s*\(', text)
    return sorted(set(n for n in names if not n.endswith('_t')))


@pytest.fixture(scope='module')
def lib():
    from awebox_b200 import build as pbuild
    pbuild.build('ampyx_ap2')
    return C.CDLL(pbuild.lib_path('ampyx_ap2'))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ('awegpu_create', 'awegpu_eval_batch', 'awegpu_eval_batch_host', 'awegpu_sparsity', 'awegpu_g',
                 'jac_awegpu_g', 'awegpu_g_sparsity_in', 'jac_awegpu_g_sparsity_out', 'awegpu_destroy'):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), 'symbol %s missing from the shared library' % name


def test_model_identity(lib):
    lib.awegpu_model_name.restype = C.c_char_p
    assert lib.awegpu_model_name() == b'ampyx_ap2'
    nv, npar = C.c_int32(), C.c_int32()
    lib.awegpu_model_dims(C.byref(nv), C.byref(npar), None, None, None, None)
    assert (nv.value, npar.value) == (59, 128)


def test_struct_mirror_matches_header_field_order():
    """ctypes ProblemStruct must list the fields of awegpu_problem_t in the header's order."""
    from awebox_b200 import cabi
    text = open(os.path.join(ROOT, 'include', 'awegpu.h')).read()
    body = text[text.index('typedef struct awegpu_problem {'):text.index('} awegpu_problem_t;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        decl = decl.split('{')[-1]
        for part in decl.split(','):
            fields.append(re.findall(r'(\w+)\s*$', part.strip())[0])
    assert fields == [f for f, _ in cabi.ProblemStruct._fields_]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from awebox_b200.evaluator import Evaluator, AwegpuError
    with pytest.raises(AwegpuError, match='no CPU fallback'):
        Evaluator('ampyx_ap2', n_k=4)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'awebox_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'build_oracle' not in text
                assert 'awebox_oracle' not in text.replace('oracle/awebox_oracle.c', '')
