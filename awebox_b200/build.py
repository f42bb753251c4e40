"""Generate the node functions of a ModelSpec and compile the CUDA library for sm_100a.

    python -m awebox_b200.build [spec ...]

Outputs (in-tree, so they travel to the GPU box with the repository snapshot):
    awebox_b200/csrc/gen/<spec>/{model_defs.h,node_coll.inc,node_shoot.inc}
    awebox_b200/lib/libawegpu_<spec>.so
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-diag-suppress', '177']
N_TU = 5        # translation units of awegpu_kernels.cu (-DAWE_TU=0..4), compiled side by side and linked


def gen_dir(spec_name):
    return os.path.join(CSRC, 'gen', spec_name)


def lib_path(spec_name):
    return os.path.join(LIBDIR, 'libawegpu_%s.so' % spec_name)


def generate(spec_name, ocp=None):
    from .spec import ModelSpec
    from .ocp.nlp import OcpProblem
    from .codegen import NodeCode
    if ocp is None:
        ocp = OcpProblem(ModelSpec.load(spec_name))
    nc = NodeCode(ocp)
    stats = nc.write(gen_dir(spec_name))
    from .ocp.outputs import OutputsMap
    stats['outputs'] = OutputsMap(ocp).write(gen_dir(spec_name), nc)
    from .ocp.hessian import LagrangianHessian, has_objective
    stats['hessian'] = LagrangianHessian(ocp, objective=has_objective(ocp)).write(gen_dir(spec_name))
    with open(os.path.join(gen_dir(spec_name), 'stats.json'), 'w') as f:
        json.dump(stats, f, indent=1, sort_keys=True)
    return ocp, stats


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: the CUDA library cannot be built')
    return nvcc


def compile_cuda(spec_name, force=False, verbose=False, extra_flags=(), out=None, gen=None):
    """nvcc: one object per translation unit (-DAWE_TU=n, in parallel: the front end is single-threaded and
    the generated programs are large), then one shared library.  Deterministic: no --split-compile."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    g = gen or gen_dir(spec_name)
    src = os.path.join(CSRC, 'awegpu_kernels.cu')
    deps = [src] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", "_reduce.cu", "_hess.cu", "_out.cu"))] + \
        [os.path.join(g, f) for f in sorted(os.listdir(g)) if f.endswith(('.h', '.inc'))] + \
        [os.path.join(ROOT, 'include', 'awegpu.h')]
    out = out or lib_path(spec_name)
    stamp = out + '.sha256'
    digest = _digest(deps) + ''.join(extra_flags)
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == digest:
        return out
    nvcc = find_nvcc()
    objdir = os.path.join(g, '_obj')
    os.makedirs(objdir, exist_ok=True)
    common = NVCC_FLAGS + list(extra_flags) + ['-I', g, '-I', os.path.join(ROOT, 'include'), '-I', CSRC]

    def one(tu):
        obj = os.path.join(objdir, 'tu%d.o' % tu)
        cmd = [nvcc] + common + ['-DAWE_TU=%d' % tu, '-c', '-o', obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return tu, obj, cmd, res

    logs, objs = [], []
    def reduce_unit(_):
        # model-independent kernels (awegpu_reduce.cu)
        obj = os.path.join(objdir, 'reduce.o')
        cmd = [nvcc] + common + ['-c', '-o', obj, os.path.join(CSRC, 'awegpu_reduce.cu')]
        return 'reduce', obj, cmd, subprocess.run(cmd, capture_output=True, text=True)

    def hess_unit(_):
        # second-order hook (awegpu_hess.cu + generated node_hess_*.inc)
        obj = os.path.join(objdir, 'hess.o')
        cmd = [nvcc] + common + ['-c', '-o', obj, os.path.join(CSRC, 'awegpu_hess.cu')]
        return 'hess', obj, cmd, subprocess.run(cmd, capture_output=True, text=True)

    def out_unit(_):
        # outputs map (awegpu_out.cu + generated node_out.inc)
        obj = os.path.join(objdir, 'out.o')
        cmd = [nvcc] + common + ['-c', '-o', obj, os.path.join(CSRC, 'awegpu_out.cu')]
        return 'out', obj, cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(N_TU + 3) as pool:
        futs = [pool.submit(one, tu) for tu in range(N_TU)] + [pool.submit(reduce_unit, 0), pool.submit(hess_unit, 0),
                                                                pool.submit(out_unit, 0)]
        results = [f.result() for f in futs]
    for tu, obj, cmd, res in results:
        logs.append(' '.join(cmd) + '\n' + res.stdout + res.stderr)
        objs.append(obj)
    failed = [r for r in results if r[3].returncode != 0]
    if not failed:
        cmd = [nvcc, '--shared', '-o', out] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(' '.join(cmd) + '\n' + res.stdout + res.stderr)
        if res.returncode != 0:
            failed = [(-1, None, cmd, res)]
    log = '\n'.join(logs)
    with open(os.path.join(g, 'nvcc.log'), 'w') as f:
        f.write(log)
    shutil.rmtree(objdir, ignore_errors=True)
    if failed:
        raise RuntimeError('nvcc failed:\n' + (failed[0][3].stdout + failed[0][3].stderr)[-4000:])
    if verbose:
        print(log)
    with open(stamp, 'w') as f:
        f.write(digest)
    return out


def build(spec_name, force=False, verbose=False):
    generate(spec_name)
    return compile_cuda(spec_name, force=force, verbose=verbose)


if __name__ == '__main__':
    from .spec import available_specs
    names = sys.argv[1:] or available_specs()
    for n in names:
        print(n, '->', build(n, verbose=True))
