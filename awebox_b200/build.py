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
              '--shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-diag-suppress', '177', '--split-compile', '0']


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
    stats = NodeCode(ocp).write(gen_dir(spec_name))
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


def compile_cuda(spec_name, force=False, verbose=False, extra_flags=(), out=None):
    os.makedirs(LIBDIR, exist_ok=True)
    g = gen_dir(spec_name)
    srcs = [os.path.join(CSRC, 'awegpu_kernels.cu')]
    deps = srcs + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(g, f) for f in sorted(os.listdir(g)) if f.endswith(('.h', '.inc'))] + \
        [os.path.join(ROOT, 'include', 'awegpu.h')]
    out = out or lib_path(spec_name)
    stamp = out + '.sha256'
    digest = _digest(deps) + ''.join(extra_flags)
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == digest:
        return out
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + ['-I', g, '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
                                                             '-o', out] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(g, 'nvcc.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log[-4000:])
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
