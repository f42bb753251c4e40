"""Experiments with the node Hessian programs: libraries of one spec with different section counts / nvcc flags.

    python scripts/build_hess_variants.py name:sections[,sync=N][,-DFLAG...] ...

Output: awebox_b200/lib/var/libawegpu_<spec>.<name>.so (select with AWEGPU_LIB=<path>)."""
import os
import shutil
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from awebox_b200 import build as pbuild
from awebox_b200.ocp import hessian as HS
from awebox_b200.spec import ModelSpec
from awebox_b200.ocp.nlp import OcpProblem

SPEC = os.environ.get('AWE_SPEC', 'ampyx_ap2')


def main(argv):
    ocp = OcpProblem(ModelSpec.load(SPEC))
    H = HS.LagrangianHessian(ocp, objective=HS.has_objective(ocp))
    jobs = []
    for arg in argv:
        name, _, rest = arg.partition(':')
        parts = rest.split(',')
        HS.HESS_SECTIONS = int(parts[0])
        flags = [f for f in parts[1:] if f and not f.startswith(('sync=', 'outsec='))]
        outsec = ([f[7:] for f in parts[1:] if f.startswith('outsec=')] or [None])[0]
        os.environ['AWE_HESS_SYNC_EVERY'] = ([f[5:] for f in parts[1:] if f.startswith('sync=')] or ['0'])[0]
        g = os.path.join(pbuild.CSRC, 'gen', '_var_' + name)
        shutil.rmtree(g, ignore_errors=True)
        shutil.copytree(pbuild.gen_dir(SPEC), g, ignore=shutil.ignore_patterns('_obj'))
        print(name, H.write(g), flush=True)
        if outsec is not None:          # outputs map cut into sections of this many operations
            from awebox_b200.ocp.outputs import OutputsMap
            from awebox_b200.codegen import NodeCode
            os.environ['AWE_OUT_SECTION_OPS'] = outsec
            print(name, 'outputs', OutputsMap(ocp).write(g, NodeCode(ocp)), flush=True)
            del os.environ['AWE_OUT_SECTION_OPS']
        out = os.path.join(pbuild.LIBDIR, 'var', 'libawegpu_%s.%s.so' % (SPEC, name))
        os.makedirs(os.path.dirname(out), exist_ok=True)
        jobs.append((name, flags, g, out))

    def run(job):
        name, flags, g, out = job
        try:
            pbuild.compile_cuda(SPEC, force=True, extra_flags=tuple(flags), out=out, gen=g)
        except RuntimeError as e:
            return name, str(e)[-1500:]
        log = open(os.path.join(g, 'nvcc.log')).read().splitlines()
        i = [n for n, l in enumerate(log) if 'awegpu_hess_node_kernel' in l and 'Compiling' in l]
        return name, ' | '.join(l.strip() for l in log[i[0] + 1:i[0] + 4]) if i else '?'

    with ThreadPoolExecutor(3) as ex:
        for name, info in ex.map(run, jobs):
            print(name, info, flush=True)


if __name__ == '__main__':
    main(sys.argv[1:])
