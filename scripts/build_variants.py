"""Kernel-tuning experiments: build several libraries of one spec with different generator parameters.

    python scripts/build_variants.py name:KEY=VAL,KEY=VAL[,-DFLAG] ...

Generator keys are module attributes of awebox_b200.codegen (POOL, PF_AHEAD, N_REGS, CHUNK, GROUP_OPS ...);
entries starting with -D go to nvcc.  Output: awebox_b200/lib/var/libawegpu_<spec>.<name>.so, selected at run
time with AWEGPU_LIB=<path>.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from awebox_b200 import build as pbuild, codegen as CG
from awebox_b200.spec import ModelSpec
from awebox_b200.ocp.nlp import OcpProblem

SPEC = os.environ.get('AWE_SPEC', 'ampyx_ap2')


def main(argv):
    ocp = OcpProblem(ModelSpec.load(SPEC))
    jobs = []
    defaults = {k: getattr(CG, k) for k in dir(CG) if k.isupper()}
    for arg in argv:
        name, _, rest = arg.partition(':')
        flags = []
        for k, v in defaults.items():
            setattr(CG, k, v)
        os.environ.pop('AWE_ILP_WINDOW', None), os.environ.pop('AWE_ILP_LATENCY', None)
        for kv in filter(None, rest.split(',')):
            if kv.startswith('-D'):
                flags.append(kv)
            elif kv.startswith('ILP_'):                       # awebox_b200.schedule: ILP_WINDOW / ILP_LATENCY
                k, v = kv.split('=')
                os.environ['AWE_' + k] = v
            else:
                k, v = kv.split('=')
                setattr(CG, k, type(defaults[k])(v))
        CG.NodeCode.emit_scheduled.__defaults__ = (CG.N_REGS, CG.LOOKAHEAD, CG.HOT_SLOTS, CG.CHUNK, 0, False, None, 0, None, None)
        g = os.path.join(pbuild.CSRC, 'gen', '_var_' + name)
        stats = CG.NodeCode(ocp).write(g)
        import shutil
        for f in ('node_hess_coll.inc', 'node_hess_shoot.inc', 'hess_defs.h', 'node_out.inc', 'out_defs.h'):
            shutil.copy(os.path.join(pbuild.gen_dir(SPEC), f), os.path.join(g, f))     # units the experiment does not touch
        out = os.path.join(pbuild.LIBDIR, 'var', 'libawegpu_%s.%s.so' % (SPEC, name))
        os.makedirs(os.path.dirname(out), exist_ok=True)
        jobs.append((name, flags, g, out))
        print(name, {k: stats[k]['hot_slots'] for k in ('k_coll', 'k_shoot')}, flush=True)

    def run(job):
        name, flags, g, out = job
        try:
            pbuild.compile_cuda(SPEC, force=True, extra_flags=tuple(flags), out=out, gen=g)
        except RuntimeError as e:
            return name, 1, -1, str(e)[-2000:]
        log = open(os.path.join(g, 'nvcc.log')).read()
        regs = [l for l in log.splitlines() if 'Used' in l]
        return name, 0, max((int(l.split('Used')[1].split()[0]) for l in regs), default=-1), ''

    with ThreadPoolExecutor(2) as ex:
        for name, rc, regs, err in ex.map(run, jobs):
            print('%s: rc=%d max regs %d %s' % (name, rc, regs, err), flush=True)


if __name__ == '__main__':
    main(sys.argv[1:])
