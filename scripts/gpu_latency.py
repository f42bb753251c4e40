"""single-instance latency through the reference-style entry points (what one IPOPT callback costs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from awebox_b200.evaluator import Evaluator
from awebox_b200 import synthetic
for spec, n_k in (('ampyx_ap2', 40), ('mpc_3dof', 10)):
    ev = Evaluator(spec, n_k=n_k)
    w = synthetic.batch(ev.ocp, 1, seed=1)[0]
    p = ev.ocp.default_p()
    for name, f in (('g_fun', ev.g_fun), ('jac_g_fun', ev.jac_g_fun)):
        for _ in range(20):
            f(w, p)
        t0 = time.perf_counter()
        n = 200
        for _ in range(n):
            f(w, p)
        dt = (time.perf_counter() - t0) / n
        print('%-10s n_k=%d  %-10s B=1: %.1f us per call (host numpy in / out, device time %.3f ms)' % (spec, n_k, name, dt * 1e6, ev.last_eval_ms()))
