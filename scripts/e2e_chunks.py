"""end-to-end throughput of the two host entries for different pipeline chunk sizes (AWEGPU_COPY_CHUNK is read once
per process: run once per value)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from awebox_b200.evaluator import Evaluator
from awebox_b200 import synthetic, cabi
B = 4096
ev = Evaluator('ampyx_ap2', n_k=40); ocp = ev.ocp
W = synthetic.batch(ocp, B, seed=3); p = ocp.default_p()
hW, hP = torch.from_numpy(W).pin_memory(), torch.from_numpy(p).pin_memory()
hG = torch.empty((B, ocp.n_g), dtype=torch.float64).pin_memory()
hJ = torch.empty((B, ocp.nnz), dtype=torch.float64).pin_memory()
hL = torch.from_numpy(np.random.default_rng(1).standard_normal(ocp.n_g)).pin_memory()
hT = torch.empty((B, ocp.n_v), dtype=torch.float64).pin_memory()
hI = torch.empty((B,), dtype=torch.float64).pin_memory()
fl = cabi.AWEGPU_P_SHARED | cabi.AWEGPU_HOST_PINNED
def full(): assert ev.lib.awegpu_eval_batch_host(ev.handle, B, hW.data_ptr(), hP.data_ptr(), hG.data_ptr(), hJ.data_ptr(), fl) == 0
def red(): assert ev.lib.awegpu_eval_batch_reduce_host(ev.handle, B, hW.data_ptr(), hP.data_ptr(), hL.data_ptr(), None, fl, 1, hG.data_ptr(), hT.data_ptr(), None, None, None, hI.data_ptr()) == 0
def red_nog(): assert ev.lib.awegpu_eval_batch_reduce_host(ev.handle, B, hW.data_ptr(), hP.data_ptr(), hL.data_ptr(), None, fl, 1, None, hT.data_ptr(), None, None, None, hI.data_ptr()) == 0
for name, fn in (('full', full), ('reduced', red), ('reduced, no g', red_nog)):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(4): fn()
    print('chunk %s  %-14s %.2f ms per 4096 instances' % (os.environ.get('AWEGPU_COPY_CHUNK', '512'), name, (time.perf_counter() - t0) / 4 * 1e3))
