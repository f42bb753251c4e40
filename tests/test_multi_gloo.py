"""world_size-2 test of the batch sharding + result gather (gloo on CPU; the per-shard evaluation is
stood in for by the CPU checker, the sharding / gather code is the product's)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, B, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from awebox_b200.parallel import shard_range, gather_results, checksum
    from awebox_b200 import synthetic, cabi
    from conftest import get_ocp, get_oracle
    ocp = get_ocp('ampyx_ap2', 4)
    W = synthetic.batch(ocp, B, seed=9)
    lo, hi = shard_range(B, rank, world)
    G, Jv = get_oracle('ampyx_ap2', 4).eval(W[lo:hi], ocp.default_p(), flags=cabi.AWEGPU_P_SHARED)
    Gall = gather_results(torch.from_numpy(G), B)
    Jall = gather_results(torch.from_numpy(Jv), B)
    cs = checksum(torch.from_numpy(G))
    if rank == 0:
        np.savez(out, G=Gall.numpy(), J=Jall.numpy(), cs=cs)
    dist.barrier()
    dist.destroy_process_group()


def _run(B, tmp_path):
    out = str(tmp_path / ('gather_%d.npz' % B))
    port = 29500 + (os.getpid() + B) % 2000
    mp.spawn(_worker, args=(2, port, B, out), nprocs=2, join=True)
    from awebox_b200 import synthetic, cabi
    from conftest import get_ocp, get_oracle
    ocp = get_ocp('ampyx_ap2', 4)
    W = synthetic.batch(ocp, B, seed=9)
    G0, J0 = get_oracle('ampyx_ap2', 4).eval(W, ocp.default_p(), flags=cabi.AWEGPU_P_SHARED)
    r = np.load(out)
    assert np.array_equal(r['G'], G0) and np.array_equal(r['J'], J0)
    assert abs(float(r['cs']) - G0.sum()) <= 1e-9 * np.abs(G0).sum()


def test_two_ranks_even_split(tmp_path):
    _run(10, tmp_path)


def test_two_ranks_ragged_split(tmp_path):
    _run(7, tmp_path)
