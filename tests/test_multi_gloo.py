"""world_size-2 test of the batch sharding + result gather (gloo on CPU; the per-shard evaluation is
stood in for by the CPU checker, the sharding / gather code is the product's)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _inputs(spec, n_k, B):
    """-> (ocp, W [B, n_v], P, flags): shared parameter vector for the periodic OCP, per-instance scenario
    parameters (p.ref, wind speed) for the MPC trial."""
    from awebox_b200 import synthetic, cabi
    from conftest import get_ocp
    ocp = get_ocp(spec, n_k)
    W = synthetic.batch(ocp, B, seed=9)
    if ocp.needs_full_p:
        return ocp, W, synthetic.scenario_params(ocp, W, seed=10), cabi.AWEGPU_P_FULL
    return ocp, W, ocp.default_p(), cabi.AWEGPU_P_SHARED


def _worker(rank, world, port, B, out, spec='ampyx_ap2', n_k=4):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from awebox_b200.parallel import shard_range, gather_results, gather_to_root, checksum
    from conftest import get_oracle
    ocp, W, P, flags = _inputs(spec, n_k, B)
    lo, hi = shard_range(B, rank, world)
    G, Jv = get_oracle(spec, n_k).eval(W[lo:hi], P if P.ndim == 1 else P[lo:hi], flags=flags)
    Gall = gather_results(torch.from_numpy(G), B)
    Jall = gather_results(torch.from_numpy(Jv), B)
    cs = checksum(torch.from_numpy(G))
    Groot = gather_to_root(torch.from_numpy(G), B)                 # the north_star's "gather per-shard results"
    assert (Groot is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(Groot, Gall)
        np.savez(out, G=Gall.numpy(), J=Jall.numpy(), cs=cs)
    dist.barrier()
    dist.destroy_process_group()


def _run(B, tmp_path, spec='ampyx_ap2', n_k=4):
    out = str(tmp_path / ('gather_%s_%d.npz' % (spec, B)))
    port = 29500 + (os.getpid() + B) % 2000
    mp.spawn(_worker, args=(2, port, B, out, spec, n_k), nprocs=2, join=True)
    from conftest import get_oracle
    ocp, W, P, flags = _inputs(spec, n_k, B)
    G0, J0 = get_oracle(spec, n_k).eval(W, P, flags=flags)
    r = np.load(out)
    assert np.array_equal(r['G'], G0) and np.array_equal(r['J'], J0)
    assert abs(float(r['cs']) - G0.sum()) <= 1e-9 * np.abs(G0).sum()


def test_two_ranks_even_split(tmp_path):
    _run(10, tmp_path)


def test_two_ranks_ragged_split(tmp_path):
    _run(7, tmp_path)


def test_power_curve_parameters_are_sharded_consistently():
    """BASELINE configs[3]: instance i of the global batch gets wind speed i mod 64 whatever the sharding"""
    from awebox_b200 import synthetic
    from awebox_b200.parallel import shard_range
    from conftest import get_ocp
    ocp = get_ocp('dual_kites', 3)
    whole = synthetic.power_curve_params(ocp, 200)
    iu = ocp.spec.theta0_labels.index('[theta0,wind,u_ref,0]')
    assert len(np.unique(whole[:, iu])) == 64 and whole[:, iu].min() == 5.0 and whole[:, iu].max() == 8.0
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = shard_range(200, r, world)
            parts.append(synthetic.power_curve_params(ocp, hi - lo, first=lo))
        assert np.array_equal(np.concatenate(parts), whole)


def test_two_ranks_mpc_scenarios(tmp_path):
    """BASELINE configs[4]: scenario batch with per-instance parameter vectors, both W and P sharded."""
    _run(9, tmp_path, spec='mpc_3dof', n_k=3)


def _worker_outputs(rank, world, port, B, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from awebox_b200.parallel import shard_range, gather_results, gather_to_root
    from awebox_b200.ocp.outputs import OutputsMap
    from awebox_b200 import synthetic
    from conftest import get_ocp
    ocp = get_ocp('mpc_3dof', 3)
    om = OutputsMap(ocp)
    W = synthetic.batch(ocp, B, seed=9)
    lo, hi = shard_range(B, rank, world)
    local = np.array([om.evaluate(W[b], ocp.default_p()).T for b in range(lo, hi)])       # [b_r, n_out, n_nodes]
    both = gather_results(torch.from_numpy(local), B)
    root = gather_to_root(torch.from_numpy(local), B)
    assert (root is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(root, both)
        np.save(out, both.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_outputs_map_sharded_and_gathered(tmp_path):
    """the outputs map of a ragged batch sharded over two ranks (3-dimensional result blocks): gather on every rank and
    on the root alone (the host restatement stands in for the kernel)"""
    from awebox_b200.ocp.outputs import OutputsMap
    from awebox_b200 import synthetic
    from conftest import get_ocp
    B = 5
    out = str(tmp_path / 'gather_outputs.npy')
    port = 29500 + (os.getpid() + 977) % 2000
    mp.spawn(_worker_outputs, args=(2, port, B, out), nprocs=2, join=True)
    ocp = get_ocp('mpc_3dof', 3)
    om = OutputsMap(ocp)
    W = synthetic.batch(ocp, B, seed=9)
    ref = np.array([om.evaluate(W[b], ocp.default_p()).T for b in range(B)])
    assert np.array_equal(np.load(out), ref)
