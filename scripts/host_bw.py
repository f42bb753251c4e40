"""aggregate host <-> device copy bandwidth with one rank per GPU copying at the same time (pinned buffers): the
ceiling of the end-to-end (host-buffer) legs of bench.py at N ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/host_bw.py
"""
import json, os, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
n = 1 << 27                       # 1 GiB of float64
h_in, h_out = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.float64, device='cuda'), torch.ones(n, dtype=torch.float64, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, h2d, d2h in (('d2h', False, True), ('h2d', True, False), ('both', True, True)):
    for rep in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = 3 * 8 * n * (int(h2d) + int(d2h)) * world / float(t.item()) / 1e9
if rank == 0:
    print(json.dumps({'bench': 'host_bw', 'ranks': world, 'aggregate_GBps': res, 'per_rank_GBps': {k: v / world for k, v in res.items()}}))
if world > 1:
    dist.destroy_process_group()
