"""Batch sharding over the GPUs of one box.

OCP instances of a batch are independent (same model, same pattern; different ``w_b``, ``p_b``), so the
path shards with NO data-path collective: rank r evaluates the contiguous block
``shard_range(B, r, world)``.  A collective is used only when the caller wants all results in one
place: ``gather_results`` is one ``all_gather`` of the per-rank blocks (NCCL over NVLink on GPUs,
gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """contiguous block of instances owned by ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(B), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(B, world):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_results(local, B, group=None):
    """all-gather per-rank result blocks ``local [b_r, n]`` into ``[B, n]`` on every rank (trailing dimensions of any
    rank, e.g. the outputs map ``[b_r, n_out, n_nodes]``, are carried along)."""
    if local.dim() != 2:
        tail = tuple(local.shape[1:])
        return gather_results(local.reshape(local.shape[0], -1), B, group).reshape((B,) + tail)
    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    n = local.shape[1]
    if len(set(sizes)) == 1:
        out = torch.empty((B, n), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, n), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def gather_to_root(local, B, root=0, group=None):
    """gather the per-rank blocks ``local [b_r, n]`` on ``root`` only -> ``[B, n]`` there, None elsewhere (the
    "NCCL over NVLink only to gather per-shard results" step of BASELINE.json: north_star)."""
    if local.dim() != 2:
        tail = tuple(local.shape[1:])
        out = gather_to_root(local.reshape(local.shape[0], -1), B, root, group)
        return None if out is None else out.reshape((B,) + tail)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(B, world)
    pad, n = max(sizes), local.shape[1]
    buf = local.contiguous()
    if local.shape[0] != pad:
        buf = torch.zeros((pad, n), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == root else None
    dist.gather(buf, parts, dst=root, group=group)
    if rank != root:
        return None
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def checksum(local, group=None):
    """order-independent checksum of the whole batch result (sum over ranks of the local sums)."""
    s = local.sum(dtype=torch.float64).reshape(1).clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return float(s.item())
