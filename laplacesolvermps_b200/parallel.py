"""Multi-GPU partitioning: instances are independent, so the only parallelism across GPUs is sharding the
instance list (one process per GPU) and ONE gather of the per-instance result records at the end
(SURVEY.md section 8e).  No collective runs inside a solve."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous block of instances owned by `rank`; block sizes differ by at most one."""
    base, extra = divmod(n_total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local, n_total, group=None):
    """all_gather of an [n_local, k] tensor of per-instance scalars into [n_total, k] on every rank
    (NCCL over NVLink on GPUs, gloo in the CPU tests).  Ragged shards are padded to the largest one."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], "local shard does not match shard_range"
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
