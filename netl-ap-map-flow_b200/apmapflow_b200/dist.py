"""Host-side plumbing for one-process-per-GPU runs (torch.distributed; NCCL on GPUs, gloo in the CPU
tests).  The LCL path shards by independent cases / maps: every rank owns a contiguous block of the
work list and no data-path collective is needed; only the timing (max over ranks) and the result
bookkeeping (gather of per-case scalars) cross ranks."""
import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard(items, rank, world_size):
    """Contiguous, balanced block of `items` for `rank` (first `len % world` ranks get one extra).
    Concatenating the blocks of ranks 0..world-1 restores the list, so results gather in order."""
    n = len(items)
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return items[start:start + base + (1 if rank < extra else 0)]


def max_over_ranks(value, device=None):
    """MAX all-reduce of a host float (the timing rule: a multi-GPU number is the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj):
    """All ranks' picklable `obj` as a list ordered by rank (rank 0 prints / writes the summary)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
