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


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores next to GPU `device_index` (NVML's ideal CPU affinity) BEFORE it
    allocates pinned host buffers: first touch then puts them on the GPU's own NUMA node, and on a two-socket
    8-GPU box the host<->device copies of the ranks stop crossing the socket interconnect.  Returns the number
    of cores bound, 0 if NVML or the affinity call is unavailable (then nothing changes)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        ids = [d for d in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if d.strip()]
        phys = int(ids[device_index]) if ids and all(d.strip().isdigit() for d in ids) and device_index < len(ids) \
            else device_index
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        cores = [c for c in cores if c < ncpu]
        if not cores:
            return 0
        os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0
