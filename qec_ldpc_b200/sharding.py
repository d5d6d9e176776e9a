"""Multi-GPU plumbing: frames are independent, so a run shards by contiguous GLOBAL frame-id ranges (disjoint Philox
streams, counters independent of the GPU count) and the only collective is one all-reduce of the counter vector.
One process per GPU under torch.distributed (NCCL on GPUs; the same code runs on gloo for the CPU tests)."""
import numpy as np


def shard_range(total_frames, rank, world_size):
    """Contiguous share [first, first+count) of `total_frames` for `rank` (matches the C++ CLI's split)."""
    lo = total_frames * rank // world_size
    hi = total_frames * (rank + 1) // world_size
    return lo, hi - lo


def allreduce_counters(counters, device=None):
    """Sum of the uint64 counter vector over all ranks (exactly one collective; < 128 bytes, latency bound)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(counters, np.uint64).copy()
    t = torch.from_numpy(np.asarray(counters, np.uint64).astype(np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def allreduce_max(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_sharded(stats_fn, total_frames, first_frame=0, device=None):
    """stats_fn(first_frame, nframes) -> counters of that range on this rank's device; returns the global counters."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    lo, cnt = shard_range(total_frames, rank, world)
    local = stats_fn(first_frame + lo, cnt) if cnt > 0 else np.zeros(12, np.uint64)
    return allreduce_counters(local, device)
