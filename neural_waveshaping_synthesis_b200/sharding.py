"""Multi-GPU plumbing for the NWS forward: utterances are independent (no cross-batch op in
NeuralWaveshaping.forward, neural_waveshaping.py:74-90), so a batch shards into contiguous
per-rank slices with the ~2 MB of weights / LUT replicated.  No collective runs on the data path;
the only exchanges are the final throughput reduction and, if the caller wants the audio on one
rank, one gather.  Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` utterances for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request total=%d rank=%d world=%d" % (total, rank, world))
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(step_ms: float, samples: float, device: torch.device,
                         group: Optional[dist.ProcessGroup] = None) -> Tuple[float, float]:
    """Whole-job figures: time = max over ranks (device-timed per rank), samples = sum over ranks."""
    t = torch.tensor([step_ms], dtype=torch.float64, device=device)
    n = torch.tensor([samples], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), float(n.item())


def gather_audio(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All ranks receive the full [total, N] batch assembled from the per-rank shards (uneven shards
    are padded to the largest for the collective and trimmed afterwards)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d utterances, expected %d" % (rank, local.shape[0], sizes[rank]))
    big = max(sizes)
    padded = local if local.shape[0] == big else torch.cat(
        [local, local.new_zeros((big - local.shape[0],) + tuple(local.shape[1:]))])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])
