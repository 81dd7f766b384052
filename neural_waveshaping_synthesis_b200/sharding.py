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
    if min(sizes) == max(sizes):
        # equal shards: one collective straight into the rows of the full batch (no staging list, no concatenation —
        # at 2048 x 64000 samples the extra copy cost a third of the gather)
        full = local.new_empty((total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
        return full
    big = max(sizes)
    padded = local if local.shape[0] == big else torch.cat(
        [local, local.new_zeros((big - local.shape[0],) + tuple(local.shape[1:]))])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def wave_bounds(total: int, rank: int, world: int, waves: int, wave: int) -> Tuple[int, int]:
    """[lo, hi) of the utterances `rank` renders in wave `wave` when a job of `total` utterances is rendered in `waves`
    waves: wave w covers the contiguous rows [w * total / waves, (w + 1) * total / waves), split evenly over the ranks —
    so the all-gather of one wave fills one contiguous block of the full batch.  `total` must divide by world * waves."""
    if total % (world * waves) or not (0 <= rank < world) or not (0 <= wave < waves):
        raise ValueError("bad wave request total=%d rank=%d world=%d waves=%d wave=%d" % (total, rank, world, waves, wave))
    per = total // (world * waves)
    lo = wave * (total // waves) + rank * per
    return lo, lo + per


def forward_and_gather(forward, f0: torch.Tensor, control: torch.Tensor, total: int, waves: int = 2,
                       group: Optional[dist.ProcessGroup] = None, **kwargs) -> torch.Tensor:
    """Renders this rank's utterances in `waves` waves and assembles the full [total, N] batch on every rank, the gather
    of wave w (asynchronous, on the collective's own stream, straight into its rows of the result) overlapping the
    forward of wave w + 1.  `f0` / `control` hold this rank's utterances wave after wave (wave_bounds order);
    `forward(f0, control, **kwargs)` is the module call.  Without a process group it is one plain forward."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return forward(f0, control, **kwargs)
    world = dist.get_world_size(group)
    if total % (world * waves) or f0.shape[0] * world != total:
        raise ValueError("forward_and_gather: %d utterances do not split into %d ranks x %d waves" % (total, world, waves))
    per = total // (world * waves)
    full, pending = None, []
    for w in range(waves):
        y = forward(f0[w * per:(w + 1) * per], control[w * per:(w + 1) * per], **kwargs)
        if full is None:
            full = y.new_empty((total,) + tuple(y.shape[1:]))
        rows = full[w * (total // waves):(w + 1) * (total // waves)]
        pending.append((dist.all_gather_into_tensor(rows, y.contiguous(), group=group, async_op=True), y))
    for work, _ in pending:
        work.wait()
    return full
