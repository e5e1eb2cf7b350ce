"""Multi-GPU plumbing for the two ways the denoising path shards:

* sampling: independent samples, split over ranks, NO data-path communication (an optional final
  gather brings the samples to rank 0 for the single-writer sample store);
* training: data parallel, ONE exchange step - an average all-reduce of the gradients over NCCL
  (NVLink 5 / NVSwitch), issued per bucket so it can overlap the rest of the step.

One process per GPU; `torch.distributed` is the plumbing (backend "nccl" on GPUs, "gloo" in the
CPU tests of this host logic)."""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of `n_items` independent units owned by `rank` (first ranks get the
    remainder), so that concatenating the shards in rank order restores the original order."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_seed(seed: int, rank: int) -> int:
    """Per-shard RNG seed: shard r of a sharded run equals an unsharded run of that sub-batch seeded
    with seed + r."""
    return seed + rank


@torch.no_grad()
def sample_sharded(diffusion, x_bcs: torch.Tensor, C, cell_idx, *, seed: int, gather: bool = False, start_from=None):
    """Ancestral sampling of this rank's slice of the batch.  x_bcs is the FULL batch (same on every
    rank); returns this rank's samples, or with gather=True the full batch on rank 0 (None elsewhere)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    rng = shard_range(x_bcs.shape[0], rank, world)
    mine = x_bcs[rng.start : rng.stop]
    torch.manual_seed(shard_seed(seed, rank))
    out = diffusion.p_sample_loop(mine, C, cell_idx, start_from=start_from) if len(rng) else mine.clone()
    if not gather or world == 1:
        return out
    return gather_shards(out, x_bcs.shape[0])


def gather_shards(local: torch.Tensor, n_total: int, dst: int = 0):
    """Concatenate the per-rank shards (shard_range order) on rank `dst`; None on the other ranks.
    Uneven shards are padded to the largest one for the collective and trimmed afterwards."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [len(shard_range(n_total, r, world)) for r in range(world)]
    big = max(sizes)
    padded = torch.zeros((big, *local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat([part[:n] for part, n in zip(parts, sizes)])


class GradientAllReduce:
    """Bucketed average all-reduce of `.grad` over the data-parallel group.

    Parameters are packed (in registration order) into flat fp32 buckets of ~`bucket_mb`; every
    bucket is one asynchronous all-reduce, so the first buckets travel while later gradients are
    still being unpacked / clipped.  With equal per-rank batch sizes the average of the rank losses'
    gradients equals the gradient of the global-batch loss (the loss is a mean over samples)."""

    def __init__(self, params, bucket_mb: float = 64.0, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets: list[list[torch.nn.Parameter]] = []
        cur, cur_bytes, cap = [], 0, int(bucket_mb * 2**20)
        for p in self.params:
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= cap:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat: list[torch.Tensor | None] = [None] * len(self.buckets)

    def __call__(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        work = []
        for i, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            flat = self._flat[i]
            if flat is None or flat.device != bucket[0].device:
                flat = self._flat[i] = torch.empty(n, dtype=torch.float32, device=bucket[0].device)
            off = 0
            for p in bucket:
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                flat[off : off + p.numel()].copy_(g.reshape(-1))
                off += p.numel()
            work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for i, bucket in enumerate(self.buckets):
            work[i].wait()
            flat = self._flat[i]
            off = 0
            for p in bucket:
                g = flat[off : off + p.numel()].view_as(p) / world
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += p.numel()
