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
    """Average all-reduce of the gradients over the data-parallel group (the ONE exchange step of the training path).

    Fast path (``attach(denoiser)``): the backward launch program already leaves every denoiser gradient in ONE flat
    fp32 buffer (engine._capture_train), so the exchange is a handful of large NCCL all-reduces over slices of that
    buffer, issued from inside ``backward`` before the gradients are handed to autograd - no per-parameter packing,
    no ~420 small torch ops on the host.  Parameters that are not part of an attached denoiser (e.g. the task's
    cell-type embedding) and eager-mode training use the generic path: flat fp32 buckets of ~``bucket_mb``, one
    asynchronous all-reduce per bucket.  With equal per-rank batch sizes the average of the rank losses' gradients
    equals the gradient of the global-batch loss (the loss is a mean over samples, ddpm.py:848-852)."""

    def __init__(self, params, bucket_mb: float = 64.0, group=None, flat_chunks: int = 4):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_mb = bucket_mb
        self.flat_chunks = max(1, int(flat_chunks))
        self._flat_synced: set[int] = set()   # ids of parameters whose gradient was already reduced inside backward
        self._attached = []
        self._buckets_for = None
        self.buckets: list[list[torch.nn.Parameter]] = []
        self._flat: list[torch.Tensor | None] = []
        self._make_buckets(self.params)

    def _make_buckets(self, params):
        self.buckets = []
        cur, cur_bytes, cap = [], 0, int(self.bucket_mb * 2**20)
        for p in params:
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= cap:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)
        self._buckets_for = tuple(id(p) for p in params)

    # ---- fast path: all-reduce the backward program's flat gradient buffer ----------------------------------------
    def attach(self, denoiser):
        """Reduce `denoiser`'s gradients inside its backward (engine.grad_sync hook: ``start(flat_part)`` is called once the
        phase-1 gradients are complete and again after phase 2, ``finish()`` before the gradients are handed to autograd).
        Returns self."""
        eng = denoiser.engine()
        ids = {id(p) for p in denoiser.parameters()}
        outer = self

        class _Hook:
            def __init__(self):
                self.pending = []

            def start(self, flat: torch.Tensor):
                if outer._world() > 1 and flat.numel() > 0:
                    self.pending.append(outer.start_flat(flat))

            def finish(self):
                for item in self.pending:
                    outer.finish_flat(item)
                self.pending = []
                if outer._world() > 1:
                    outer._flat_synced |= ids

            def __call__(self, flat: torch.Tensor):  # whole buffer at once
                self.start(flat)
                self.finish()

        eng.grad_sync = _Hook()
        self._attached.append(denoiser)
        return self

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def start_flat(self, flat: torch.Tensor):
        """Start the in-place average of a flat gradient buffer over the group: asynchronous all-reduces over chunks of
        ~1/flat_chunks of the denoiser's gradients (the first chunks travel while the later ones are being enqueued and,
        for the phase-1 part, while the rest of the backward pass runs)."""
        avg = dist.get_backend(self.group) == "nccl"
        n = flat.numel()
        total = sum(p.numel() for d in self._attached for p in d.parameters()) or n
        step = max(1, -(-total // self.flat_chunks))
        work = [dist.all_reduce(flat[i : i + step], op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
                for i in range(0, n, step)]
        return flat, work, avg

    def finish_flat(self, item):
        flat, work, avg = item
        for w in work:
            w.wait()
        if not avg:
            flat.div_(self._world())

    def reduce_flat(self, flat: torch.Tensor):
        """In-place average of a flat gradient buffer over the group (start_flat + finish_flat)."""
        self.finish_flat(self.start_flat(flat))

    # ---- generic path --------------------------------------------------------------------------------------------
    def __call__(self):
        if self._world() == 1:
            self._flat_synced.clear()
            return
        todo = [p for p in self.params if id(p) not in self._flat_synced]
        self._flat_synced.clear()
        if not todo:
            return
        if self._buckets_for != tuple(id(p) for p in todo):
            self._make_buckets(todo)
        world = self._world()
        work = []
        for i, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            flat = self._flat[i]
            if flat is None or flat.device != bucket[0].device:
                flat = self._flat[i] = torch.empty(n, dtype=torch.float32, device=bucket[0].device)
            views = list(flat.split([p.numel() for p in bucket]))
            torch._foreach_copy_(views, [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
            work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for i, bucket in enumerate(self.buckets):
            work[i].wait()
            flat = self._flat[i]
            flat.div_(world)
            views = [v.view_as(p) for v, p in zip(flat.split([p.numel() for p in bucket]), bucket)]
            for p, v in zip(bucket, views):
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in bucket], views)
