"""Shared helpers for the GPU parity tests."""

from __future__ import annotations

import contextlib

import numpy as np
import torch
import torch.nn.functional as F


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def to_halo(x: torch.Tensor, dtype=torch.float32, ld: int | None = None, c0: int = 0, pad_rows: int = 0) -> torch.Tensor:
    """NCDHW -> halo grid [B, X+2, Y+2, Z+2, ld] (replicate halo), channels at [c0, c0+C).
    pad_rows > 0: the grid is a view into a buffer with that many NaN rows in front and behind
    (tdb_conv3d_bf16_fold reads, but must never use, such padding)."""
    B, C = x.shape[:2]
    xp = F.pad(x.float(), (1, 1, 1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 4, 1)
    ld = ld or C
    rows = xp.shape[0] * xp.shape[1] * xp.shape[2] * xp.shape[3]
    flat = torch.full((rows + 2 * pad_rows, ld), float("nan") if pad_rows else 0.0, dtype=dtype, device=x.device)
    out = flat[pad_rows : pad_rows + rows].view(*xp.shape[:4], ld)
    out.zero_()
    out[..., c0 : c0 + C] = xp.to(dtype)
    return out


def from_halo(g: torch.Tensor, C: int | None = None, c0: int = 0) -> torch.Tensor:
    """Interior of a halo grid -> NCDHW fp32."""
    C = C or g.shape[-1]
    return g[:, 1:-1, 1:-1, 1:-1, c0 : c0 + C].permute(0, 4, 1, 2, 3).float().contiguous()


def halo_is_replicated(g: torch.Tensor) -> bool:
    inner = g[:, 1:-1, 1:-1, 1:-1, :].permute(0, 4, 1, 2, 3).float()
    want = F.pad(inner, (1, 1, 1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 4, 1)
    return bool(torch.equal(want, g.float()))


@contextlib.contextmanager
def cpu_seeded_randn(seed: int):
    """Make torch.randn_like / torch.randint draw from a CPU generator (the stream the CPU
    reference consumed when the golden vectors were made) regardless of the tensor's device."""
    gen = torch.Generator().manual_seed(seed)
    real_randn_like, real_randint = torch.randn_like, torch.randint

    def randn_like(x, **kw):
        return torch.randn(x.shape, generator=gen, dtype=torch.float32).to(x.device)

    def randint(low, high, size, **kw):
        dev = kw.pop("device", None)
        return real_randint(low, high, size, generator=gen, dtype=kw.get("dtype", torch.long)).to(dev or "cpu")

    torch.randn_like, torch.randint = randn_like, randint
    try:
        yield
    finally:
        torch.randn_like, torch.randint = real_randn_like, real_randint


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))
