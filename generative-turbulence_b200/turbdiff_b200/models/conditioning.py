"""The interface type of the conditioning dict ``C`` (reference: models/conditioning.py:15-93).

Only the pieces the denoising path touches: the ``Type`` enum used as dict keys and the two
helpers that split ``C`` into local (per-voxel, unbatched ``(c, X, Y, Z)``) and global parts.
Keys are duck-typed on ``.local`` / ``.global_`` so the reference's own enum members work."""

from __future__ import annotations

import enum

import torch


class ConditioningType(enum.Enum):
    CELL_TYPE = enum.auto()
    CELL_POS = enum.auto()

    @property
    def local(self) -> bool:
        return True

    @property
    def global_(self) -> bool:
        return False


class Conditioning:
    Type = ConditioningType


def local_conditioning(C) -> torch.Tensor | None:
    parts = [v for k, v in C.items() if getattr(k, "local", True)]
    return torch.cat(parts, dim=0) if parts else None


def global_conditioning(C) -> torch.Tensor | None:
    parts = [v for k, v in C.items() if getattr(k, "global_", False)]
    return torch.cat(parts, dim=0) if parts else None
