"""Inside-cell gather/scatter helpers (reference: models/utils.py:8-28), CUDA-backed and
bit-exact: ``where_cells`` is order independent, so it runs as a byte-mask select;
``select_cells`` keeps the order of ``cell_idx``."""

from __future__ import annotations

import torch

from .. import _lib
from .._lib import call, ptr

_mask_cache: dict = {}


def broadcast_right(x: torch.Tensor, other: torch.Tensor):
    assert other.ndim >= x.ndim
    return x.reshape(*x.shape, *((1,) * (other.ndim - x.ndim)))


def ravel_cells(x: torch.Tensor):
    return x.flatten(start_dim=-3)


def inside_mask(cell_idx: torch.Tensor, nvox: int) -> torch.Tensor:
    """uint8 (nvox,) mask with 1 on the voxels listed in cell_idx (cached per index tensor)."""
    _lib.require_cuda(cell_idx, "cell_idx")
    key = (cell_idx.data_ptr(), cell_idx.numel(), cell_idx._version, nvox, str(cell_idx.device))
    hit = _mask_cache.get(key)
    # an entry is valid only for the very tensor object it was built from: the cache keeps that tensor alive, so its
    # address cannot be recycled for another geometry with the same cell count while the entry exists
    if hit is not None and hit[0] is cell_idx:
        return hit[1]
    if len(_mask_cache) >= 8:
        _mask_cache.pop(next(iter(_mask_cache)))
    idx = cell_idx.to(torch.int64).contiguous()
    m = torch.zeros(nvox, dtype=torch.uint8, device=cell_idx.device)
    call("tdb_build_mask", idx.data_ptr(), m.data_ptr(), idx.numel(), nvox, _lib.stream_ptr())
    _mask_cache[key] = (cell_idx, m)
    return m


def select_cells(x: torch.Tensor, cell_idx: torch.Tensor):
    _lib.require_cuda(x, "x")
    nvox = x.shape[-3] * x.shape[-2] * x.shape[-1]
    lead = x.shape[:-3]
    xf = x.to(torch.float32).contiguous()
    idx = cell_idx.to(device=x.device, dtype=torch.int64).contiguous()
    rows = xf.numel() // nvox if nvox else 0
    out = torch.empty((*lead, idx.numel()), dtype=torch.float32, device=x.device)
    call("tdb_select_cells", xf.data_ptr(), idx.data_ptr(), out.data_ptr(), rows, nvox, idx.numel(), _lib.stream_ptr())
    return out


def where_cells(cell_idx, cell_values: torch.Tensor, other: torch.Tensor | None = None):
    _lib.require_cuda(cell_values, "cell_values")
    nvox = cell_values.shape[-3] * cell_values.shape[-2] * cell_values.shape[-1]
    a = cell_values.to(torch.float32).contiguous()
    o = None if other is None else other.to(torch.float32).expand_as(a).contiguous()
    mask = inside_mask(cell_idx.to(cell_values.device), nvox)
    out = torch.empty_like(a)
    call("tdb_where_cells", a.data_ptr(), ptr(o), mask.data_ptr(), out.data_ptr(), a.numel() // nvox if nvox else 0, nvox,
         _lib.stream_ptr())
    return out


def scatter_normalize(samples: torch.Tensor, cell_idx: torch.Tensor, cell_counts, mean: torch.Tensor, std: torch.Tensor):
    """Fused ``OpenFOAMData.grid_embedding`` + ``Normalization.normalize_grid`` (data/ofles.py:220-232,
    models/normalization.py:20-24): (B, n_cells, F) channels-last cell values -> normalised (B, F, *cell_counts) grid,
    non-cell voxels holding the normalised zero.  One launch; bit-exact with the reference's op sequence.  (FIXED_VALUE
    boundary cells, if any, are written by the caller: ``grid[..., f, idx] = addcmul(-mean/std, 1/std, value)``.)"""
    _lib.require_cuda(samples, "samples")
    B, n_cells, F = samples.shape
    nvox = int(cell_counts[0]) * int(cell_counts[1]) * int(cell_counts[2])
    s = samples.to(torch.float32).contiguous()
    idx = cell_idx.to(device=s.device, dtype=torch.int64).contiguous()
    mean, std = mean.to(s.device, torch.float32), std.to(s.device, torch.float32)
    scale, shift = torch.reciprocal(std).contiguous(), (-mean / std).contiguous()  # exactly the reference's operands
    grid = torch.empty((B, F, *cell_counts), dtype=torch.float32, device=s.device)
    call("tdb_scatter_normalize", s.data_ptr(), idx.data_ptr(), inside_mask(idx, nvox).data_ptr(), scale.data_ptr(), shift.data_ptr(),
         grid.data_ptr(), B, F, nvox, idx.numel(), _lib.stream_ptr())
    return grid


def gather_denormalize(x: torch.Tensor, cell_idx: torch.Tensor, mean: torch.Tensor, std: torch.Tensor):
    """Fused ``Normalization.denormalize_grid`` + ``select_cells`` + channels-last rearrangement of
    ``SampleStore.add_samples`` (models/normalization.py:26-30, models/utils.py:14-15, models/metrics.py:50-57):
    (B, F, X, Y, Z) samples -> de-normalised (B, n_cells, F) cell values."""
    _lib.require_cuda(x, "x")
    B, F = x.shape[:2]
    nvox = x.shape[-3] * x.shape[-2] * x.shape[-1]
    xf = x.to(torch.float32).contiguous()
    idx = cell_idx.to(device=x.device, dtype=torch.int64).contiguous()
    scale, shift = std.to(x.device, torch.float32).contiguous(), mean.to(x.device, torch.float32).contiguous()
    out = torch.empty((B, idx.numel(), F), dtype=torch.float32, device=x.device)
    call("tdb_gather_denormalize", xf.data_ptr(), idx.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), B, F, nvox,
         idx.numel(), _lib.stream_ptr())
    return out
