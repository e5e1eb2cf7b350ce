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
    m = _mask_cache.get(key)
    if m is None:
        if len(_mask_cache) > 64:
            _mask_cache.clear()
        idx = cell_idx.to(torch.int64).contiguous()
        m = torch.zeros(nvox, dtype=torch.uint8, device=cell_idx.device)
        call("tdb_build_mask", idx.data_ptr(), m.data_ptr(), idx.numel(), nvox, _lib.stream_ptr())
        _mask_cache[key] = m
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
