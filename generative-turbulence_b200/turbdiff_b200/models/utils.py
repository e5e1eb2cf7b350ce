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
    if torch.is_grad_enabled() and (cell_values.requires_grad or (other is not None and other.requires_grad)):
        # differentiable use (the learned-variance ELBO differentiates through x_start, ddpm.py:745-747, 853-870): the same
        # select as a torch op, so that autograd sees it; values are identical
        m = inside_mask(cell_idx.to(cell_values.device), nvox).bool().view(cell_values.shape[-3:])
        return torch.where(m, cell_values, torch.zeros_like(cell_values) if other is None else other.expand_as(cell_values))
    a = cell_values.to(torch.float32).contiguous()
    o = None if other is None else other.to(torch.float32).expand_as(a).contiguous()
    mask = inside_mask(cell_idx.to(cell_values.device), nvox)
    out = torch.empty_like(a)
    call("tdb_where_cells", a.data_ptr(), ptr(o), mask.data_ptr(), out.data_ptr(), a.numel() // nvox if nvox else 0, nvox,
         _lib.stream_ptr())
    return out


def fixed_values_of(metadata, variables):
    """The FIXED_VALUE boundary writes of ``OpenFOAMData.grid_embedding`` (data/ofles.py:233-238) as a list of
    ``(voxel_idx, first_channel, values)`` in the reference's write order, read off a reference ``OpenFOAMMetadata``
    (duck-typed: ``boundary_conditions[v][name].type / .value``, ``boundaries[name]["idx"]``, ``v.dims``)."""
    out, f0 = [], 0
    for v in variables:
        for name, desc in metadata.boundary_conditions.get(v, {}).items():
            if getattr(desc.type, "name", str(desc.type)) == "FIXED_VALUE":
                out.append((metadata.boundaries[name]["idx"], f0, torch.as_tensor(desc.value, dtype=torch.float32).reshape(-1)))
        f0 += v.dims
    return out


def boundary_code(cell_idx: torch.Tensor, nvox: int, F: int, fixed_values):
    """Per-voxel code byte + class tables for tdb_scatter_normalize.  Overlapping boundaries are resolved per channel
    in write order (a later write wins, as the reference's sequential index assignments do); voxels with the same
    per-channel winners share a class.  Tiny torch ops on the index tensors, once per geometry."""
    code = inside_mask(cell_idx, nvox).clone()
    if not fixed_values:
        return code, None, None
    cls, bc_has, bc_val = boundary_classes(nvox, F, fixed_values, cell_idx.device)
    code |= (cls << 1).to(torch.uint8)
    return code, bc_has, bc_val


def boundary_classes(nvox: int, F: int, fixed_values, dev):
    """(class per voxel, bc_has [n_cls+1][F] uint8, bc_val [n_cls+1][F] fp32) - pure index arithmetic (see boundary_code)."""
    winner = torch.zeros((F, nvox), dtype=torch.int64, device=dev)  # per channel: 1 + index of the write that wins
    vals = [None]
    for k, (idx, f0, value) in enumerate(fixed_values):
        idx = idx.to(device=dev, dtype=torch.int64)
        winner[f0 : f0 + value.numel(), idx] = k + 1
        vals.append((f0, value.to(dev)))
    key = torch.zeros(nvox, dtype=torch.int64, device=dev)
    for f in range(F):
        key = key * (len(fixed_values) + 1) + winner[f]
    uniq, inv = torch.unique(key, return_inverse=True)  # uniq[0] == 0 (no write) whenever some voxel is untouched
    has_zero = bool(uniq[0] == 0)
    n_cls = uniq.numel() - (1 if has_zero else 0)
    if n_cls > 127:
        raise RuntimeError(f"turbdiff_b200: {n_cls} distinct boundary-value classes (max 127)")
    cls = inv + (0 if has_zero else 1)
    # class tables from one representative voxel per class
    rep = torch.zeros(n_cls + 1, dtype=torch.int64, device=dev)
    rep[cls] = torch.arange(nvox, device=dev)
    w = winner[:, rep]                                   # (F, n_cls + 1)
    bc_has = (w > 0).t().contiguous().to(torch.uint8)    # (n_cls + 1, F)
    bc_has[0] = 0
    bc_val = torch.zeros((n_cls + 1, F), dtype=torch.float32, device=dev)
    for k in range(1, len(vals)):
        f0, value = vals[k]
        for d in range(value.numel()):
            sel = w[f0 + d] == k
            bc_val[sel, f0 + d] = value[d]
    bc_val[0] = 0
    return cls, bc_has, bc_val


def scatter_normalize(samples: torch.Tensor, cell_idx: torch.Tensor, cell_counts, mean: torch.Tensor, std: torch.Tensor,
                      fixed_values=None, tables=None):
    """Fused ``OpenFOAMData.grid_embedding`` + ``Normalization.normalize_grid`` (data/ofles.py:220-240,
    models/normalization.py:20-24): (B, n_cells, F) channels-last cell values -> normalised (B, F, *cell_counts) grid with
    the FIXED_VALUE boundary values written into their padding voxels and every other non-cell voxel holding the
    normalised zero.  One launch; bit-exact with the reference's op sequence.  ``fixed_values``: see
    :func:`fixed_values_of`; ``tables`` = a cached :func:`boundary_code` result for this geometry."""
    _lib.require_cuda(samples, "samples")
    B, n_cells, F = samples.shape
    nvox = int(cell_counts[0]) * int(cell_counts[1]) * int(cell_counts[2])
    s = samples.to(torch.float32).contiguous()
    idx = cell_idx.to(device=s.device, dtype=torch.int64).contiguous()
    mean, std = mean.to(s.device, torch.float32), std.to(s.device, torch.float32)
    scale, shift = torch.reciprocal(std).contiguous(), (-mean / std).contiguous()  # exactly the reference's operands
    grid = torch.empty((B, F, *cell_counts), dtype=torch.float32, device=s.device)
    code, bc_has, bc_val = tables if tables is not None else boundary_code(idx, nvox, F, fixed_values)
    call("tdb_scatter_normalize", s.data_ptr(), idx.data_ptr(), code.data_ptr(), ptr(bc_has), ptr(bc_val), scale.data_ptr(),
         shift.data_ptr(), grid.data_ptr(), B, F, nvox, idx.numel(), _lib.stream_ptr())
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
