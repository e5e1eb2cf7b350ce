"""Oracle (test infrastructure, CPU, numpy): the integer/indexing steps either side of
the denoising path, plus the synthetic channel-with-a-hole geometry used everywhere.

Follows ``data/ofles.py:220-240`` (grid_embedding scatter + FIXED_VALUE boundary
values), ``models/cell_type_embeddings.py:29-58`` (cell-type map),
``models/normalization.py:20-30`` and ``scripts/grid-embedding.py:41-90`` (cell_idx and
boundary index conventions: one padding layer, row-major ravel over the padded grid,
boundary voxel = fluid cell + outward face normal).

All results here are exact (integer indexing / copies): parity is bit-exact.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

CELL_TYPES = {"inside": 0, "outside": 1, "walls": 2, "inlets": 3, "outlets": 4, "empties": 5}


@dataclass
class Geometry:
    """A padded voxel grid with its inside cells and boundary voxel sets."""

    padded: tuple[int, int, int]
    cell_idx: np.ndarray  # int64 (n_cells,), flat indices into the padded grid, mesh order
    boundaries: dict[str, np.ndarray] = field(default_factory=dict)  # name -> int64 idx (sorted by name)

    @property
    def n_vox(self) -> int:
        return int(np.prod(self.padded))


def channel_geometry(cells=(32, 16, 16), hole=((4, 8), (5, 11), (0, 10)), seed=0, shuffle=True) -> Geometry:
    """Channel of ``cells`` fluid cells padded by one layer per side, with one solid
    box ``hole`` = ((x0,x1),(y0,y1),(z0,z1)) in cell coordinates cut out.  Inlet at
    x=0 padding plane, outlet at x=X-1, walls on the y/z padding planes and on solid voxels
    face-adjacent to fluid (grid-embedding.py:56-62).  ``cell_idx`` is emitted in a
    shuffled (mesh-like, unsorted) order so order-dependent gathers are exercised."""

    nx, ny, nz = cells
    P = (nx + 2, ny + 2, nz + 2)
    fluid = np.zeros(P, dtype=bool)
    fluid[1:-1, 1:-1, 1:-1] = True
    if hole is not None:
        (x0, x1), (y0, y1), (z0, z1) = hole
        fluid[1 + x0 : 1 + x1, 1 + y0 : 1 + y1, 1 + z0 : 1 + z1] = False
    coords = np.argwhere(fluid)
    idx = np.ravel_multi_index(coords.T, P).astype(np.int64)
    if shuffle:
        rng = np.random.Generator(np.random.PCG64(seed))
        # block-wise shuffle: mimic OpenFOAM's block-by-block cell order
        blocks = np.array_split(idx, 7)
        order = rng.permutation(len(blocks))
        idx = np.concatenate([blocks[i][rng.permutation(len(blocks[i]))] if i % 2 else blocks[i] for i in order])

    inlets, outlets, walls = set(), set(), set()
    for d, (ax, step) in enumerate([(0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1)]):
        nb = coords.copy()
        nb[:, ax] += step
        is_solid = ~fluid[nb[:, 0], nb[:, 1], nb[:, 2]]
        nb = nb[is_solid]
        flat = np.ravel_multi_index(nb.T, P)
        if ax == 0 and step == -1:
            on_plane = nb[:, 0] == 0
            inlets.update(flat[on_plane].tolist())
            walls.update(flat[~on_plane].tolist())
        elif ax == 0 and step == 1:
            on_plane = nb[:, 0] == P[0] - 1
            outlets.update(flat[on_plane].tolist())
            walls.update(flat[~on_plane].tolist())
        else:
            walls.update(flat.tolist())
    b = {
        "inlets": np.array(sorted(inlets), dtype=np.int64),
        "outlets": np.array(sorted(outlets), dtype=np.int64),
        "walls": np.array(sorted(walls), dtype=np.int64),
    }
    return Geometry(P, idx, b)


def cell_type_map(geo: Geometry) -> np.ndarray:
    """outside everywhere, inside on cell_idx, then each boundary in dict order
    (cell_type_embeddings.py:47-58)."""
    ct = np.full(geo.n_vox, CELL_TYPES["outside"], dtype=np.int64)
    ct[geo.cell_idx] = CELL_TYPES["inside"]
    for name, idx in geo.boundaries.items():
        ct[idx] = CELL_TYPES[name]
    return ct.reshape(geo.padded)


def cell_type_embedding(geo: Geometry, table: np.ndarray) -> np.ndarray:
    """Embedding(6, dim) lookup moved to channels-first (cell_type_embeddings.py:66-69)."""
    return np.moveaxis(table[cell_type_map(geo)], -1, 0)


def grid_embedding(geo: Geometry, samples: list[np.ndarray], fixed_values: list[dict[str, np.ndarray]]) -> np.ndarray:
    """Scatter per-cell samples into a zero dense grid and write FIXED_VALUE boundary
    vectors (ofles.py:220-240).  ``samples[v]``: (B, n_cells, dims_v);
    ``fixed_values[v]``: boundary name -> (dims_v,) vector for that variable."""
    B = samples[0].shape[0]
    F = sum(s.shape[-1] for s in samples)
    x = np.zeros((B, F, geo.n_vox), dtype=np.float32)
    f0 = 0
    for s, fv in zip(samples, fixed_values):
        d = s.shape[-1]
        x[:, f0 : f0 + d, geo.cell_idx] = np.swapaxes(s, 1, 2)
        for name, value in fv.items():
            x[:, f0 : f0 + d, geo.boundaries[name]] = np.asarray(value, dtype=np.float32)[None, :, None]
        f0 += d
    return x.reshape(B, F, *geo.padded)


def select_cells(x: np.ndarray, cell_idx: np.ndarray) -> np.ndarray:
    """models/utils.py:14-15."""
    return x.reshape(*x.shape[:-3], -1)[..., cell_idx]


def where_cells(cell_idx, cell_values, other=None):
    """models/utils.py:22-28."""
    out = np.zeros_like(cell_values) if other is None else other.copy()
    of = out.reshape(*out.shape[:-3], -1)
    of[..., cell_idx] = cell_values.reshape(*cell_values.shape[:-3], -1)[..., cell_idx]
    return out


def inside_mask(geo: Geometry) -> np.ndarray:
    m = np.zeros(geo.n_vox, dtype=np.uint8)
    m[geo.cell_idx] = 1
    return m


def normalize_grid(x, mean, std):
    """addcmul(-mean/std, 1/std, x) per channel, fp32 (normalization.py:20-24)."""
    mean = np.asarray(mean, np.float32).reshape(-1, 1, 1, 1)
    std = np.asarray(std, np.float32).reshape(-1, 1, 1, 1)
    return (-mean / std) + (np.float32(1) / std) * x


def denormalize_grid(x, mean, std):
    """addcmul(mean, std, x) (normalization.py:26-30)."""
    mean = np.asarray(mean, np.float32).reshape(-1, 1, 1, 1)
    std = np.asarray(std, np.float32).reshape(-1, 1, 1, 1)
    return mean + std * x
