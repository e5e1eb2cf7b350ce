"""Synthetic workload of the shapes configuration for benchmarks and profiling (no dataset, no checkpoint).

A padded channel grid (192x48x48 cells + one padding layer per side = 194x50x50, scripts/grid-embedding.py:45-69 of the
reference) with one solid pillar cut out (scripts/generate-performance-dataset.py:25), the in-domain cell list in a
mesh-like unsorted order, a cell-type map (outside / inside / inlet / outlet / wall) embedded through a random 6x4 table
as the local conditioning, and Gaussian boundary-value fields.  Only shapes and index structure matter here; parity
against the reference's own helpers is the job of tests/ and oracle/."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

CELLS = (192, 48, 48)
PILLAR = ((12, 24), (16, 32), (0, 32))
OUTSIDE, INSIDE, INLET, OUTLET, WALL = 0, 1, 2, 3, 4


@dataclass
class SyntheticGeometry:
    padded: tuple
    cell_idx: np.ndarray   # (n_cells,) int64 flat voxel indices of the in-domain cells
    cell_type: np.ndarray  # padded-grid int64 map


def channel_geometry(cells=CELLS, pillar=PILLAR, seed: int = 0) -> SyntheticGeometry:
    nx, ny, nz = cells
    P = (nx + 2, ny + 2, nz + 2)
    fluid = np.zeros(P, dtype=bool)
    fluid[1:-1, 1:-1, 1:-1] = True
    if pillar is not None:
        (x0, x1), (y0, y1), (z0, z1) = pillar
        fluid[1 + x0 : 1 + x1, 1 + y0 : 1 + y1, 1 + z0 : 1 + z1] = False
    idx = np.flatnonzero(fluid.reshape(-1)).astype(np.int64)
    rng = np.random.Generator(np.random.PCG64(seed))
    # block-wise order with some blocks permuted: gathers / scatters must not rely on sorted indices
    blocks = np.array_split(idx, 8)
    idx = np.concatenate([blocks[i][rng.permutation(len(blocks[i]))] if i % 2 else blocks[i] for i in rng.permutation(len(blocks))])
    # solid voxels face-adjacent to fluid are boundaries: inlet on the x = 0 plane, outlet on x = X-1, walls elsewhere
    near = np.zeros(P, dtype=bool)
    for ax in range(3):
        for shift in (-1, 1):
            rolled = np.roll(fluid, shift, axis=ax)
            edge = [slice(None)] * 3
            edge[ax] = 0 if shift == 1 else -1
            rolled[tuple(edge)] = False
            near |= rolled
    boundary = near & ~fluid
    ct = np.full(P, OUTSIDE, dtype=np.int64)
    ct[fluid] = INSIDE
    ct[boundary] = WALL
    ct[0][boundary[0]] = INLET
    ct[-1][boundary[-1]] = OUTLET
    return SyntheticGeometry(P, idx, ct)


def synthetic_inputs(batch: int, seed: int, cells=CELLS, pillar=PILLAR):
    """(geometry, x_bcs (B, 4, X, Y, Z) fp32 ~ N(0, 1), c_local (4, X, Y, Z) fp32 cell-type embedding) on the host."""
    geo = channel_geometry(cells, pillar, seed=0)
    rng = np.random.Generator(np.random.PCG64(seed))
    table = rng.standard_normal((6, 4)).astype(np.float32)
    c_local = torch.from_numpy(np.ascontiguousarray(np.moveaxis(table[geo.cell_type], -1, 0)))
    x = torch.from_numpy(rng.standard_normal((batch, 4, *geo.padded)).astype(np.float32))
    return geo, x, c_local


def conv_flops_per_sample(spatial, dim: int = 32, levels: int = 4, in_features: int = 4, c_local_features: int = 4, out_features: int = 4,
                          attn_hidden: int = 128, haloed: bool = False) -> float:
    """Algorithmic FLOPs of one denoiser forward: 2*Cin*Cout*k^3*voxels over every 3x3x3 and 1x1x1 convolution
    (666.2 GFLOP at the shapes configuration).  haloed=True counts the rows the implicit-GEMM kernels actually process:
    the GEMM's M dimension runs over the halo grids ((X+2)(Y+2)(Z+2) rows per level), whose halo rows are computed and
    discarded - 9 % more work at level 0, 3.2x at the 12x3x3 bottleneck."""
    from .engine import level_sizes

    vox = [int(np.prod([d + 2 for d in s] if haloed else s)) for s in level_sizes(tuple(spatial), levels)]

    def block(cin, cout, lvl):
        f = 2.0 * 27 * vox[lvl] * (cin * cout + cout * cout)
        return f + (2.0 * vox[lvl] * cin * cout if cin != cout else 0.0)

    c0 = dim + (dim if c_local_features > 0 else 0)
    total = 2.0 * vox[0] * (in_features * dim + c_local_features * dim)  # 1x1 encoders
    total += block(c0, 2 * dim, 0) + sum(block(dim * 2**l, dim * 2 ** (l + 1), l) for l in range(1, levels))
    cd = dim * 2**levels
    total += 2 * block(cd, cd, levels) + 2.0 * vox[levels] * (cd * 3 * attn_hidden + attn_hidden * cd)
    total += sum(block(2 * dim * 2 ** (l + 1), dim * 2**l, l) for l in range(levels))
    total += block(dim, dim, 0) + 2.0 * vox[0] * dim * out_features
    return total
