"""Oracle (test infrastructure): synthetic batches built from the reference's OWN dataclasses,
so the unmodified ``DiffusionTraining`` / ``OpenFOAMData.grid_embedding`` / ``Conditioning`` / ``Normalization`` run on
them without HDF5 files.  Geometry = oracle.grid_ref.channel_geometry (inlet / outlet / wall voxel sets as
scripts/grid-embedding.py:56-62 defines them), boundary conditions of the LES template (scripts/foam2h5.py:139-147)."""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import grid_ref


def make_batch(ns, cells=(32, 16, 16), hole=((4, 8), (5, 11), (0, 10)), batch=2, seed=0, device="cpu", u_in=20.0):
    """(OpenFOAMBatch, geometry).  `ns` = oracle.ref_shim.load()."""
    ofles = ns.ofles
    V, BC = ofles.Variable, ofles.BoundaryCondition
    geo = grid_ref.channel_geometry(cells=cells, hole=hole, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 17))
    n = len(geo.cell_idx)
    u = torch.from_numpy((rng.standard_normal((batch, n, 3)) * 3.0 + np.array([u_in / 2, 0, 0])).astype(np.float32)).to(device)
    p = torch.from_numpy((rng.standard_normal((batch, n, 1)) * 40.0).astype(np.float32)).to(device)
    bnd = {k: {"type": "patch", "idx": torch.from_numpy(v).to(device)} for k, v in geo.boundaries.items()}
    bcs = {
        V.U: {"inlets": BC(BC.Type.FIXED_VALUE, torch.tensor([u_in, 0.0, 0.0], device=device)),
              "outlets": BC(BC.Type.INLET_OUTLET),
              "walls": BC(BC.Type.FIXED_VALUE, torch.tensor([0.0, 0.0, 0.0], device=device))},
        V.P: {"inlets": BC(BC.Type.ZERO_GRADIENT), "outlets": BC(BC.Type.FIXED_VALUE, torch.tensor([0.0], device=device)),
              "walls": BC(BC.Type.ZERO_GRADIENT)},
    }
    md = ofles.OpenFOAMMetadata(file=Path("/synthetic/case-0/data.h5"), nu=1e-5, h=torch.ones(3), cell_counts=np.array(geo.padded),
                                cell_idx=torch.from_numpy(geo.cell_idx).to(device), boundaries=bnd, boundary_conditions=bcs, holes=[])
    data = ofles.OpenFOAMData(md, torch.zeros(batch, device=device), {V.U: u, V.P: p})

    def st(x):
        x = x.reshape(-1, x.shape[-1])
        return {"mean": x.mean(0), "std": x.std(0), "min": x.min(0).values, "max": x.max(0).values}

    stats = ofles.OpenFOAMStats({"u": st(u), "p": st(p), "norm(u)": st(u.norm(dim=-1, keepdim=True))})
    return ofles.OpenFOAMBatch(data, stats), geo
