"""Oracle (test infrastructure): the named, fully deterministic test cases shared by the
golden generator and the parity tests.  Inputs come from numpy PCG64 seeds only."""

from __future__ import annotations

import numpy as np
import torch

from . import grid_ref
from .unet_ref import UNetSpec

CASES = {
    # smallest case with the full structure: GroupNorm(8), 2 levels, attention centre.
    "micro": dict(
        spec=UNetSpec(dim=8, u_net_levels=2, timesteps=10, groups=8),
        cells=(16, 8, 8), hole=((3, 7), (2, 6), (0, 5)), batch=2, seed=101, save_taps=True,
    ),
    # BASELINE.json configs[0]: tiny TurbDiff (2 levels, 16 ch) on a 32x16x16 grid, B=2.
    "tiny": dict(
        spec=UNetSpec(dim=16, u_net_levels=2, timesteps=10, groups=8),
        cells=(32, 16, 16), hole=((4, 8), (5, 11), (0, 10)), batch=2, seed=202, save_taps=False,
    ),
    # dim=32 (the channel plan of the shapes config: 64/128/256 ...) on a small grid with
    # 3 levels: exercises every channel width the tcgen05 path is specialised for.
    "dim32": dict(
        spec=UNetSpec(dim=32, u_net_levels=3, timesteps=500, groups=8),
        cells=(30, 14, 12), hole=((5, 9), (3, 8), (0, 7)), batch=2, seed=303, save_taps=False,
    ),
    # other normalisation variants of ddpm.py:424-431
    "micro-layer": dict(
        spec=UNetSpec(dim=8, u_net_levels=2, timesteps=10, groups=1),
        cells=(16, 8, 8), hole=None, batch=1, seed=404, save_taps=False,
    ),
    "micro-instance": dict(
        spec=UNetSpec(dim=8, u_net_levels=2, timesteps=10, groups=None),
        cells=(16, 8, 8), hole=None, batch=1, seed=505, save_taps=False,
    ),
}

# Cases WITHOUT reference goldens (checked against the oracle only, which the cases above pin to the reference):
# dim=64 with 4 levels doubles every channel count of the shapes config (128 ... 1024, 2048 -> 512 in the first up block),
# i.e. more than 512 output channels per convolution and FiLM rows of 2048.
WIDE_CASES = {
    "dim64": dict(
        spec=UNetSpec(dim=64, u_net_levels=4, timesteps=100, groups=8),
        cells=(22, 6, 6), hole=None, batch=2, seed=606, save_taps=False,
    ),
    # everything "odd" at once: velocity only (F = 3), cell-position features on top of the cell-type embedding (7 local
    # channels, conditioning.py:52-61), an odd batch, and a grid whose z lines (66 voxels with the halo) are longer than
    # what the row-window kernels take, so the bf16 path runs on the kz-folded / per-tap kernels.
    "odd": dict(
        spec=UNetSpec(in_features=3, out_features=3, c_local_features=7, dim=16, u_net_levels=2, timesteps=10, groups=8),
        cells=(10, 6, 64), hole=((2, 5), (1, 4), (10, 30)), batch=3, seed=707, save_taps=False,
    ),
}


def case_inputs(case):
    """(x, t, c_local, geometry) for a case.  x ~ N(0,1) fp32 (B,F,X,Y,Z) on the padded
    grid, t spread over [0,T), c_local = a seeded 6x4 cell-type table looked up on the
    geometry's cell-type map."""
    spec: UNetSpec = case["spec"]
    geo = grid_ref.channel_geometry(cells=case["cells"], hole=case["hole"], seed=case["seed"])
    rng = np.random.Generator(np.random.PCG64(case["seed"] + 1))
    B = case["batch"]
    x = rng.standard_normal((B, spec.in_features, *geo.padded)).astype(np.float32)
    t = np.array([(3 + 5 * b) % spec.timesteps for b in range(B)], dtype=np.int64)
    table = rng.standard_normal((6, spec.c_local_features)).astype(np.float32)
    c_local = np.ascontiguousarray(grid_ref.cell_type_embedding(geo, table))
    return torch.from_numpy(x), torch.from_numpy(t), torch.from_numpy(c_local), geo


# ---- full shapes configuration (BASELINE.json configs[1]/[2]) --------------------------------------------------------
SHAPES_T = 500  # config/model/diffusion.yaml:12 of the reference
SHAPES_SEED = 0          # synth_state_dict seed of the full-size fixtures
SHAPES_INPUT_SEED = 100  # turbdiff_b200.synthetic.synthetic_inputs(1, SHAPES_INPUT_SEED)
SHAPES_FWD_T = 137       # timestep of the full-size forward fixture


def shapes_spec(T: int = SHAPES_T) -> UNetSpec:
    return UNetSpec(in_features=4, out_features=4, c_local_features=4, timesteps=T, dim=32, u_net_levels=4, groups=8)


def sub3(v):
    """Strided sub-sample of the trailing three (spatial) axes used by the full-size fixtures (every 3rd voxel)."""
    return v[..., ::3, ::3, ::3]


def tap_sample(v):
    """Sub-sample of a per-block tap (B, C, X, Y, Z): every 4th channel, every 5th voxel per axis (offset 1)."""
    return v[:, ::4, 1::5, 1::5, 1::5]


def grad_sample(g):
    """<= 1024 evenly strided entries of a flattened gradient tensor."""
    flat = g.reshape(-1)
    return flat[:: max(1, flat.numel() // 1024)][:1024]


# ---- learned variances (ddpm.py:732-741, 853-870) --------------------------------------------------------------------
LV_ELBO_WEIGHT = 0.05
LV_VARIANTS = ((True, True), (False, True), (True, False))  # (noise_bcs, detach_elbo_mean)
LV_SEEDS = (4321, 4)  # both draw t = 0 for one of the two samples: the log-likelihood branch of the ELBO is exercised


def lv_case(base: str = "micro"):
    """A configuration with learned variances: the denoiser predicts 2F channels (diffusion.py:114)."""
    import dataclasses

    case = dict(CASES[base])
    case["spec"] = dataclasses.replace(case["spec"], out_features=2 * case["spec"].in_features)
    return case
