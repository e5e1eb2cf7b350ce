"""Oracle (test infrastructure, CPU): the turbulent-kinetic-energy spectrum statistic of the reference,
``TurbulentKineticEnergySpectrum`` (models/metrics.py:270-320), its log-domain trilinear interpolation ``interp3``
(:211-267) and ``LogTKESpectrumL2Distance`` (:323-378), restated as plain functions.  The sphere quadrature (Lebedev
nodes ``p`` (N,3) and weights ``w`` (N,), taken by the reference from its ``numgrids.pickle``) is an INPUT here.

Pinned by tests/golden/tke.npz (outputs of the unmodified reference classes, tests/golden/make_golden.py gen_tke)."""

from __future__ import annotations

import numpy as np
import torch
from scipy.special import roots_legendre


def interp3(grid: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """Trilinear interpolation of ``grid`` (..., X, Y, Z) at ``points`` (..., 3) given in index coordinates; neighbour
    indices are clamped to the grid but the weights are taken against the CLAMPED lower index, as metrics.py:238-266."""
    shape = torch.tensor(grid.shape[-3:])
    p0 = torch.minimum(torch.clamp(torch.floor(points).long(), min=0), shape - 1)
    p1 = torch.minimum(torch.clamp(torch.floor(points).long() + 1, min=0), shape - 1)
    x0, y0, z0 = p0.unbind(-1)
    x1, y1, z1 = p1.unbind(-1)
    wx, wy, wz = (points - p0).unbind(-1)
    g = grid
    return ((1 - wx) * (1 - wy) * (1 - wz) * g[..., x0, y0, z0] + (1 - wx) * (1 - wy) * wz * g[..., x0, y0, z1]
            + (1 - wx) * wy * (1 - wz) * g[..., x0, y1, z0] + (1 - wx) * wy * wz * g[..., x0, y1, z1]
            + wx * (1 - wy) * (1 - wz) * g[..., x1, y0, z0] + wx * (1 - wy) * wz * g[..., x1, y0, z1]
            + wx * wy * (1 - wz) * g[..., x1, y1, z0] + wx * wy * wz * g[..., x1, y1, z1])


def tke_spectrum(u_perturbation: torch.Tensor, k: torch.Tensor, p: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """E(k) (..., K) of a perturbation velocity field (..., 3, X, Y, Z): TKE = |u'|^2 / 2 per voxel, 3-D FFT shifted to
    the centre, |F|^2 interpolated in the LOG domain onto spheres of radius k around the zero frequency, integrated
    with the quadrature (weights sum to 1) and scaled by the sphere area 4 pi k^2 (metrics.py:296-320)."""
    tke = 0.5 * (u_perturbation**2).sum(dim=-4)
    f = torch.fft.fftshift(torch.fft.fftn(tke, dim=(-3, -2, -1)), dim=(-3, -2, -1))
    center = k.new_tensor([s // 2 for s in u_perturbation.shape[-3:]])
    q = k[:, None, None] * p + center
    vals = interp3((f.abs() ** 2).log(), q).exp().float()  # (..., K, N)
    return torch.matmul(vals, w) * (4 * torch.pi * k**2)


def legendre_k(n_nodes: int, spatial) -> tuple[torch.Tensor, torch.Tensor, float]:
    """Gauss-Legendre nodes mapped to k in [1, (min(spatial) - 1) // 2] (metrics.py:346-362): (k, weights, slope)."""
    nodes, weights = roots_legendre(n_nodes)
    nodes, weights = torch.tensor(nodes).float(), torch.tensor(weights).float()
    k_min, k_max = 1.0, float((min(spatial) - 1) // 2)
    slope = (k_max - k_min) / 2
    return slope * nodes + ((k_max - k_min) / 2 + k_min), weights, slope


def log_tke_l2_distance(u_a, u_b, u_mean, p, w, n_nodes: int = 64):
    """Pairwise L2 distances between the log-TKE spectra of two sets of velocity fields (metrics.py:354-378):
    (D (A,B), log E_a (A,K), log E_b (B,K), k (K,))."""
    k, lw, slope = legendre_k(n_nodes, u_a.shape[-3:])
    la = tke_spectrum(u_a - u_mean, k, p, w).log()
    lb = tke_spectrum(u_b - u_mean, k, p, w).log()
    d = slope * torch.einsum("ijk,k->ij", (la[:, None] - lb[None]) ** 2, lw)
    return torch.sqrt(d), la, lb, k


def synthetic_velocity(batch: int, n: int, seed: int) -> np.ndarray:
    """Deterministic broadband test fields (batch, 3, n, n, n) fp32: white noise shaped to a k^(-5/3)-like spectrum."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.fft.fftfreq(n) * n
    kk = np.sqrt(f[:, None, None] ** 2 + f[None, :, None] ** 2 + f[None, None, :] ** 2)
    shape = (1.0 + kk) ** (-11.0 / 6.0)
    u = rng.standard_normal((batch, 3, n, n, n))
    u = np.fft.ifftn(np.fft.fftn(u, axes=(-3, -2, -1)) * shape, axes=(-3, -2, -1)).real
    u = u / u.std() + np.array([1.0, 0.1, -0.2])[None, :, None, None, None]
    return u.astype(np.float32)
