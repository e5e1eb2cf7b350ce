"""GPU implementation of the reference's turbulent-kinetic-energy spectrum statistic (turbdiff/models/metrics.py:270-378):
``TurbulentKineticEnergySpectrum`` and ``LogTKESpectrumL2Distance`` with the reference's constructor / forward signatures.
It is the acceptance statistic of the sampler ("agreement of sample TKE / energy-spectrum statistics") and runs after every
sampling pass (WassersteinTKE, metrics.py:381-476), per cube region of the channel.

The sphere quadrature (Lebedev nodes) is data of the reference package (``turbdiff/models/numgrids.pickle``): it is read
from the installed reference when ``points`` / ``weights`` are not given."""

from __future__ import annotations

import pickle
from pathlib import Path

import torch
from scipy.special import roots_legendre
from torch import nn

from . import _lib
from ._lib import call, ptr


def reference_quadrature(n: int):
    """(points (n,3), weights (n,)) of the reference's Lebedev grid with n nodes (metrics.py:283-291)."""
    import importlib.util

    spec = importlib.util.find_spec("turbdiff")
    if spec is None or not spec.submodule_search_locations:
        raise RuntimeError("turbdiff_b200.metrics: the Lebedev quadrature is read from the reference package "
                           "(turbdiff/models/numgrids.pickle), which is not importable; pass points= and weights= instead")
    grids = pickle.loads((Path(list(spec.submodule_search_locations)[0]) / "models" / "numgrids.pickle").read_bytes())
    if n not in grids:
        raise RuntimeError(f"n={n} is not supported by numgrid.")
    x, y, z, w = grids[n]
    return torch.tensor([x, y, z]).T.float(), torch.tensor(w).float()


class TurbulentKineticEnergySpectrum(nn.Module):
    """Estimate the turbulent kinetic energy spectrum of a 3D flow field (metrics.py:270-320)."""

    def __init__(self, n: int = 5810, points: torch.Tensor | None = None, weights: torch.Tensor | None = None):
        super().__init__()
        if points is None or weights is None:
            points, weights = reference_quadrature(n)
        self.n = int(points.shape[0])
        self.register_buffer("p", points.float().contiguous())
        self.register_buffer("w", weights.float().contiguous())

    def forward(self, u_perturbation: torch.Tensor, k: torch.Tensor, u_mean: torch.Tensor | None = None):
        """E(k) with shape (..., K) for u_perturbation (..., 3, X, Y, Z); with ``u_mean`` (3, X, Y, Z) the subtraction
        u - u_mean is fused into the kernel (pass the raw velocity as the first argument then)."""
        _lib.require_cuda(u_perturbation, "u_perturbation")
        assert u_perturbation.shape[-4] == 3
        lead = u_perturbation.shape[:-4]
        n0, n1, n2 = u_perturbation.shape[-3:]
        u = u_perturbation.reshape(-1, 3, n0, n1, n2).to(torch.float32).contiguous()
        dev = u.device
        kk = k.to(device=dev, dtype=torch.float32).contiguous()
        um = None if u_mean is None else u_mean.to(device=dev, dtype=torch.float32).expand(3, n0, n1, n2).contiguous()
        B, K = u.shape[0], kk.numel()
        work = torch.empty(4 * B * n0 * n1 * n2, dtype=torch.float32, device=dev)
        E = torch.empty((B, K), dtype=torch.float32, device=dev)
        p, w = self.p.to(dev), self.w.to(dev)
        call("tdb_tke_spectrum", u.data_ptr(), ptr(um), B, n0, n1, n2, kk.data_ptr(), K, p.data_ptr(), w.data_ptr(), self.n,
             work.data_ptr(), E.data_ptr(), _lib.stream_ptr())
        return E.reshape(*lead, K)


class LogTKESpectrumL2Distance(nn.Module):
    """L2 distance between the log-TKE spectrum functions E(k) of two sets of flows with Gauss-Legendre integration
    (metrics.py:323-378)."""

    def __init__(self, tke_spectrum: nn.Module, n: int = 64):
        super().__init__()
        self.tke_spectrum = tke_spectrum
        self.n = n
        nodes, weights = roots_legendre(n)
        self.register_buffer("legendre_nodes", torch.tensor(nodes).float())
        self.register_buffer("legendre_weights", torch.tensor(weights).float())

    def forward(self, u_a: torch.Tensor, u_b: torch.Tensor, u_mean: torch.Tensor):
        assert u_a.shape[-4] == 3 and u_b.shape[-4] == 3 and u_mean.shape[-4] == 3
        assert u_a.shape[-3:] == u_b.shape[-3:] and u_a.shape[-3:] == u_mean.shape[-3:]
        k_min = 1.0
        k_max = float((min(u_a.shape[-3:]) - 1) // 2)
        slope = (k_max - k_min) / 2
        k = (slope * self.legendre_nodes + ((k_max - k_min) / 2 + k_min)).to(u_a.device)
        log_tke_a = self.tke_spectrum(u_a, k, u_mean=u_mean).log()
        log_tke_b = self.tke_spectrum(u_b, k, u_mean=u_mean).log()
        D = slope * torch.einsum("ijk, k -> ij", (log_tke_a[:, None] - log_tke_b[None]) ** 2, self.legendre_weights.to(u_a.device))
        return torch.sqrt(D), log_tke_a, log_tke_b, k
