"""TKE-spectrum statistic on the GPU (tdb_tke_spectrum, turbdiff_b200.metrics) against the unmodified reference classes
(tests/golden/tke.npz), and the statistical-parity check north_star asks for: the energy spectra of samples drawn by this
repo's sampler agree with those of the reference's sampler (same weights, same noise seeds)."""

import numpy as np
import pytest
import torch

from util import cpu_seeded_randn, rel_l2

pytestmark = pytest.mark.gpu


def _modules(golden, n_nodes):
    from turbdiff_b200.metrics import LogTKESpectrumL2Distance, TurbulentKineticEnergySpectrum

    g = golden["tke"]
    spec = TurbulentKineticEnergySpectrum(points=torch.from_numpy(g["p110"]), weights=torch.from_numpy(g["w110"])).cuda()
    return spec, LogTKESpectrumL2Distance(spec, n=n_nodes).cuda()


@pytest.mark.parametrize("n", [16, 24])
def test_tke_spectrum_matches_reference_golden(golden, n):
    from oracle import tke_ref

    g = golden["tke"]
    spec, dist = _modules(golden, 16)
    u = torch.from_numpy(tke_ref.synthetic_velocity(3, n, {16: 1, 24: 2}[n])).cuda()
    um = u.mean(0)
    k = torch.from_numpy(g[f"synthetic/{n}/k"]).cuda()
    E = spec(u - um, k)
    assert rel_l2(E, g[f"synthetic/{n}/E"]) < 2e-5
    assert torch.equal(spec(u, k, u_mean=um), E)  # fused mean subtraction = the reference's fp32 subtraction
    D, la, lb, kk = dist(u[:2], u[1:], um)
    np.testing.assert_allclose(kk.cpu().numpy(), g[f"synthetic/{n}/k"], rtol=1e-6)
    assert rel_l2(la, g[f"synthetic/{n}/log_a"]) < 1e-5 and rel_l2(lb, g[f"synthetic/{n}/log_b"]) < 1e-5
    np.testing.assert_allclose(D.cpu().numpy(), g[f"synthetic/{n}/D"], rtol=2e-3, atol=2e-4)


def test_tke_spectrum_production_size(golden):
    """48^3 cube, 64 radii, the reference's 5810-point Lebedev grid (read from the installed reference package)."""
    from oracle import ref_shim, tke_ref
    from turbdiff_b200.metrics import LogTKESpectrumL2Distance, TurbulentKineticEnergySpectrum

    if not ref_shim.available():
        pytest.skip("reference package not installed")
    ref_shim.load()  # puts the reference on sys.path: the quadrature table is read from it
    g = golden["tke"]
    dist = LogTKESpectrumL2Distance(TurbulentKineticEnergySpectrum(), n=64).cuda()
    assert dist.tke_spectrum.n == 5810
    u = torch.from_numpy(tke_ref.synthetic_velocity(2, 48, 3)).cuda()
    D, la, lb, k = dist(u[:1], u[1:], u.mean(0))
    assert rel_l2(la, g["synthetic/48/log_a"]) < 1e-5 and rel_l2(lb, g["synthetic/48/log_b"]) < 1e-5
    # (two fields and their own mean: the perturbations are each other's negatives, the spectra coincide and D is pure
    # rounding noise - 8e-7 in the reference, 2e-6 here)
    np.testing.assert_allclose(D.cpu().numpy(), g["synthetic/48/D"], atol=2e-5)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_sample_energy_spectra_agree_with_the_reference_sampler(golden, precision, tol):
    """16 ancestral-sampling chains of the tiny configuration (8 noise seeds x batch 2) through GaussianDiffusion.p_sample_loop
    of this repo; their log-TKE spectra on the two 16^3 cubes of the channel against the spectra of the reference's chains
    (same weights, same noise): per-sample agreement, and the spectrum distance between matching samples is far below the
    spread of the ensemble (the statistic WassersteinTKE is built on, metrics.py:381-476)."""
    from oracle.cases import CASES, case_inputs
    from test_gpu_model import build, key_of
    from turbdiff_b200 import GaussianDiffusion

    g = golden["tke"]
    case = CASES["tiny"]
    m = build(case, precision)
    gd = GaussianDiffusion(m, timesteps=case["spec"].timesteps, beta_schedule="log-snr-linear", loss_type="l2", noise_bcs=True).cuda()
    x, _, c_local, geo = case_inputs(case)
    C = {key_of(): c_local.cuda()}
    idx = torch.from_numpy(geo.cell_idx).cuda()
    chains = []
    for seed in g["stat/seeds"].tolist():
        with cpu_seeded_randn(seed):
            chains.append(gd.p_sample_loop(x.cuda(), C, idx))
    s = torch.cat(chains)
    u = s[:, :3, 1:-1, 1:-1, 1:-1]
    cubes = torch.stack((u[..., :16, :, :], u[..., 16:, :, :]), dim=1)  # (16, 2, 3, 16, 16, 16)
    u_mean = torch.from_numpy(g["stat/u_mean"]).cuda()
    _, dist = _modules(golden, 16)
    want = torch.from_numpy(g["stat/log_tke"]).cuda()
    for c in range(2):
        D, la, _, _ = dist(cubes[:, c].contiguous(), cubes[:, c].contiguous(), u_mean[c])
        err = rel_l2(la, want[:, c])
        # distance of every sample of ours to the reference ensemble: matching samples vs the ensemble spread
        Dx = torch.sqrt(dist.legendre_weights.cuda().new_tensor(0.0) + ((la[:, None] - want[None, :, c]) ** 2 @ dist.legendre_weights.cuda())
                        * ((float((16 - 1) // 2) - 1.0) / 2))
        ref_D = torch.from_numpy(g[f"stat/D/{c}"])
        spread = float(np.median(ref_D.numpy()[~np.eye(16, dtype=bool)]))
        print(f"cube {c} {precision}: log-spectrum rel-L2 {err:.2e}, max matched distance {float(Dx.diagonal().max()):.3e}, ensemble spread {spread:.3f}")
        assert err < tol
        assert float(Dx.diagonal().max()) < (0.01 if precision == "fp32" else 0.25) * spread
        # ensemble statistics: mean log-spectrum of the 16 samples
        assert rel_l2(la.mean(0), want[:, c].mean(0)) < tol
