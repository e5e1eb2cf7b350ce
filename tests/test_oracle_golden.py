"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is the parity pin of the oracle."""

import json

import numpy as np
import pytest
import torch

from oracle import grid_ref
from oracle.cases import CASES, case_inputs
from oracle.diffusion_ref import BUFFER_NAMES, SCHEDULES, DiffusionRef, diffusion_buffers
from oracle.unet_ref import UNetSpec, denoiser_forward, state_dict_layout, synth_state_dict, time_embedding, trilinear_resample

from conftest import GOLDEN


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_state_dict_layout_matches_reference():
    ref = json.loads((GOLDEN / "state_dict_layout_shapes.json").read_text())
    mine = [[k, list(s)] for k, s in state_dict_layout(UNetSpec())]
    assert mine == ref
    assert len(ref) == 139
    assert sum(int(np.prod(s)) for _, s in ref) == 55_246_788


@pytest.mark.parametrize("name", SCHEDULES)
@pytest.mark.parametrize("T", [10, 500, 1000])
def test_schedule_buffers_bit_exact(golden, name, T):
    g = golden["schedules"]
    buf = diffusion_buffers(name, T)
    for b in BUFFER_NAMES:
        want = g[f"{name}/{T}/{b}"]
        got = buf[b].numpy()
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, want, err_msg=f"{name}/{T}/{b}")


def test_schedule_known_answers():
    # SURVEY.md section 8c known-answer values (log-snr-linear, T=500)
    b = diffusion_buffers("log-snr-linear", 500)
    np.testing.assert_allclose(b["betas"][[0, 1, 250, 499]].numpy(),
                               [9.990009712e-04, 3.756604201e-05, 3.300226107e-02, 3.624184802e-02], rtol=1e-6)
    np.testing.assert_allclose(b["posterior_log_var"][[0, 1, 250, 499]].numpy(),
                               [-6.933759212, -10.22628784, -3.414535046, -3.317541361], rtol=1e-6)
    np.testing.assert_allclose(float(b["sqrt_recip_alphas_cumprod"][-1]), 3.162293701e02, rtol=1e-6)


@pytest.mark.parametrize("dim,T", [(32, 500), (32, 1000), (16, 10), (8, 10)])
def test_time_embedding(golden, dim, T):
    g = golden["time_embedding"]
    got = time_embedding(torch.arange(T), dim, T).numpy()
    np.testing.assert_allclose(got, g[f"{dim}/{T}"], atol=2e-6, rtol=0)


def test_grid_helpers_bit_exact(golden):
    g = golden["grid"]
    geo = grid_ref.channel_geometry(cells=(12, 6, 5), hole=((3, 6), (1, 4), (0, 3)), seed=3)
    rng = np.random.Generator(np.random.PCG64(11))
    B, n = 2, len(geo.cell_idx)
    u = rng.standard_normal((B, n, 3)).astype(np.float32)
    p = rng.standard_normal((B, n, 1)).astype(np.float32)
    grid = grid_ref.grid_embedding(
        geo, [u, p],
        [{"inlets": [20.0, 0.0, 0.0], "walls": [0.0, 0.0, 0.0]}, {"outlets": [0.0]}],
    )
    np.testing.assert_array_equal(grid, g["grid_embedding"])
    np.testing.assert_array_equal(grid_ref.cell_type_map(geo), g["cell_types"])
    table = rng.standard_normal((6, 4)).astype(np.float32)
    np.testing.assert_array_equal(table, g["table"])
    np.testing.assert_array_equal(grid_ref.cell_type_embedding(geo, table), g["cell_type_embedding"])
    other = g["other"]
    np.testing.assert_array_equal(grid_ref.where_cells(geo.cell_idx, grid, other), g["where_cells"])
    np.testing.assert_array_equal(grid_ref.where_cells(geo.cell_idx, other), g["where_cells_zero"])
    np.testing.assert_array_equal(grid_ref.select_cells(other, geo.cell_idx), g["select_cells"])
    # scatter -> gather round trip returns the samples in cell_idx order
    np.testing.assert_array_equal(np.swapaxes(grid_ref.select_cells(grid[:, :3], geo.cell_idx), 1, 2), u)
    # mask select == where_cells (order independence)
    m = grid_ref.inside_mask(geo).reshape(geo.padded).astype(bool)
    np.testing.assert_array_equal(np.where(m, grid, other), g["where_cells"])


def test_trilinear_matches_torch():
    x = torch.randn(2, 3, 11, 7, 5, generator=torch.Generator().manual_seed(0))
    for size in [(5, 3, 3), (22, 14, 10), (11, 7, 5), (3, 3, 3)]:
        want = torch.nn.functional.interpolate(x, size=size, mode="trilinear", align_corners=True)
        assert rel_l2(trilinear_resample(x, size), want) < 1e-6


@pytest.mark.parametrize("cname", list(CASES))
def test_denoiser_forward(golden, cname):
    g = golden["unet"]
    case = CASES[cname]
    spec = case["spec"]
    sd = synth_state_dict(spec, case["seed"])
    x, t, c_local, _ = case_inputs(case)
    taps = {}
    with torch.no_grad():
        y = denoiser_forward(sd, spec, x, t, c_local, taps)
    assert rel_l2(y, g[f"{cname}/out"]) < 1e-5
    for k in [k for k in g.files if k.startswith(f"{cname}/tap/")]:
        assert rel_l2(taps[k.split("/")[-1]], g[k]) < 1e-5, k
    for k in [k for k in g.files if k.startswith(f"{cname}/tapsum/")]:
        v = taps[k.split("/")[-1]].double()
        np.testing.assert_allclose([v.sum().item(), v.pow(2).sum().item()], g[k], rtol=2e-4, atol=1e-3)


def test_denoiser_forward_fp64_close_to_fp32(golden):
    case = CASES["micro"]
    spec = case["spec"]
    sd = synth_state_dict(spec, case["seed"], dtype=torch.float64)
    x, t, c_local, _ = case_inputs(case)
    with torch.no_grad():
        y = denoiser_forward(sd, spec, x.double(), t, c_local.double())
    assert rel_l2(y, golden["unet"]["micro/out"]) < 1e-5


def _diffusion(cname, noise_bcs, loss_type="l2"):
    case = CASES[cname]
    spec = case["spec"]
    sd = {k: v.requires_grad_() for k, v in synth_state_dict(spec, case["seed"]).items()}
    x, t, c_local, geo = case_inputs(case)
    eps = lambda x_t, tt: denoiser_forward(sd, spec, x_t, tt, c_local)
    d = DiffusionRef(eps, timesteps=spec.timesteps, beta_schedule="log-snr-linear", loss_type=loss_type, noise_bcs=noise_bcs)
    return d, sd, x, torch.from_numpy(geo.cell_idx)


@pytest.mark.parametrize("cname", ["micro", "tiny"])
@pytest.mark.parametrize("noise_bcs", [True, False])
def test_sampling_loop(golden, cname, noise_bcs):
    g = golden["diffusion"]
    tag = f"{cname}/noise_bcs={int(noise_bcs)}"
    d, _, x, idx = _diffusion(cname, noise_bcs)
    torch.manual_seed(1234)
    assert rel_l2(d.sample_loop(x, idx), g[f"{tag}/sample"]) < 1e-5
    torch.manual_seed(1234)
    assert rel_l2(d.sample_loop(x, idx, start_from=4), g[f"{tag}/sample_from4"]) < 1e-5
    for tt in (3, 0):
        with torch.no_grad():
            _, _, mean, lv = d.predictions(x, torch.full((x.shape[0],), tt), idx)
        assert rel_l2(mean, g[f"{tag}/p_sample_mean/{tt}"]) < 1e-5
        np.testing.assert_array_equal(lv.numpy(), g[f"{tag}/p_sample_logvar/{tt}"])


@pytest.mark.parametrize("cname", ["micro", "tiny"])
@pytest.mark.parametrize("noise_bcs", [True, False])
def test_training_loss_and_grads(golden, cname, noise_bcs):
    g = golden["diffusion"]
    tag = f"{cname}/noise_bcs={int(noise_bcs)}"
    d, sd, x, idx = _diffusion(cname, noise_bcs)
    torch.manual_seed(4321)
    loss, t = d.forward(x, idx)
    np.testing.assert_array_equal(t.numpy(), g[f"{tag}/t"])
    np.testing.assert_allclose(loss.item(), g[f"{tag}/loss"], rtol=1e-5)
    loss.backward()
    for k, p in sd.items():
        gs = g[f"{tag}/gradsum/{k}"]
        got = [p.grad.double().sum().item(), p.grad.double().pow(2).sum().item()]
        np.testing.assert_allclose(got[1], gs[1], rtol=1e-3, atol=1e-10, err_msg=k)
        key = f"{tag}/grad/{k}"
        if key in g.files:
            # a conv bias in front of a one-channel-per-group norm has an exactly zero gradient
            assert rel_l2(p.grad, g[key]) < 1e-4 or np.abs(g[key]).max() < 1e-7, k
    if cname == "micro":
        d1, _, x, idx = _diffusion(cname, noise_bcs, "l1")
        torch.manual_seed(4321)
        np.testing.assert_allclose(d1.forward(x, idx)[0].item(), g[f"{tag}/loss_l1"], rtol=1e-5)


# ---- the FULL shapes configuration (194x50x50, 4 levels, 55.2 M parameters): oracle vs the reference's golden ---------


@pytest.fixture(scope="module")
def shapes_case():
    from oracle.cases import SHAPES_INPUT_SEED, SHAPES_SEED, shapes_spec
    from turbdiff_b200.synthetic import synthetic_inputs

    spec = shapes_spec()
    geo, x, c_local = synthetic_inputs(1, SHAPES_INPUT_SEED)
    return spec, synth_state_dict(spec, SHAPES_SEED), geo, x, c_local


def test_full_size_denoiser_forward(golden, shapes_case):
    from oracle.cases import SHAPES_FWD_T, sub3, tap_sample

    g = golden["shapes"]
    spec, sd, geo, x, c_local = shapes_case
    taps = {}
    with torch.no_grad():
        y = denoiser_forward(sd, spec, x, torch.tensor([SHAPES_FWD_T]), c_local, taps)
    assert y.shape == (1, 4, 194, 50, 50)
    assert rel_l2(sub3(y), g["out/sub"]) < 1e-5
    np.testing.assert_allclose(y.double().pow(2).sum().item(), g["out/sum"][1], rtol=1e-4)
    names = [k.split("/")[1] for k in g.files if k.startswith("tap/") and k.endswith("/sub")]
    assert len(names) == 12
    for name in names:
        assert rel_l2(tap_sample(taps[name]), g[f"tap/{name}/sub"]) < 1e-5, name
        np.testing.assert_allclose(taps[name].double().pow(2).sum().item(), g[f"tap/{name}/sum"][1], rtol=1e-4, err_msg=name)


def test_full_size_sampling_and_training_step(golden, shapes_case):
    from oracle.cases import SHAPES_T, grad_sample, sub3

    g = golden["shapes"]
    spec, sd, geo, x, c_local = shapes_case
    sd = {k: v.clone().requires_grad_() for k, v in sd.items()}
    idx = torch.from_numpy(geo.cell_idx)
    d = DiffusionRef(lambda x_t, tt: denoiser_forward(sd, spec, x_t, tt, c_local), timesteps=SHAPES_T, beta_schedule="log-snr-linear",
                     loss_type="l2", noise_bcs=True)
    torch.manual_seed(77)
    with torch.no_grad():
        s = d.sample_loop(x, idx, start_from=3)
    assert rel_l2(sub3(s), g["sample_from3/sub"]) < 1e-5
    torch.manual_seed(4321)
    loss, t = d.forward(x, idx)
    np.testing.assert_array_equal(t.numpy(), g["loss_t"])
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)
    loss.backward()
    for k, p in sd.items():
        want = g[f"grad/{k}/sub"]
        if np.abs(want).max() < 1e-7:  # conv bias in front of GroupNorm: exactly zero up to rounding noise
            continue
        assert rel_l2(grad_sample(p.grad), want) < 2e-4, k


# ---- learned variances + ELBO (ddpm.py:732-741, 853-870) -------------------------------------------------------------


def _lv_diffusion(noise_bcs, detach):
    from oracle.cases import LV_ELBO_WEIGHT, lv_case

    case = lv_case()
    spec = case["spec"]
    sd = {k: v.requires_grad_() for k, v in synth_state_dict(spec, case["seed"]).items()}
    x, t, c_local, geo = case_inputs(case)
    d = DiffusionRef(lambda x_t, tt: denoiser_forward(sd, spec, x_t, tt, c_local), timesteps=spec.timesteps, beta_schedule="log-snr-linear",
                     loss_type="l2", noise_bcs=noise_bcs, learned_variances=True, elbo_weight=LV_ELBO_WEIGHT, detach_elbo_mean=detach)
    return d, sd, x, torch.from_numpy(geo.cell_idx)


@pytest.mark.parametrize("noise_bcs,detach", [(True, True), (False, True), (True, False)])
def test_learned_variances_loss_and_grads(golden, noise_bcs, detach):
    from oracle.cases import LV_SEEDS

    g = golden["diffusion_lv"]
    tag = f"noise_bcs={int(noise_bcs)}/detach={int(detach)}"
    d, sd, x, idx = _lv_diffusion(noise_bcs, detach)
    if detach:
        assert "only one dimension" in str(g[f"{tag}/sample_loop_error"])  # the reference cannot sample this variant
        for tt in (3, 0):
            with torch.no_grad():
                _, _, mean, lv = d.predictions(x, torch.full((x.shape[0],), tt), idx)
            assert rel_l2(mean, g[f"{tag}/p_sample_mean/{tt}"]) < 1e-5
            assert rel_l2(lv, g[f"{tag}/p_sample_logvar/{tt}"]) < 1e-6
    for seed in LV_SEEDS:
        for p in sd.values():
            p.grad = None
        torch.manual_seed(seed)
        loss, t = d.forward(x, idx)
        np.testing.assert_array_equal(t.numpy(), g[f"{tag}/t/{seed}"])
        assert 0 in t.tolist()
        np.testing.assert_allclose(loss.item(), g[f"{tag}/loss/{seed}"], rtol=1e-5)
        loss.backward()
        for k, p in sd.items():
            key = f"{tag}/grad/{seed}/{k}"
            if key in g.files and np.abs(g[key]).max() > 1e-7:
                assert rel_l2(p.grad, g[key]) < 2e-4, k


# ---- TKE-spectrum statistic (models/metrics.py:270-378) --------------------------------------------------------------


@pytest.mark.parametrize("n", [16, 24])
def test_tke_spectrum_and_distance(golden, n):
    from oracle import tke_ref

    g = golden["tke"]
    p, w = torch.from_numpy(g["p110"]), torch.from_numpy(g["w110"])
    u = torch.from_numpy(tke_ref.synthetic_velocity(3, n, {16: 1, 24: 2}[n]))
    um = u.mean(0)
    D, la, lb, k = tke_ref.log_tke_l2_distance(u[:2], u[1:], um, p, w, 16)
    np.testing.assert_array_equal(k.numpy(), g[f"synthetic/{n}/k"])
    assert rel_l2(tke_ref.tke_spectrum(u - um, k, p, w), g[f"synthetic/{n}/E"]) < 1e-6
    assert rel_l2(la, g[f"synthetic/{n}/log_a"]) < 1e-6 and rel_l2(lb, g[f"synthetic/{n}/log_b"]) < 1e-6
    np.testing.assert_allclose(D.numpy(), g[f"synthetic/{n}/D"], rtol=1e-4, atol=1e-5)


def test_tke_spectrum_production_size(golden):
    """One 48^3 cube, 64 Gauss-Legendre radii, 5810 Lebedev points (the reference's defaults); the quadrature comes from
    the reference package (its numgrids.pickle), so this needs /root/reference or the oracle/_ref install."""
    from oracle import ref_shim, tke_ref

    if not ref_shim.available():
        pytest.skip("reference package not installed")
    import pickle

    x, y, z, w = pickle.loads((ref_shim.reference_root() / "turbdiff" / "models" / "numgrids.pickle").read_bytes())[5810]
    p, w = torch.tensor([x, y, z]).T.float(), torch.tensor(w).float()
    g = golden["tke"]
    u = torch.from_numpy(tke_ref.synthetic_velocity(2, 48, 3))
    D, la, lb, k = tke_ref.log_tke_l2_distance(u[:1], u[1:], u.mean(0), p, w, 64)
    assert rel_l2(la, g["synthetic/48/log_a"]) < 1e-6 and rel_l2(lb, g["synthetic/48/log_b"]) < 1e-6
    np.testing.assert_allclose(D.numpy(), g["synthetic/48/D"], rtol=1e-4)
