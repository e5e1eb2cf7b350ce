"""Parity at BASELINE.json's FULL shapes configuration (194x50x50 padded grid, u+p, dim 32, 4 levels, 55.2 M
parameters).

First against the UNMODIFIED REFERENCE: tests/golden/shapes.npz holds the reference's full-size denoiser output, its
twelve per-block taps, a 3-step sampling chain and one training loss with all 139 parameter gradients (strided
sub-samples + checksums, tests/golden/make_golden.py gen_shapes); fp32 path <= 1e-5 per block, bf16 path <= 2e-2
(BASELINE.json north_star).  Then size-independent properties:

* the bf16 tensor-core path against the fp32 parity path of the same library (the fp32 path is pinned to the
  reference's golden vectors at small sizes; tolerance = BASELINE north_star's 2e-2);
* batch-order equivariance (samples are independent: GroupNorm is per sample);
* the sampling loop leaves every non-interior cell at its boundary value bit-exactly, and is reproducible;
* gradients: bf16 against fp32 (cosine), a central finite difference of the loss along the gradient direction (fp32 path),
  and the closed form of the decoder-bias gradient.
"""

import numpy as np
import pytest
import torch

from util import rel_l2

pytestmark = pytest.mark.gpu

T = 1000


@pytest.fixture(scope="module")
def full():
    import bench
    from oracle.unet_ref import synth_state_dict
    from turbdiff_b200 import DenoisingModel

    spec = bench.shapes_spec(T)
    sd = synth_state_dict(spec, 0)
    geo, x, c_local = bench.synthetic_inputs(2, 100)

    def make(precision):
        m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=T, dim=32, u_net_levels=4,
                           norm_type="group", precision=precision)
        m.load_state_dict(sd, strict=True)
        return m.cuda().eval()

    return {"geo": geo, "x": x.cuda(), "c_local": c_local.cuda(), "make": make}


def _key():
    from turbdiff_b200.models.conditioning import Conditioning

    return Conditioning.Type.CELL_TYPE


def test_full_size_bf16_path_matches_fp32_path(full):
    t = torch.tensor([17, 803], dtype=torch.long, device="cuda")
    with torch.no_grad():
        ref = full["make"]("fp32")(full["x"], t, {_key(): full["c_local"]}).clone()
        got = full["make"]("bf16")(full["x"], t, {_key(): full["c_local"]}).clone()
    assert ref.shape == (2, 4, 194, 50, 50)
    assert torch.isfinite(got).all()
    err = rel_l2(got, ref)
    print("full-size bf16 vs fp32 rel-L2:", err)
    assert err < 2e-2  # BASELINE.json north_star: 2e-2 for the bf16 path


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_full_size_batch_order_equivariance(full, precision):
    m = full["make"](precision)
    t = torch.tensor([5, 640], dtype=torch.long, device="cuda")
    C = {_key(): full["c_local"]}
    with torch.no_grad():
        a = m(full["x"], t, C).clone()
        b = m(full["x"].flip(0).contiguous(), t.flip(0).contiguous(), C).clone()
    # fp32 path: the only order-dependent arithmetic is the atomic accumulation of GroupNorm moments in double.
    # bf16 path: the moments come from fp32 partial sums whose grouping follows the tile -> CTA assignment, and a
    # 1e-7 change of a statistic flips bf16 roundings downstream: equivariance holds to the bf16 noise floor
    # (measured 6e-3, the same size as the bf16-vs-fp32 error), so the bound is the path's tolerance
    err = rel_l2(b.flip(0), a)
    print("full-size batch-order equivariance", precision, err)
    assert err < (1e-5 if precision == "fp32" else 2e-2)


def test_full_size_sampling_keeps_boundary_cells_and_is_reproducible(full):
    from turbdiff_b200 import GaussianDiffusion

    geo = full["geo"]
    m = full["make"]("bf16")
    gd = GaussianDiffusion(m, timesteps=T, beta_schedule="log-snr-linear", noise_bcs=True).cuda()
    idx = torch.from_numpy(geo.cell_idx).cuda()
    C = {_key(): full["c_local"]}
    outs = []
    for _ in range(2):
        torch.manual_seed(7)
        with torch.no_grad():
            outs.append(gd.p_sample_loop(full["x"], C, idx, start_from=3).clone())
    s = outs[0]
    assert s.shape == full["x"].shape and torch.isfinite(s).all()
    from turbdiff_b200.models.utils import inside_mask

    nvox = int(np.prod(geo.padded))
    inside = inside_mask(idx, nvox).bool()
    outside = ~inside.view(1, 1, *geo.padded).expand_as(s)
    assert torch.equal(s[outside], full["x"][outside])          # boundary / padding cells: bit-exact x_bcs (ddpm.py:812-814)
    assert not torch.equal(s[~outside], full["x"][~outside])    # the interior was actually sampled
    assert rel_l2(outs[1], outs[0]) < 5e-3                      # same seed, same chain (up to atomic summation order; measured 1e-4)


def _loss_and_grads(full, precision, direction=None, eps=0.0):
    m = full["make"](precision).train()
    if direction is not None:
        with torch.no_grad():
            for (_, p), d in zip(m.named_parameters(), direction):
                p.add_(eps * d)
    t = torch.tensor([42, 911], dtype=torch.long, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    G = torch.randn(full["x"].shape, generator=g, device="cuda") / full["x"].numel() ** 0.5
    if direction is not None:  # plain evaluation of the loss at the shifted parameters
        with torch.no_grad():
            out = m(full["x"], t, {_key(): full["c_local"]})
            return float((out.double() * G.double()).sum()), None, G
    cl = full["c_local"].clone().requires_grad_()
    out = m(full["x"], t, {_key(): cl})
    loss = (out.double() * G.double()).sum()
    loss.backward()
    return float(loss.detach()), {k: p.grad.clone() for k, p in m.named_parameters()}, G


def test_full_size_gradients(full):
    l32, g32, G = _loss_and_grads(full, "fp32")
    _, g16, _ = _loss_and_grads(full, "bf16")
    # closed form: d loss / d decode.1.bias[f] = sum of G over samples and voxels of feature f
    assert rel_l2(g32["decode.1.bias"], G.sum(dim=(0, 2, 3, 4))) < 1e-4
    # bf16 gradients against fp32 gradients: direction of the whole gradient and of the big weight tensors
    a = torch.cat([v.flatten().double() for v in g16.values()])
    b = torch.cat([g32[k].flatten().double() for k in g16])
    cos = float((a @ b) / (a.norm() * b.norm()))
    print("full-size gradient cosine bf16 vs fp32:", cos, "rel-L2:", float((a - b).norm() / b.norm()))
    assert cos > 0.995
    # central finite difference of the fp32 loss along the (normalised) gradient direction: slope = |g|
    gnorm = float(b.norm())
    direction = [g32[k] / gnorm for k in g32]
    h = 2e-4  # |g| ~ 1e2: the loss moves by ~3e-2, far above fp32 noise and well inside the linear regime
    lp, _, _ = _loss_and_grads(full, "fp32", direction, +h)
    lm, _, _ = _loss_and_grads(full, "fp32", direction, -h)
    fd = (lp - lm) / (2 * h)
    print("full-size directional derivative: analytic", gnorm, "finite difference", fd, "loss", l32)
    assert abs(fd - gnorm) <= 5e-2 * gnorm


# ---- against the unmodified reference at full size (tests/golden/shapes.npz) ------------------------------------------


@pytest.fixture(scope="module")
def shapes():
    from oracle.cases import SHAPES_INPUT_SEED, SHAPES_SEED, SHAPES_T, shapes_spec
    from oracle.unet_ref import synth_state_dict
    from turbdiff_b200 import DenoisingModel, GaussianDiffusion
    from turbdiff_b200.synthetic import synthetic_inputs

    sd = synth_state_dict(shapes_spec(), SHAPES_SEED)
    geo, x, c_local = synthetic_inputs(1, SHAPES_INPUT_SEED)

    def make(precision):
        m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=SHAPES_T, dim=32,
                           u_net_levels=4, norm_type="group", precision=precision)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        gd = GaussianDiffusion(m, timesteps=SHAPES_T, beta_schedule="log-snr-linear", loss_type="l2", noise_bcs=True).cuda()
        return m, gd

    return {"geo": geo, "x": x.cuda(), "c_local": c_local.cuda(), "idx": torch.from_numpy(geo.cell_idx).cuda(), "make": make}


TOL = {"fp32": 1e-5, "bf16": 2e-2}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_forward_matches_reference_golden(golden, shapes, precision):
    """DenoisingModel.forward (ddpm.py:477-505) at 194x50x50 / 4 levels / 512-1024 channels: output and every block."""
    from oracle.cases import SHAPES_FWD_T, sub3, tap_sample

    g = golden["shapes"]
    m, _ = shapes["make"](precision)
    t = torch.tensor([SHAPES_FWD_T], dtype=torch.long, device="cuda")
    taps = {}
    with torch.no_grad():
        eps = m.engine().forward(shapes["x"], t, shapes["c_local"], taps=taps).clone()
    errs = {"out": rel_l2(sub3(eps), g["out/sub"])}
    for name, v in taps.items():
        errs[name] = rel_l2(tap_sample(v), g[f"tap/{name}/sub"])
        np.testing.assert_allclose(v.double().pow(2).sum().item(), g[f"tap/{name}/sum"][1], rtol=1e-4 if precision == "fp32" else 2e-2,
                                   err_msg=name)
    print("full-size vs reference", precision, {k: f"{v:.2e}" for k, v in errs.items()})
    assert len(taps) == 12
    assert max(errs.values()) < TOL[precision], errs


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_sampling_chain_matches_reference_golden(golden, shapes, precision):
    """p_sample_loop(start_from=3) (ddpm.py:767-816) on the reference's noise stream."""
    from oracle.cases import sub3
    from util import cpu_seeded_randn

    g = golden["shapes"]
    _, gd = shapes["make"](precision)
    C = {_key(): shapes["c_local"]}
    with cpu_seeded_randn(77):
        s = gd.p_sample_loop(shapes["x"], C, shapes["idx"], start_from=3)
    err = rel_l2(sub3(s), g["sample_from3/sub"])
    print("full-size 3-step chain vs reference", precision, err)
    assert err < (2e-5 if precision == "fp32" else 2e-2)
    np.testing.assert_allclose(s.double().pow(2).sum().item(), g["sample_from3/sum"][1], rtol=1e-4 if precision == "fp32" else 2e-2)


GRAD_TOL = {"fp32": 5e-5, "bf16": 3e-2}  # measured on B200: 1.3e-5 / 1.75e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_training_step_matches_reference_golden(golden, shapes, precision):
    """GaussianDiffusion.forward + loss.backward() (ddpm.py:833-882): loss and every one of the 139 gradients."""
    from oracle.cases import grad_sample
    from util import cpu_seeded_randn

    g = golden["shapes"]
    m, gd = shapes["make"](precision)
    m.train()

    class MD:
        cell_idx = shapes["idx"]

    with cpu_seeded_randn(4321):
        loss, t = gd(shapes["x"], {_key(): shapes["c_local"]}, MD, None)
    np.testing.assert_array_equal(t.cpu().numpy(), g["loss_t"])
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=2e-5 if precision == "fp32" else 2e-2)
    loss.backward()
    assert m.engine().graph_fallbacks == 0
    errs = {}
    for k, p in m.named_parameters():
        want = g[f"grad/{k}/sub"]
        if np.abs(want).max() < 1e-7:  # conv bias in front of GroupNorm: zero up to rounding noise
            assert float(p.grad.abs().max()) < 1e-4, k
            continue
        errs[k] = rel_l2(grad_sample(p.grad), want)
        np.testing.assert_allclose(p.grad.double().pow(2).sum().item(), g[f"grad/{k}/sum"][1], rtol=1e-3 if precision == "fp32" else 8e-2,
                                   err_msg=k)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("full-size gradients vs reference", precision, [(k, f"{v:.2e}") for k, v in worst])
    assert worst[0][1] < GRAD_TOL[precision], worst


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_full_size_upsampling_kernels_agree(precision):
    """The up-sampling into the level-0 concat buffer (97x25x25 -> 194x50x50, 64 channels into a 128-channel pitch) at the exact
    shape the bench runs: the two-stage line kernel (what B >= 3 selects at this size) equals the line walker (B = 2: below the
    1.5 M-row switch) bit for bit, and both equal torch's trilinear interpolation of the same input."""
    import torch.nn.functional as F
    from turbdiff_b200 import _lib

    td, code = (torch.bfloat16, 1) if precision == "bf16" else (torch.float32, 0)  # TDB_BF16 / TDB_F32
    B, C, (Xi, Yi, Zi), (Xo, Yo, Zo) = 2, 64, (97, 25, 25), (194, 50, 50)
    x = torch.randn(B, C, Xi, Yi, Zi, generator=torch.Generator().manual_seed(3)).cuda().to(td)
    xin = torch.zeros((B, Xi + 2, Yi + 2, Zi + 2, C), device="cuda", dtype=td)
    xin[:, 1:-1, 1:-1, 1:-1] = x.permute(0, 2, 3, 4, 1)
    outs = []
    for flag in (0, _lib.TRILINEAR_LINE):
        out = torch.zeros((B, Xo + 2, Yo + 2, Zo + 2, 2 * C), device="cuda", dtype=td)
        _lib.call("tdb_trilinear", xin.data_ptr(), C, Xi, Yi, Zi, out.data_ptr(), 2 * C, Xo, Yo, Zo, B, C, code | flag, _lib.stream_ptr())
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert float(outs[1][..., C:].float().abs().max()) == 0.0  # the skip half of the concat buffer is untouched
    want = F.interpolate(x.float(), size=(Xo, Yo, Zo), mode="trilinear", align_corners=True)
    got = outs[1][:, 1:-1, 1:-1, 1:-1, :C].permute(0, 4, 1, 2, 3).float()
    assert rel_l2(got, want) < (2e-6 if precision == "fp32" else 5e-3)
