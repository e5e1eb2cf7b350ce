"""Per-kernel parity on a B200: every C-ABI entry point against the CPU oracle on identical
inputs.  Tolerances: fp32 path rel-L2 <= 1e-5; bf16 path <= 2e-2 (BASELINE.json north_star);
integer / indexing / update-arithmetic kernels bit-exact."""

import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import from_halo, halo_is_replicated, rel_l2, to_halo

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 2e-2}


@pytest.fixture(scope="module")
def lib():
    from turbdiff_b200 import _lib

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.load()
    return _lib


def _dt(prec):
    return (0, torch.float32) if prec == "fp32" else (1, torch.bfloat16)


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


# --------------------------------------------------------------------------- convolution
CONV_CASES = [
    # B, X, Y, Z, Cin, Cout, ntaps
    (2, 9, 7, 6, 16, 16, 27),
    (1, 17, 9, 9, 32, 64, 27),
    (2, 12, 6, 5, 64, 64, 27),
    (1, 8, 6, 6, 128, 32, 27),
    (1, 6, 3, 3, 256, 512, 27),
    (2, 10, 6, 6, 64, 128, 1),
    (1, 12, 3, 3, 512, 384, 1),
    (1, 34, 18, 18, 64, 64, 27),
]


def _conv_ref(x, w, b, ntaps):
    if ntaps == 27:
        return F.conv3d(F.pad(x, (1,) * 6, mode="replicate"), w, b)
    return F.conv3d(x, w, b)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3d_f32(lib, case):
    B, X, Y, Z, Cin, Cout, ntaps = case
    k = 3 if ntaps == 27 else 1
    x = gen(B, Cin, X, Y, Z, seed=1)
    w = gen(Cout, Cin, k, k, k, seed=2, scale=1 / math.sqrt(Cin * ntaps))
    b = gen(Cout, seed=3, scale=0.1)
    xin = to_halo(x, ld=Cin + 8, c0=8)  # exercise a channel-slice view
    wp = w.permute(2, 3, 4, 1, 0).reshape(ntaps, Cin, Cout).contiguous()
    out = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda")
    lib.call("tdb_conv3d_f32", xin.data_ptr() + 8 * 4, Cin + 8, wp.data_ptr(), b.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z,
             Cin, Cout, ntaps, lib.stream_ptr())
    want = _conv_ref(x.double().cpu(), w.double().cpu(), b.double().cpu(), ntaps)
    assert rel_l2(from_halo(out), want) < 1e-5


@pytest.mark.parametrize("case", CONV_CASES + [(2, 12, 3, 3, 512, 512, 27), (1, 6, 6, 6, 1024, 256, 27)])
@pytest.mark.parametrize("fused_stats", [False, True])
@pytest.mark.parametrize("splitk,multicast", [(False, False), (True, False), (False, True)])
def test_conv3d_bf16_tensor_core(lib, case, fused_stats, splitk, multicast):
    B, X, Y, Z, Cin, Cout, ntaps = case
    k = 3 if ntaps == 27 else 1
    x = gen(B, Cin, X, Y, Z, seed=1).bfloat16().float()
    w = gen(Cout, Cin, k, k, k, seed=2, scale=1 / math.sqrt(Cin * ntaps)).bfloat16().float()
    b = gen(Cout, seed=3, scale=0.1)
    xin = to_halo(x, dtype=torch.bfloat16)
    wp = w.permute(0, 2, 3, 4, 1).reshape(Cout, ntaps * Cin).contiguous().bfloat16()
    out = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda", dtype=torch.bfloat16)
    G = 8
    stats = torch.zeros((B, G, 2), dtype=torch.float64, device="cuda")
    scratch = torch.empty(B * (X + 2) * (Y + 2) * (Z + 2) * Cout, dtype=torch.float32, device="cuda") if splitk else None
    lib.call("tdb_conv3d_bf16", xin.data_ptr(), Cin, wp.data_ptr(), b.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
             ntaps, stats.data_ptr() if fused_stats else None, G, lib.CONV_CLUSTER_MC if multicast else 0, lib.ptr(scratch), lib.stream_ptr())
    torch.cuda.synchronize()
    want = _conv_ref(x.double().cpu(), w.double().cpu(), b.double().cpu(), ntaps)
    # inputs are exactly representable in bf16, accumulation is fp32: only the bf16 output rounding remains
    assert rel_l2(from_halo(out), want) < 4e-3
    if fused_stats:
        wg = want.reshape(B, G, -1)
        # with split-K the moments are taken from the bf16-rounded output
        np.testing.assert_allclose(stats[..., 0].cpu().numpy(), wg.sum(-1).numpy(), rtol=1e-4, atol=1e-2 if not splitk else 0.5)
        np.testing.assert_allclose(stats[..., 1].cpu().numpy(), (wg**2).sum(-1).numpy(), rtol=1e-4 if not splitk else 2e-3)


FOLD_CASES = [c for c in CONV_CASES if c[6] == 27 and c[5] in (16, 32, 64)] + [
    (2, 20, 9, 7, 32, 32, 27),     # weights resident in shared memory, several tiles per CTA
    (1, 40, 30, 30, 64, 64, 27),   # 288 tiles over 148 persistent CTAs: both TMEM stages recycle
    (3, 21, 11, 9, 128, 32, 27),
    (1, 13, 6, 5, 48, 16, 27),
]


FOLD2_CASES = [c for c in FOLD_CASES if c[4] % 64 == 0 and c[5] in (32, 64)] + [
    (1, 20, 12, 12, 256, 64, 27), (2, 194, 6, 5, 64, 64, 27),
    # 128-channel N tiles (two N=192 MMAs per K step, single accumulator stage)
    (2, 25, 13, 12, 64, 128, 27), (1, 20, 12, 12, 128, 256, 27), (2, 12, 6, 6, 256, 512, 27), (1, 48, 12, 12, 128, 128, 27),
]


@pytest.mark.parametrize("case,entry", [(c, "tdb_conv3d_bf16_fold") for c in FOLD_CASES] + [(c, "tdb_conv3d_bf16_fold2") for c in FOLD2_CASES])
@pytest.mark.parametrize("fused_stats", [False, True])
def test_conv3d_bf16_kz_folded(lib, case, entry, fused_stats):
    B, X, Y, Z, Cin, Cout, _ = case
    x = gen(B, Cin, X, Y, Z, seed=1).bfloat16().float()
    w = gen(Cout, Cin, 3, 3, 3, seed=2, scale=1 / math.sqrt(Cin * 27)).bfloat16().float()
    b = gen(Cout, seed=3, scale=0.1)
    pad = (Y + 2) * (Z + 2) + 2 * (Z + 2) + 256
    xin = to_halo(x, dtype=torch.bfloat16, pad_rows=pad)
    tile = Cout if Cout < 128 else 128
    wf = w.reshape(Cout // tile, tile, Cin, 3, 3, 3).permute(0, 5, 1, 3, 4, 2).reshape(3 * Cout, 9 * Cin).contiguous().bfloat16()
    out = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda", dtype=torch.bfloat16)
    G = 8
    stats = torch.zeros((B, G, 2), dtype=torch.float64, device="cuda")
    extra = ()
    fuse = entry.endswith("fold2") and Cout <= 64
    if entry.endswith("fold2"):
        # the fused 1x1 residual projection of the same input (only where the engine uses it: Cout <= 64)
        wp = gen(Cout, Cin, 1, 1, 1, seed=21, scale=1 / math.sqrt(Cin)).bfloat16().float()
        bp = gen(Cout, seed=22, scale=0.1)
        outp = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout + 8), device="cuda", dtype=torch.bfloat16)
        wpp = wp.reshape(Cout, Cin).contiguous().bfloat16()
        extra = (wpp.data_ptr(), bp.data_ptr(), outp.data_ptr() + 16, Cout + 8) if fuse else (None, None, None, 0)
    lib.call(entry, xin.data_ptr(), Cin, pad, wf.data_ptr(), b.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
             stats.data_ptr() if fused_stats else None, G, 0, *extra, lib.stream_ptr())
    torch.cuda.synchronize()
    if fuse:
        want_p = _conv_ref(x.double().cpu(), wp.double().cpu(), bp.double().cpu(), 1)
        assert rel_l2(from_halo(outp, Cout, 8), want_p) < 4e-3
    want = _conv_ref(x.double().cpu(), w.double().cpu(), b.double().cpu(), 27)
    assert rel_l2(from_halo(out), want) < 4e-3
    if fused_stats:
        wg = want.reshape(B, G, -1)
        np.testing.assert_allclose(stats[..., 0].cpu().numpy(), wg.sum(-1).numpy(), rtol=1e-4, atol=1e-2)
        np.testing.assert_allclose(stats[..., 1].cpu().numpy(), (wg**2).sum(-1).numpy(), rtol=1e-4)


# --------------------------------------------------------------------------- GroupNorm / pointwise
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("C,G", [(64, 8), (16, 8), (32, 1), (16, 16), (512, 8)])
def test_gn_stats_and_pointwise(lib, prec, C, G):
    code, td = _dt(prec)
    B, X, Y, Z = 2, 7, 5, 6
    x = (gen(B, C, X, Y, Z, seed=4) * 1.5 + 0.3).to(td).float()
    res = gen(B, C, X, Y, Z, seed=5).to(td).float()
    gamma, beta = gen(C, seed=6) * 0.2 + 1, gen(C, seed=7) * 0.1
    film = gen(B, 2 * C + 6, seed=8) * 0.3
    raw = to_halo(x, dtype=td)
    raw[:, 0] = 77.0  # halo rows of a conv output are garbage: must never be read
    resg = to_halo(res, dtype=td)
    stats = torch.zeros((B, G, 2), dtype=torch.float64, device="cuda")
    lib.call("tdb_gn_stats", raw.data_ptr(), C, stats.data_ptr(), B, X, Y, Z, C, G, code, lib.stream_ptr())
    xg = x.double().cpu().reshape(B, G, -1)
    np.testing.assert_allclose(stats[..., 0].cpu().numpy(), xg.sum(-1).numpy(), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(stats[..., 1].cpu().numpy(), (xg**2).sum(-1).numpy(), rtol=1e-6)

    out = torch.zeros((B, X + 2, Y + 2, Z + 2, C), device="cuda", dtype=td)
    lib.call("tdb_pointwise", raw.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), film.data_ptr() + 4 * 3,
             2 * C + 6, resg.data_ptr(), C, out.data_ptr(), C, B, X, Y, Z, C, G, 1e-5, lib.PW_SILU, code, lib.stream_ptr())
    xd = x.double().cpu()
    h = F.group_norm(xd, G, gamma.double().cpu(), beta.double().cpu(), eps=1e-5)
    fl = film.double().cpu()[:, 3 : 3 + 2 * C]
    h = fl[:, C:, None, None, None] + (fl[:, :C, None, None, None] + 1) * h
    want = F.silu(h) + res.double().cpu()
    assert rel_l2(from_halo(out), want) < (2e-6 if prec == "fp32" else 6e-3)
    assert halo_is_replicated(out)

    # plain residual add without norm / activation, interior only
    out2 = torch.full_like(out, 5.0)
    lib.call("tdb_pointwise", raw.data_ptr(), C, None, None, None, None, 0, resg.data_ptr(), C, out2.data_ptr(), C, B, X, Y, Z, C,
             1, 1e-5, lib.PW_NOHALO, code, lib.stream_ptr())
    assert rel_l2(from_halo(out2), xd + res.double().cpu()) < (1e-6 if prec == "fp32" else 6e-3)
    assert float(out2[:, 0].float().min()) == 5.0


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("C", [16, 24, 64, 256])  # 16-byte channel vectors per voxel: 2/3/8/32 (bf16), 4/6/16/64 (fp32)
@pytest.mark.parametrize("sizes", [((11, 7, 5), (5, 3, 3)), ((5, 3, 3), (11, 7, 5)), ((12, 3, 3), (24, 6, 6)), ((6, 6, 6), (6, 6, 6)),
                                   ((12, 6, 6), (25, 12, 12)), ((4, 3, 3), (14, 9, 3)), ((3, 3, 7), (5, 4, 30))])
def test_trilinear(lib, prec, sizes, C):
    """Down-sampling runs on the gather kernel; up-sampling on the line walker and, when forced (or for large outputs) and the
    channel vectors of a voxel divide a warp (C = 16, 64, 256 in bf16; 16, 64 in fp32), on the two-stage line kernel."""
    code, td = _dt(prec)
    (Xi, Yi, Zi), (Xo, Yo, Zo) = sizes
    B = 2
    x = gen(B, C, Xi, Yi, Zi, seed=9).to(td).float()
    xin = to_halo(x, dtype=td)
    out = torch.zeros((B, Xo + 2, Yo + 2, Zo + 2, 2 * C), device="cuda", dtype=td)
    lib.call("tdb_trilinear", xin.data_ptr(), C, Xi, Yi, Zi, out.data_ptr() + C * out.element_size(), 2 * C, Xo, Yo, Zo, B, C, code,
             lib.stream_ptr())
    want = F.interpolate(x.double().cpu(), size=(Xo, Yo, Zo), mode="trilinear", align_corners=True)
    assert rel_l2(from_halo(out, C, C), want) < (2e-6 if prec == "fp32" else 5e-3)
    assert float(out[..., :C].float().abs().max()) == 0.0  # the other half of the concat buffer is untouched
    assert halo_is_replicated(out[..., C:].contiguous())
    if Xo * Yo * Zo > Xi * Yi * Zi:
        # the two-stage line kernel (production: outputs of >= 1.5 M rows) must equal the walker bit for bit
        out2 = torch.zeros_like(out)
        lib.call("tdb_trilinear", xin.data_ptr(), C, Xi, Yi, Zi, out2.data_ptr() + C * out.element_size(), 2 * C, Xo, Yo, Zo, B, C,
                 code | lib.TRILINEAR_LINE, lib.stream_ptr())
        assert torch.equal(out2, out)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("C", [16, 64, 512])
@pytest.mark.parametrize("sizes", [((11, 7, 5), (5, 3, 3)), ((5, 3, 3), (11, 7, 5)), ((12, 3, 3), (24, 6, 6)), ((6, 6, 6), (6, 6, 6)),
                                   ((3, 3, 3), (6, 6, 6)), ((24, 6, 6), (12, 3, 3)), ((25, 12, 12), (50, 25, 25)), ((4, 3, 3), (14, 9, 3))])
def test_trilinear_bwd_is_adjoint(lib, prec, C, sizes):
    """tdb_trilinear_bwd = autograd of F.interpolate(trilinear, align_corners=True) (reference ddpm.py:358-369), halo rows of
    the result zero; with TDB_TRIBWD_ACCUMULATE the transposed gradient is added onto d_in's interior and its halo is kept."""
    code, td = _dt(prec)
    (Xi, Yi, Zi), (Xo, Yo, Zo) = sizes
    if C == 512 and Xi * Yi * Zi > 2000:
        pytest.skip("large grid x wide channels adds nothing")
    B = 2
    g = gen(B, C, Xo, Yo, Zo, seed=19).to(td).float()
    x = torch.zeros(B, C, Xi, Yi, Zi, dtype=torch.float64, requires_grad=True)
    (want,) = torch.autograd.grad(F.interpolate(x, size=(Xo, Yo, Zo), mode="trilinear", align_corners=True), x, g.double().cpu())
    gg = to_halo(g, dtype=td, ld=2 * C, c0=C)
    gg[:, 0] = 55.0  # halo rows of the incoming gradient are never read
    d_in = torch.full((B, Xi + 2, Yi + 2, Zi + 2, C), 3.0, device="cuda", dtype=td)
    off = C * gg.element_size()
    lib.call("tdb_trilinear_bwd", gg.data_ptr() + off, 2 * C, Xo, Yo, Zo, d_in.data_ptr(), C, Xi, Yi, Zi, B, C, code, 0, lib.stream_ptr())
    tol = 2e-6 if prec == "fp32" else 6e-3
    assert rel_l2(from_halo(d_in), want) < tol
    halo = d_in.clone()
    halo[:, 1:-1, 1:-1, 1:-1] = 0
    assert float(halo.float().abs().max()) == 0.0
    # accumulate form
    base = gen(B, C, Xi, Yi, Zi, seed=20).to(td).float()
    acc = to_halo(base, dtype=td)
    acc[:, 0] = 7.0
    lib.call("tdb_trilinear_bwd", gg.data_ptr() + off, 2 * C, Xo, Yo, Zo, acc.data_ptr(), C, Xi, Yi, Zi, B, C, code,
             lib.TRIBWD_ACCUMULATE, lib.stream_ptr())
    assert rel_l2(from_halo(acc), want + base.double().cpu()) < tol
    assert float(acc[:, 0].float().min()) == 7.0 and float(acc[:, 0].float().max()) == 7.0


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("C,G,with_film,size", [(64, 8, True, (9, 7, 6)), (16, 8, False, (7, 5, 6)), (32, 1, True, (20, 9, 11)),
                                                (512, 8, True, (6, 3, 3)), (32, 32, False, (40, 12, 10))])
def test_pointwise_bwd_matches_autograd(lib, prec, C, G, with_film, size):
    """tdb_pointwise_bwd_reduce / _finalize / _apply against torch.autograd of GroupNorm -> FiLM -> SiLU
    (reference Block.forward, ddpm.py:168-177): input gradient, norm / FiLM / bias-sum gradients."""
    code, td = _dt(prec)
    X, Y, Z = size
    B = 2
    x = (gen(B, C, X, Y, Z, seed=4) * 1.5 + 0.3).to(td).float()
    g = gen(B, C, X, Y, Z, seed=5).to(td).float()
    gamma, beta = gen(C, seed=6) * 0.2 + 1, gen(C, seed=7) * 0.1
    film = gen(B, 2 * C, seed=8) * 0.3
    xd = x.double().cpu().requires_grad_(True)
    gm, bt, fl = gamma.double().cpu().requires_grad_(True), beta.double().cpu().requires_grad_(True), film.double().cpu().requires_grad_(True)
    h = F.group_norm(xd, G, gm, bt, eps=1e-5)
    if with_film:
        h = fl[:, C:, None, None, None] + (fl[:, :C, None, None, None] + 1) * h
    y = F.silu(h)
    want = torch.autograd.grad(y, [xd, gm, bt] + ([fl] if with_film else []), g.double().cpu())

    raw = to_halo(x, dtype=td)
    raw[:, 0] = 77.0
    gg = to_halo(g, dtype=td)
    gg[:, 0] = -31.0
    stats = torch.zeros((B, G, 2), dtype=torch.float64, device="cuda")
    lib.call("tdb_gn_stats", raw.data_ptr(), C, stats.data_ptr(), B, X, Y, Z, C, G, code, lib.stream_ptr())
    red = torch.zeros((B, C, 4), dtype=torch.float64, device="cuda")
    fptr = film.data_ptr() if with_film else None
    lib.call("tdb_pointwise_bwd_reduce", gg.data_ptr(), C, raw.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), fptr,
             2 * C, red.data_ptr(), B, X, Y, Z, C, G, 1e-5, lib.PW_SILU, code, lib.stream_ptr())
    out = torch.empty((4, C), dtype=torch.float32, device="cuda")
    grp = torch.empty((B, G, 2), dtype=torch.float32, device="cuda")
    dfilm = torch.zeros((B, 2 * C), dtype=torch.float32, device="cuda")
    lib.call("tdb_pointwise_bwd_finalize", red.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), fptr, 2 * C,
             grp.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
             dfilm.data_ptr() if with_film else None, 2 * C, B, X, Y, Z, C, G, 1e-5, lib.stream_ptr())
    d_raw = torch.full((B, X + 2, Y + 2, Z + 2, C), 9.0, device="cuda", dtype=td)
    lib.call("tdb_pointwise_bwd_apply", gg.data_ptr(), C, raw.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), fptr,
             2 * C, grp.data_ptr(), d_raw.data_ptr(), C, B, X, Y, Z, C, G, 1e-5, lib.PW_SILU, code, lib.stream_ptr())
    tol = 2e-5 if prec == "fp32" else 1.2e-2
    assert rel_l2(from_halo(d_raw), want[0]) < tol
    halo = d_raw.clone()
    halo[:, 1:-1, 1:-1, 1:-1] = 0
    assert float(halo.float().abs().max()) == 0.0
    ptol = 2e-5 if prec == "fp32" else 4e-3
    assert rel_l2(out[1], want[1]) < ptol and rel_l2(out[2], want[2]) < ptol
    assert rel_l2(out[0], want[0].sum(dim=(0, 2, 3, 4))) < (1e-3 if prec == "fp32" else 3e-2) or float(want[0].sum(dim=(0, 2, 3, 4)).abs().max()) < 1e-6
    assert rel_l2(out[3], g.double().cpu().sum(dim=(0, 2, 3, 4))) < 1e-5
    if with_film:
        assert rel_l2(dfilm, want[3]) < ptol
    # the fused form (finalize folded into apply) gives the same input gradient and parameter gradients
    d_raw2 = torch.full_like(d_raw, 9.0)
    out2 = torch.empty_like(out)
    dfilm2 = torch.zeros_like(dfilm)
    lib.call("tdb_pointwise_bwd_apply_fused", gg.data_ptr(), C, raw.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), fptr,
             2 * C, red.data_ptr(), d_raw2.data_ptr(), C, out2[0].data_ptr(), out2[1].data_ptr(), out2[2].data_ptr(), out2[3].data_ptr(),
             dfilm2.data_ptr() if with_film else None, 2 * C, B, X, Y, Z, C, G, 1e-5, lib.PW_SILU, code, lib.stream_ptr())
    assert torch.equal(d_raw2, d_raw)
    torch.testing.assert_close(out2, out, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(dfilm2, dfilm, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("spatial", [(12, 3, 3), (8, 4, 4), (4, 2, 2), (3, 3, 16), (10, 7, 7)])  # S = 108, 128, 16, 144, 490
def test_attention(lib, prec, spatial):
    code, td = _dt(prec)
    X, Y, Z = spatial
    B, heads, dh = 2, 4, 32
    hid = heads * dh
    qkv = gen(B, 3 * hid, X, Y, Z, seed=10).to(td).float()
    qg = to_halo(qkv, dtype=td)
    out = torch.zeros((B, X + 2, Y + 2, Z + 2, hid), device="cuda", dtype=td)
    lib.call("tdb_attention", qg.data_ptr(), 3 * hid, out.data_ptr(), hid, B, X, Y, Z, heads, dh, code, lib.stream_ptr())
    S = X * Y * Z
    q, k, v = (qkv.double().cpu()[:, i * hid : (i + 1) * hid].reshape(B, heads, dh, S).transpose(2, 3) for i in range(3))
    want = F.scaled_dot_product_attention(q, k, v).transpose(2, 3).reshape(B, hid, X, Y, Z)
    assert rel_l2(from_halo(out), want) < (2e-6 if prec == "fp32" else 5e-3)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("spatial", [(12, 3, 3), (8, 4, 4), (5, 3, 2), (3, 3, 16), (9, 5, 5), (10, 7, 7)])
def test_attention_backward_matches_autograd(lib, prec, spatial):
    """tdb_attention_bwd against torch.autograd of scaled_dot_product_attention (reference attention.py:9-15).  Sequences of
    up to 135 voxels run on the kernel that keeps the S x S matrices in shared memory, longer ones (144, 225, 490 here) on
    the streaming form."""
    code, td = _dt(prec)
    X, Y, Z = spatial
    B, heads, dh = 3, 4, 32
    hid, S = heads * dh, X * Y * Z
    qkv = gen(B, 3 * hid, X, Y, Z, seed=10).to(td).float()
    go = gen(B, hid, X, Y, Z, seed=11).to(td).float()
    qd = qkv.double().cpu().requires_grad_(True)
    q, k, v = (qd[:, i * hid : (i + 1) * hid].reshape(B, heads, dh, S).transpose(2, 3) for i in range(3))
    o = F.scaled_dot_product_attention(q, k, v).transpose(2, 3).reshape(B, hid, X, Y, Z)
    (want,) = torch.autograd.grad(o, qd, go.double().cpu())
    qg, gg = to_halo(qkv, dtype=td), to_halo(go, dtype=td)
    d_qkv = torch.zeros((B, X + 2, Y + 2, Z + 2, 3 * hid), device="cuda", dtype=td)
    lib.call("tdb_attention_bwd", qg.data_ptr(), 3 * hid, gg.data_ptr(), hid, d_qkv.data_ptr(), 3 * hid, B, X, Y, Z, heads, dh, code,
             lib.stream_ptr())
    assert rel_l2(from_halo(d_qkv), want) < (5e-6 if prec == "fp32" else 8e-3)


def test_time_film(lib):
    from oracle.unet_ref import UNetSpec, process_time, synth_state_dict

    spec = UNetSpec(dim=32, u_net_levels=1, timesteps=500)
    sd = synth_state_dict(spec, 5)
    t = torch.tensor([0, 1, 250, 499], dtype=torch.int64)
    c_ref = process_time(t, sd, spec, torch.float32)
    fw = torch.randn(96, 32, generator=torch.Generator().manual_seed(1)) * 0.2
    fb = torch.randn(96, generator=torch.Generator().manual_seed(2)) * 0.1
    film_ref = F.linear(c_ref, fw, fb)
    from turbdiff_b200.models.ddpm import NyquistFrequencyEmbedding

    emb = NyquistFrequencyEmbedding(32, 500).cuda()
    d = {k: v.cuda() for k, v in sd.items()}
    c = torch.zeros(4, 32, device="cuda")
    film = torch.zeros(4, 96, device="cuda")
    fwc, fbc, tc = fw.t().contiguous().cuda(), fb.cuda(), t.cuda()
    lib.call("tdb_time_film", tc.data_ptr(), emb.scale.data_ptr(), emb.bias.data_ptr(), d["process_c.0.weight"].data_ptr(),
             d["process_c.0.bias"].data_ptr(), d["process_c.2.weight"].data_ptr(), d["process_c.2.bias"].data_ptr(), fwc.data_ptr(),
             fbc.data_ptr(), c.data_ptr(), film.data_ptr(), 4, 32, 96, lib.stream_ptr())
    assert rel_l2(c, c_ref) < 1e-5
    assert rel_l2(film, film_ref) < 1e-5


# --------------------------------------------------------------------------- encode / decode
@pytest.mark.parametrize("B", [1, 4, 9])
def test_time_film_backward_matches_autograd(lib, B):
    """tdb_time_film_bwd against torch.autograd of the Nyquist embedding -> process_c MLP -> FiLM projections
    (reference ddpm.py:147-148, 447-452, 184/191)."""
    dim, rows, T = 32, 2 * (32 + 64 + 64 + 128) + 14, 500
    g = torch.Generator().manual_seed(17)
    scale, bias = torch.rand(dim, generator=g) * 0.9 + 0.01, torch.rand(dim, generator=g) * 1.5
    w1, b1 = torch.randn(4 * dim, dim, generator=g) * 0.2, torch.randn(4 * dim, generator=g) * 0.1
    w2, b2 = torch.randn(dim, 4 * dim, generator=g) * 0.1, torch.randn(dim, generator=g) * 0.1
    fw, fb = torch.randn(rows, dim, generator=g) * 0.2, torch.randn(rows, generator=g) * 0.1
    t = torch.randint(0, T, (B,), generator=g)
    d_film = torch.randn(B, rows, generator=g)
    P = [v.double().requires_grad_(True) for v in (w1, b1, w2, b2, fw, fb)]
    emb = torch.sin(bias.double() + scale.double() * t.double()[:, None])
    c = F.silu(F.linear(F.silu(F.linear(emb, P[0], P[1])), P[2], P[3]))
    film = F.linear(c, P[4], P[5])
    want = torch.autograd.grad(film, P, d_film.double())
    dev = [v.cuda().contiguous() for v in (t, scale, bias, w1, b1, w2, b2, fw.t().contiguous(), d_film)]
    c_dev = torch.empty((B, dim), device="cuda")
    film_dev = torch.empty((B, rows), device="cuda")
    lib.call("tdb_time_film", dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr(), dev[5].data_ptr(),
             dev[6].data_ptr(), dev[7].data_ptr(), fb.cuda().data_ptr(), c_dev.data_ptr(), film_dev.data_ptr(), B, dim, rows, lib.stream_ptr())
    assert rel_l2(film_dev, film) < 1e-5
    out = [torch.full(s, 7.0, device="cuda") for s in ((rows, dim), (rows,), (4 * dim, dim), (4 * dim,), (dim, 4 * dim), (dim,), (B, dim))]
    lib.call("tdb_time_film_bwd", dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr(),
             dev[5].data_ptr(), dev[6].data_ptr(), dev[7].data_ptr(), c_dev.data_ptr(), dev[8].data_ptr(), *(o.data_ptr() for o in out), B, dim,
             rows, lib.stream_ptr())
    got = {"w1": out[2], "b1": out[3], "w2": out[4], "b2": out[5], "fw": out[0], "fb": out[1]}
    for name, w in zip(("w1", "b1", "w2", "b2", "fw", "fb"), want):
        assert rel_l2(got[name], w) < 2e-5, name


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_encode_decode(lib, prec):
    code, td = _dt(prec)
    B, Fx, Fc, dim, X, Y, Z = 2, 4, 4, 16, 9, 6, 5
    x, cl = gen(B, Fx, X, Y, Z, seed=11), gen(Fc, X, Y, Z, seed=12)
    wx, bx, wc, bc = gen(dim, Fx, seed=13), gen(dim, seed=14), gen(dim, Fc, seed=15), gen(dim, seed=16)
    out = torch.zeros((B, X + 2, Y + 2, Z + 2, 2 * dim), device="cuda", dtype=td)
    lib.call("tdb_encode_input", x.data_ptr(), cl.data_ptr(), wx.data_ptr(), bx.data_ptr(), wc.data_ptr(), bc.data_ptr(),
             out.data_ptr(), 2 * dim, B, Fx, Fc, dim, X, Y, Z, 3, code, lib.stream_ptr())
    ex = F.conv3d(x.double().cpu(), wx.double().cpu()[:, :, None, None, None], bx.double().cpu())
    ec = F.conv3d(cl.double().cpu()[None], wc.double().cpu()[:, :, None, None, None], bc.double().cpu()).expand(B, -1, -1, -1, -1)
    tol = 2e-6 if prec == "fp32" else 5e-3
    assert rel_l2(from_halo(out), torch.cat((ex, ec), 1)) < tol
    assert halo_is_replicated(out)

    wd, bd = gen(Fx, 2 * dim, seed=17), gen(Fx, seed=18)
    y = torch.zeros((B, Fx, X, Y, Z), device="cuda")
    lib.call("tdb_decode_output", out.data_ptr(), 2 * dim, wd.data_ptr(), bd.data_ptr(), y.data_ptr(), B, X, Y, Z, 2 * dim, Fx, code,
             lib.stream_ptr())
    want = F.conv3d(from_halo(out).double().cpu(), wd.double().cpu()[:, :, None, None, None], bd.double().cpu())
    assert rel_l2(y, want) < 2e-6


# --------------------------------------------------------------------------- diffusion kernels (bit-exact)
def _coef(T=10, name="log-snr-linear"):
    from oracle.diffusion_ref import diffusion_buffers

    b = diffusion_buffers(name, T)
    tab = torch.stack((b["sqrt_recip_alphas_cumprod"], b["sqrt_recipm1_alphas_cumprod"], b["posterior_mean_coef1"],
                       b["posterior_mean_coef2"], (b["log_betas"] / 2).exp(), b["sqrt_alphas_cumprod"],
                       b["sqrt_one_minus_alphas_cumprod"], torch.zeros(T)), dim=1).contiguous()
    return b, tab


@pytest.mark.parametrize("noise_bcs", [True, False])
@pytest.mark.parametrize("clip", [False, True])
@pytest.mark.parametrize("nvox", [7 * 5 * 4, 6 * 5 * 3 + 1])
def test_ddpm_step_bit_exact(lib, noise_bcs, clip, nvox):
    from oracle.diffusion_ref import DiffusionRef

    B, Fx, T = 2, 4, 10
    shape = (B, Fx, nvox, 1, 1)
    g = torch.Generator().manual_seed(3)
    x_t, eps, z, zbc, xb = (torch.randn(shape, generator=g) for _ in range(5))
    idx = torch.randperm(nvox, generator=g)[: nvox * 2 // 3]
    mask = torch.zeros(nvox, dtype=torch.uint8)
    mask[idx] = 1
    b, tab = _coef(T)
    d = DiffusionRef(None, timesteps=T, beta_schedule="log-snr-linear", noise_bcs=noise_bcs, clip_denoised=clip)
    from oracle.diffusion_ref import where_cells

    for t in (7, 1, 0):
        tt = torch.full((B,), t, dtype=torch.long)
        x0 = d.predict_start(x_t, tt, eps)
        if not noise_bcs:
            x0 = where_cells(idx, x0, x_t)
        if clip:
            x0 = x0.clamp(-1, 1)
        mean = d.posterior_mean(x0, x_t, tt)
        if t == 0:
            want = mean
        else:
            zz = z if noise_bcs else where_cells(idx, z)
            want = mean + (b["log_betas"][t] / 2).exp() * zz
            if noise_bcs:
                want = where_cells(idx, want, d.q_sample(xb, tt, zbc))
        for final in (False, True):
            w2 = where_cells(idx, want, xb) if final else want
            flags = (lib.STEP_NOISE_BCS if noise_bcs else 0) | (lib.STEP_CLIP if clip else 0) | (lib.STEP_FINAL if final else 0)
            dx, de, dz, dzb, dxb, dm, dtab = (v.cuda().contiguous() for v in (x_t, eps, z, zbc, xb, mask, tab))
            t_dev = torch.tensor([t], dtype=torch.int32, device="cuda")
            out = torch.empty_like(dx)
            lib.call("tdb_ddpm_step", dx.data_ptr(), de.data_ptr(), dz.data_ptr(), dzb.data_ptr(), dxb.data_ptr(), dm.data_ptr(),
                     dtab.data_ptr(), t_dev.data_ptr(), out.data_ptr(), B, Fx, nvox, flags, lib.stream_ptr())
            assert torch.equal(out.cpu(), w2), (t, final)


def test_q_sample_loss_and_cells(lib):
    from oracle.diffusion_ref import DiffusionRef, select_cells, where_cells

    B, Fx, T, nvox = 3, 4, 10, 6 * 5 * 4
    shape = (B, Fx, 6, 5, 4)
    g = torch.Generator().manual_seed(5)
    x0, noise, eps = (torch.randn(shape, generator=g) for _ in range(3))
    idx = torch.randperm(nvox, generator=g)[:77]
    t = torch.tensor([0, 4, 9])
    _, tab = _coef(T)
    d = DiffusionRef(None, timesteps=T, beta_schedule="log-snr-linear")
    dm = torch.zeros(nvox, dtype=torch.uint8, device="cuda")
    didx = idx.cuda()
    lib.call("tdb_build_mask", didx.data_ptr(), dm.data_ptr(), idx.numel(), nvox, lib.stream_ptr())
    want_mask = torch.zeros(nvox, dtype=torch.uint8)
    want_mask[idx] = 1
    assert torch.equal(dm.cpu(), want_mask)
    dx0, dn, de, dtab, dt = x0.cuda(), noise.cuda(), eps.cuda(), tab.cuda(), t.cuda()
    for noise_bcs in (1, 0):
        out = torch.empty_like(dx0)
        lib.call("tdb_q_sample", dx0.data_ptr(), dn.data_ptr(), dt.data_ptr(), dtab.data_ptr(), dm.data_ptr(), out.data_ptr(), B, Fx,
                 nvox, noise_bcs, lib.stream_ptr())
        want = d.q_sample(x0, t, noise)
        if not noise_bcs:
            want = where_cells(idx, want, x0)
        assert torch.equal(out.cpu(), want)
    # masked loss + grad
    for l1 in (0, 1):
        acc = torch.zeros(1, dtype=torch.float64, device="cuda")
        grad = torch.empty_like(de)
        lib.call("tdb_masked_loss", de.data_ptr(), dn.data_ptr(), dm.data_ptr(), acc.data_ptr(), grad.data_ptr(), B, Fx, nvox,
                 idx.numel(), l1, lib.stream_ptr())
        e = eps.double().requires_grad_()
        per = (e - noise.double()).abs() if l1 else (e - noise.double()) ** 2
        loss = per.flatten(-3)[..., idx].reshape(B, -1).mean(1).mean()
        loss.backward()
        np.testing.assert_allclose(acc.item(), loss.item(), rtol=1e-6)
        assert rel_l2(grad, e.grad) < 1e-6
    # where / select / scatter
    other = torch.randn(shape, generator=g)
    out = torch.empty_like(dx0)
    lib.call("tdb_where_cells", dx0.data_ptr(), other.cuda().data_ptr(), dm.data_ptr(), out.data_ptr(), B * Fx, nvox, lib.stream_ptr())
    assert torch.equal(out.cpu(), where_cells(idx, x0, other))
    lib.call("tdb_where_cells", dx0.data_ptr(), None, dm.data_ptr(), out.data_ptr(), B * Fx, nvox, lib.stream_ptr())
    assert torch.equal(out.cpu(), where_cells(idx, x0))
    sel = torch.empty((B, Fx, idx.numel()), device="cuda")
    lib.call("tdb_select_cells", dx0.data_ptr(), didx.data_ptr(), sel.data_ptr(), B * Fx, nvox, idx.numel(), lib.stream_ptr())
    assert torch.equal(sel.cpu(), select_cells(x0, idx))
    samples = torch.randn(B, idx.numel(), Fx, generator=g)
    grid = torch.zeros((B, Fx, nvox), device="cuda")
    lib.call("tdb_scatter_cells", samples.cuda().data_ptr(), didx.data_ptr(), grid.data_ptr(), B, Fx, nvox, idx.numel(), lib.stream_ptr())
    want = torch.zeros(B, Fx, nvox)
    want[:, :, idx] = samples.transpose(1, 2)
    assert torch.equal(grid.cpu(), want)
    # empty index set
    lib.call("tdb_select_cells", dx0.data_ptr(), didx.data_ptr(), sel.data_ptr(), B * Fx, nvox, 0, lib.stream_ptr())


# --------------------------------------------------------------------------- convolution weight gradient
WGRAD_CASES = [
    # B, X, Y, Z, Cin, Cout, ntaps
    (2, 12, 6, 5, 64, 64, 27),
    (1, 17, 9, 9, 128, 32, 27),
    (2, 9, 7, 6, 32, 32, 27),
    (1, 10, 6, 6, 64, 128, 27),
    (1, 8, 5, 5, 256, 64, 27),
    (1, 6, 4, 4, 128, 256, 27),
    (1, 6, 3, 3, 512, 512, 27),
    (2, 10, 6, 6, 64, 128, 1),
    (1, 34, 18, 18, 64, 64, 27),
    (1, 40, 20, 50, 32, 32, 27),
    (2, 20, 12, 12, 256, 64, 27),
    (1, 30, 12, 30, 128, 32, 27),
]


def _wgrad_ref(x, dy, ntaps):
    """d/dw of sum(conv(x, w) * dy) in float64 on the CPU (torch.autograd, like the reference's loss.backward())."""
    x, dy = x.double().cpu(), dy.double().cpu()
    k = 3 if ntaps == 27 else 1
    w = torch.zeros(dy.shape[1], x.shape[1], k, k, k, dtype=torch.float64, requires_grad=True)
    (g,) = torch.autograd.grad(_conv_ref(x, w, None, ntaps), w, dy)
    return g


def _wgrad_inputs(case, ld_extra=0):
    B, X, Y, Z, Cin, Cout, ntaps = case
    x = gen(B, Cin, X, Y, Z, seed=31).bfloat16().float()
    dy = gen(B, Cout, X, Y, Z, seed=32).bfloat16().float()
    xin = to_halo(x, dtype=torch.bfloat16, ld=Cin + ld_extra, c0=ld_extra)
    dyh = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda", dtype=torch.bfloat16)  # zero halo
    dyh[:, 1:-1, 1:-1, 1:-1, :] = dy.permute(0, 2, 3, 4, 1).bfloat16()
    return x, dy, xin, dyh


def _dw_to_torch(dw, ntaps, Cin, Cout):
    k = 3 if ntaps == 27 else 1
    return dw.view(k, k, k, Cin, Cout).permute(4, 3, 0, 1, 2)


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_conv3d_wgrad_tensor_core(lib, case, mode):
    """tcgen05 weight gradient (MN-major operands from the halo grids); mode 1 shares one row window per kz triple,
    mode 3 additionally stacks the kz taps on the N side for Cout in {32, 64}."""
    B, X, Y, Z, Cin, Cout, ntaps = case
    ld_extra = 8 if Cin == 64 else 0  # channel-pitched input view (concat slices)
    x, dy, xin, dyh = _wgrad_inputs(case, ld_extra)
    dw = torch.zeros((ntaps, Cin, Cout), dtype=torch.float32, device="cuda")
    lib.call("tdb_conv3d_wgrad_tc", xin.data_ptr() + 2 * ld_extra, Cin + ld_extra, dyh.data_ptr(), Cout, dw.data_ptr(), B, X, Y, Z,
             Cin, Cout, ntaps, mode, lib.stream_ptr())
    torch.cuda.synchronize()
    assert rel_l2(_dw_to_torch(dw, ntaps, Cin, Cout), _wgrad_ref(x, dy, ntaps)) < 1e-4


@pytest.mark.parametrize("case", WGRAD_CASES[:4] + WGRAD_CASES[7:8])
@pytest.mark.parametrize("prec,zero_halo", [("fp32", False), ("bf16", False), ("bf16", True)])
def test_conv3d_wgrad_dispatch(lib, case, prec, zero_halo):
    """tdb_conv3d_wgrad: masked SIMT path (halo rows of d_out hold garbage) and the zero-halo tensor-core path."""
    B, X, Y, Z, Cin, Cout, ntaps = case
    code, dt = _dt(prec)
    x, dy, xin, dyh = _wgrad_inputs(case)
    xin, dyh = xin.to(dt), dyh.to(dt)
    if not zero_halo:
        dyh = to_halo(dy, dtype=dt)  # non-zero halo rows must be ignored
    dw = torch.zeros((ntaps, Cin, Cout), dtype=torch.float32, device="cuda")
    lib.call("tdb_conv3d_wgrad", xin.data_ptr(), Cin, dyh.data_ptr(), Cout, dw.data_ptr(), B, X, Y, Z, Cin, Cout, ntaps, code,
             lib.WGRAD_ZERO_HALO if zero_halo else 0, lib.stream_ptr())
    torch.cuda.synchronize()
    assert rel_l2(_dw_to_torch(dw, ntaps, Cin, Cout), _wgrad_ref(x, dy, ntaps)) < 1e-4


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("C,F_,size", [(32, 4, (21, 9, 7)), (16, 8, (6, 5, 4)), (32, 3, (40, 12, 11)), (24, 4, (5, 4, 3))])
def test_cl_nc_outer_and_colsum(lib, prec, C, F_, size):
    """tdb_cl_nc_outer: weight (and bias) gradient of a 1x1x1 convolution between a halo grid and NCDHW planes
    (reference encode_x / encode_c_local / decode[1], ddpm.py:433,436,459, through autograd)."""
    code, td = _dt(prec)
    X, Y, Z = size
    B = 3
    g = gen(B, C, X, Y, Z, seed=51).to(td).float()
    q = gen(B, F_, X, Y, Z, seed=52)
    gh = to_halo(g, dtype=td, ld=C + 8, c0=8)
    gh[:, 0] = 9.0
    out = torch.zeros((C, F_), dtype=torch.float32, device="cuda")
    cs = torch.zeros(C, dtype=torch.float32, device="cuda")
    lib.call("tdb_cl_nc_outer", gh.data_ptr() + 8 * gh.element_size(), C + 8, q.data_ptr(), F_ * X * Y * Z, out.data_ptr(), cs.data_ptr(),
             B, X, Y, Z, C, F_, code, lib.stream_ptr())
    want = torch.einsum("bcxyz,bfxyz->cf", g.double().cpu(), q.double().cpu())
    assert rel_l2(out, want) < 1e-5
    assert rel_l2(cs, g.double().cpu().sum(dim=(0, 2, 3, 4))) < 1e-5
    # unbatched Q (q_bstride = 0), no colsum
    out2 = torch.zeros((C, F_), dtype=torch.float32, device="cuda")
    lib.call("tdb_cl_nc_outer", gh.data_ptr() + 8 * gh.element_size(), C + 8, q.data_ptr(), 0, out2.data_ptr(), None, B, X, Y, Z, C, F_, code,
             lib.stream_ptr())
    want2 = torch.einsum("bcxyz,fxyz->cf", g.double().cpu(), q[0].double().cpu())
    assert rel_l2(out2, want2) < 1e-5


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("size", [(6, 4, 3), (2, 5, 2), (4, 1, 3), (13, 4, 4)])
def test_halo_fold_is_adjoint_of_replicate_pad(lib, prec, size):
    """tdb_halo_fold: border voxels collect the gradient of their halo images (autograd of F.pad(mode="replicate")),
    and the halo rows are left zero."""
    code, dt = _dt(prec)
    B, C = 2, 16
    X, Y, Z = size
    g = gen(B, X + 2, Y + 2, Z + 2, C, seed=41).to(dt)
    x = torch.zeros(B, C, X, Y, Z, dtype=torch.float64, requires_grad=True)
    (want,) = torch.autograd.grad(F.pad(x, (1,) * 6, mode="replicate"), x, g.double().cpu().permute(0, 4, 1, 2, 3))
    buf = g.clone()
    lib.call("tdb_halo_fold", buf.data_ptr(), C, B, X, Y, Z, C, code, lib.stream_ptr())
    torch.cuda.synchronize()
    assert rel_l2(from_halo(buf), want) < (1e-6 if prec == "fp32" else 8e-3)
    halo = buf.clone()
    halo[:, 1:-1, 1:-1, 1:-1, :] = 0
    assert float(halo.abs().max()) == 0.0


# --------------------------------------------------------------------------- row-window CTA-pair convolution
WIN_CASES = [
    # B, X, Y, Z, Cin, Cout, fused projection
    (2, 12, 6, 5, 64, 64, False),
    (1, 17, 9, 9, 128, 32, True),
    (2, 9, 7, 6, 32, 32, False),
    (1, 8, 6, 6, 32, 128, False),
    (1, 10, 6, 6, 32, 64, True),
    (1, 34, 18, 18, 64, 64, False),
    (1, 40, 30, 30, 64, 64, False),    # several tiles per CTA pair: both TMEM stages and the window ring recycle
    (2, 40, 20, 50, 32, 32, False),    # widest supported z-line class (Z + 2 = 52): 240-row windows
    (3, 21, 11, 9, 128, 32, False),
    # streamed weights, N tiles of 128 channels (weights too large to stay resident)
    (2, 25, 13, 12, 64, 128, True), (1, 20, 12, 12, 128, 256, True), (2, 12, 6, 6, 256, 512, False), (1, 48, 12, 12, 128, 128, False),
    (1, 12, 6, 6, 512, 128, True),
    # tiny grids: 64-channel N tiles of the streamed-weight variant (the bottleneck level)
    (4, 12, 3, 3, 512, 512, False), (1, 12, 3, 3, 256, 256, True),
]


WINZ_CASES = [c for c in WIN_CASES if c[5] in (32, 64) and 27 * c[4] * c[5] <= 116 * 1024] + [
    (1, 30, 12, 62, 32, 32, False),    # Z + 2 = 64: the widest window (256 rows)
    (2, 11, 5, 3, 64, 32, True), (1, 194, 6, 5, 128, 32, True),
]


@pytest.mark.parametrize("case,entry", [(c, "tdb_conv3d_bf16_win") for c in WIN_CASES] + [(c, "tdb_conv3d_bf16_winz") for c in WINZ_CASES])
@pytest.mark.parametrize("variant", ["plain", "stats", "all_rows"])
def test_conv3d_bf16_row_window(lib, case, entry, variant):
    B, X, Y, Z, Cin, Cout, with_proj = case
    x = gen(B, Cin, X, Y, Z, seed=1).bfloat16().float()
    w = gen(Cout, Cin, 3, 3, 3, seed=2, scale=1 / math.sqrt(Cin * 27)).bfloat16().float()
    b = gen(Cout, seed=3, scale=0.1)
    xin = to_halo(x, dtype=torch.bfloat16, ld=Cin + 8, c0=8)  # channel-pitched input view, no padding rows
    if entry.endswith("winz"):  # kz folded into N: row = kz*Cout + co, column = (kx*3 + ky)*Cin + ci
        wk = w.permute(4, 0, 2, 3, 1).reshape(3 * Cout, 9 * Cin).contiguous().bfloat16()
    else:
        wk = w.permute(0, 2, 3, 4, 1).reshape(Cout, 27 * Cin).contiguous().bfloat16()
    out = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda", dtype=torch.bfloat16)
    G = 8
    stats = torch.zeros((B, G, 2), dtype=torch.float64, device="cuda")
    fuse = with_proj and variant != "all_rows"
    extra = (None, None, None, 0)
    if fuse:
        wp = gen(Cout, Cin, 1, 1, 1, seed=21, scale=1 / math.sqrt(Cin)).bfloat16().float()
        bp = gen(Cout, seed=22, scale=0.1)
        outp = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout + 8), device="cuda", dtype=torch.bfloat16)
        wpp = wp.reshape(Cout, Cin).contiguous().bfloat16()
        extra = (wpp.data_ptr(), bp.data_ptr(), outp.data_ptr() + 16, Cout + 8)
    lib.call(entry, xin.data_ptr() + 16, Cin + 8, wk.data_ptr(), b.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin,
             Cout, stats.data_ptr() if variant == "stats" else None, G, lib.CONV_ALL_ROWS if variant == "all_rows" else 0, *extra,
             lib.stream_ptr())
    torch.cuda.synchronize()
    want = _conv_ref(x.double().cpu(), w.double().cpu(), b.double().cpu(), 27)
    assert rel_l2(from_halo(out), want) < 4e-3
    if fuse:
        want_p = _conv_ref(x.double().cpu(), wp.double().cpu(), bp.double().cpu(), 1)
        assert rel_l2(from_halo(outp, Cout, 8), want_p) < 4e-3
    if variant == "stats":
        wg = want.reshape(B, G, -1)
        np.testing.assert_allclose(stats[..., 0].cpu().numpy(), wg.sum(-1).numpy(), rtol=1e-4, atol=1e-2)
        np.testing.assert_allclose(stats[..., 1].cpu().numpy(), (wg**2).sum(-1).numpy(), rtol=1e-4)
    if variant == "all_rows":
        # input gradients: the input has a ZERO halo, and every row (halo rows too) is the zero-padded convolution
        xz = torch.zeros((B, X + 2, Y + 2, Z + 2, Cin + 8), device="cuda", dtype=torch.bfloat16)
        xz[:, 1:-1, 1:-1, 1:-1, 8:] = x.permute(0, 2, 3, 4, 1).bfloat16()
        lib.call(entry, xz.data_ptr() + 16, Cin + 8, wk.data_ptr(), b.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z,
                 Cin, Cout, None, G, lib.CONV_ALL_ROWS, None, None, None, 0, lib.stream_ptr())
        torch.cuda.synchronize()
        full = F.conv3d(F.pad(F.pad(x.double().cpu(), (1,) * 6), (1,) * 6), w.double().cpu(), b.double().cpu())
        assert rel_l2(out.permute(0, 4, 1, 2, 3).float(), full) < 4e-3


ADD1X1_CASES = [c for c in WIN_CASES if c[5] == 128 or c[5] % 128 == 0] + [(2, 20, 10, 10, 32, 128, False), (1, 25, 13, 13, 64, 256, False)]


@pytest.mark.parametrize("case", ADD1X1_CASES)
def test_conv3d_bf16_row_window_add1x1(lib, case):
    """tdb_conv3d_bf16_win_add1x1: out = conv3x3x3(in) + conv1x1(in2) in one launch, every row stored - the input gradient
    of a ResnetBlock's first convolution plus that of its residual projection (autograd of reference ddpm.py:190-197)."""
    B, X, Y, Z, Cin, Cout, _ = case
    x = gen(B, Cin, X, Y, Z, seed=1).bfloat16().float()
    x2 = gen(B, Cin, X, Y, Z, seed=11).bfloat16().float()
    w = gen(Cout, Cin, 3, 3, 3, seed=2, scale=1 / math.sqrt(Cin * 27)).bfloat16().float()
    w2 = gen(Cout, Cin, 1, 1, 1, seed=12, scale=1 / math.sqrt(Cin)).bfloat16().float()
    wk = w.permute(0, 2, 3, 4, 1).reshape(Cout, 27 * Cin).contiguous().bfloat16()
    w2k = w2.reshape(Cout, Cin).contiguous().bfloat16()

    def zero_halo(v, ld, c0):
        g = torch.zeros((B, X + 2, Y + 2, Z + 2, ld), device="cuda", dtype=torch.bfloat16)
        g[:, 1:-1, 1:-1, 1:-1, c0 : c0 + Cin] = v.permute(0, 2, 3, 4, 1).bfloat16()
        return g

    xz, x2z = zero_halo(x, Cin + 8, 8), zero_halo(x2, 2 * Cin, Cin)
    out = torch.full((B, X + 2, Y + 2, Z + 2, Cout), 3.0, device="cuda", dtype=torch.bfloat16)
    lib.call("tdb_conv3d_bf16_win_add1x1", xz.data_ptr() + 16, Cin + 8, wk.data_ptr(), None, out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
             lib.CONV_ALL_ROWS, x2z.data_ptr() + 2 * Cin, 2 * Cin, w2k.data_ptr(), lib.stream_ptr())
    torch.cuda.synchronize()
    full = F.conv3d(F.pad(F.pad(x.double().cpu(), (1,) * 6), (1,) * 6), w.double().cpu())
    full = full + F.pad(F.conv3d(x2.double().cpu(), w2.double().cpu()), (1,) * 6)
    assert rel_l2(out.permute(0, 4, 1, 2, 3).float(), full) < 4e-3


@pytest.mark.parametrize("cout,cin,taps", [(64, 64, 27), (32, 128, 27), (512, 256, 27), (16, 16, 27), (48, 80, 27), (384, 512, 1), (32, 128, 1)])
@pytest.mark.parametrize("transpose", [0, 1])
@pytest.mark.parametrize("folded", [0, 1])
def test_pack_conv_weights_matches_the_torch_layouts(lib, cout, cin, taps, transpose, folded):
    """tdb_pack_conv_weights: bf16 kernel layouts (per-tap / kz-folded, forward / tap-reversed transpose for input gradients)
    bit-identical to the torch permute + flip + cast sequence they replace."""
    if folded and taps != 27:
        pytest.skip("the kz-folded layout exists for 3x3x3 weights only")
    k = 3 if taps == 27 else 1
    w = gen(cout, cin, k, k, k, seed=77)
    wl = w.flip(2, 3, 4).transpose(0, 1) if transpose else w  # logical (O, I, k, k, k)
    O, I = wl.shape[:2]
    tile = O if O < 128 else 128
    if folded and O % tile:
        pytest.skip("N tile does not divide the channel count")
    if folded:
        want = wl.reshape(O // tile, tile, I, 3, 3, 3).permute(0, 5, 1, 3, 4, 2).reshape(3 * O, 9 * I).contiguous().to(torch.bfloat16)
    else:
        want = wl.permute(0, 2, 3, 4, 1).reshape(O, taps * I).contiguous().to(torch.bfloat16)
    got = torch.full_like(want, 7.0)
    lib.call("tdb_pack_conv_weights", w.data_ptr(), got.data_ptr(), cout, cin, taps, folded, tile, transpose, lib.stream_ptr())
    assert torch.equal(got, want)


@pytest.mark.parametrize("cout,cin,taps", [(64, 64, 27), (32, 128, 27), (512, 256, 27), (48, 80, 27), (20, 12, 27), (384, 512, 1), (128, 32, 1), (4, 32, 1)])
def test_unpack_wgrad_equals_the_torch_permute(lib, cout, cin, taps):
    """tdb_unpack_wgrad: [taps][Cin][Cout] -> (Cout, Cin, k, k, k), bit-identical to the permute + contiguous it replaces."""
    k = 3 if taps == 27 else 1
    dw = gen(taps, cin, cout, seed=88)
    want = dw.view(k, k, k, cin, cout).permute(4, 3, 0, 1, 2).contiguous()
    got = torch.full_like(want, 7.0)
    lib.call("tdb_unpack_wgrad", dw.data_ptr(), got.data_ptr(), cout, cin, taps, lib.stream_ptr())
    assert torch.equal(got, want)


def test_pack_conv_weights_batch_equals_the_single_launches(lib):
    """tdb_pack_conv_weights_batch (the job table as a kernel parameter, 64 weights per launch): 70 weights of mixed shapes,
    tap counts and layouts in two launches, each bit-identical to its own tdb_pack_conv_weights launch."""
    import ctypes

    shapes = [(64, 64, 27), (32, 128, 27), (512, 256, 27), (16, 16, 27), (48, 80, 27), (384, 512, 1), (32, 128, 1), (128, 64, 27), (256, 128, 1),
              (64, 32, 27)]
    jobs, want, got, keep = [], [], [], []
    for n in range(70):
        cout, cin, taps = shapes[n % len(shapes)]
        transpose = (n // 2) % 2
        O = cin if transpose else cout
        tile = O if O < 128 else 128
        folded = 1 if (taps == 27 and O % tile == 0 and n % 3 == 0) else 0
        k = 3 if taps == 27 else 1
        w = gen(cout, cin, k, k, k, seed=300 + n)
        a = torch.full((cout * cin * taps,), 7.0, device="cuda", dtype=torch.bfloat16)
        b = torch.full_like(a, 5.0)
        lib.call("tdb_pack_conv_weights", w.data_ptr(), a.data_ptr(), cout, cin, taps, folded, tile, transpose, lib.stream_ptr())
        jobs.append(lib.PackJob(w.data_ptr(), b.data_ptr(), cout, cin, taps, folded, tile, transpose))
        want.append(a)
        got.append(b)
        keep.append(w)
    table = (lib.PackJob * len(jobs))(*jobs)
    n0 = lib.launch_count()
    lib.call("tdb_pack_conv_weights_batch", ctypes.cast(table, ctypes.c_void_p), len(jobs), lib.stream_ptr())
    assert lib.launch_count() - n0 == 2
    for n, (a, b) in enumerate(zip(want, got)):
        assert torch.equal(a, b), n


# --------------------------------------------------------------------------- fused scatter/normalise, gather/de-normalise
@pytest.mark.parametrize("B", [1, 3])
def test_scatter_normalize_and_gather_denormalize(lib, B):
    """grid_embedding + normalize_grid, and denormalize_grid + select_cells + 'b f c -> b c f', each in
    one launch - bit-exact against the reference's torch op sequence on the same device, and equal to the CPU oracle."""
    from oracle import grid_ref
    from turbdiff_b200.models import utils as U

    geo = grid_ref.channel_geometry(cells=(12, 7, 5), hole=((3, 6), (2, 5), (0, 3)), seed=3)
    n_cells, F = len(geo.cell_idx), 4
    g = torch.Generator().manual_seed(5)
    samples = torch.randn(B, n_cells, F, generator=g)
    mean = torch.tensor([0.3, -1.2, 0.05, 101.5])
    std = torch.tensor([1.7, 0.4, 2.5, 13.0])
    idx = torch.from_numpy(geo.cell_idx).cuda()

    # reference op sequence (ofles.py:220-232, normalization.py:20-24) with torch on the GPU
    x = torch.zeros((B, F, geo.n_vox), device="cuda")
    x.transpose(-1, -2)[..., idx, :] = samples.cuda()
    x = x.view(B, F, *geo.padded)
    m3, s3 = mean.cuda().view(F, 1, 1, 1), std.cuda().view(F, 1, 1, 1)
    want = torch.addcmul(-m3 / s3, torch.reciprocal(s3), x)
    got = U.scatter_normalize(samples.cuda(), idx, geo.padded, mean, std)
    assert torch.equal(got, want)
    oracle = grid_ref.normalize_grid(grid_ref.grid_embedding(geo, [samples.numpy()], [{}]), mean.numpy(), std.numpy())
    np.testing.assert_allclose(got.cpu().numpy(), oracle, rtol=1e-6, atol=1e-6)

    # denormalize_grid + select_cells + rearrange (normalization.py:26-30, utils.py:14-15, metrics.py:50-57)
    xs = torch.randn(B, F, *geo.padded, generator=g).cuda()
    want2 = torch.addcmul(m3, s3, xs).flatten(-3)[..., idx].permute(0, 2, 1).contiguous()
    got2 = U.gather_denormalize(xs, idx, mean, std)
    assert torch.equal(got2, want2)
    oracle2 = np.swapaxes(grid_ref.select_cells(grid_ref.denormalize_grid(xs.cpu().numpy(), mean.numpy(), std.numpy()), geo.cell_idx), 1, 2)
    np.testing.assert_allclose(got2.cpu().numpy(), oracle2, rtol=1e-6, atol=1e-5)
    # round trip: scatter/normalise then de-normalise/gather returns the samples (to fp32 rounding)
    back = U.gather_denormalize(got, idx, mean, std)
    np.testing.assert_allclose(back.cpu().numpy(), samples.numpy(), rtol=1e-5, atol=2e-5)


def test_scatter_normalize_writes_fixed_value_boundaries(lib, golden):
    """data/ofles.py:233-238: FIXED_VALUE boundary vectors (inlet U = (20,0,0), wall U = 0, outlet p = 0) written into the
    padding voxels by the same launch: identity normalisation against the reference's grid_embedding golden (bit-exact),
    a real normalisation against the reference's torch op sequence on the device (bit-exact)."""
    from test_host_cpu import _golden_grid_case
    from turbdiff_b200.models import utils as U

    geo, samples, fixed = _golden_grid_case()
    idx = torch.from_numpy(geo.cell_idx).cuda()
    F_ = 4
    got = U.scatter_normalize(samples.cuda(), idx, geo.padded, torch.zeros(F_), torch.ones(F_), fixed_values=fixed)
    np.testing.assert_array_equal(got.cpu().numpy(), golden["grid"]["grid_embedding"])

    mean, std = torch.tensor([0.3, -1.2, 0.05, 101.5]), torch.tensor([1.7, 0.4, 2.5, 13.0])
    x = torch.zeros((2, F_, geo.n_vox), device="cuda")
    xt = x.transpose(-1, -2)
    xt[..., idx, :] = samples.cuda()
    for bidx, f0, v in fixed:
        xt[..., bidx.cuda(), f0 : f0 + v.numel()] = v.cuda()
    m3, s3 = mean.cuda().view(F_, 1, 1, 1), std.cuda().view(F_, 1, 1, 1)
    want = torch.addcmul(-m3 / s3, torch.reciprocal(s3), x.view(2, F_, *geo.padded))
    tables = U.boundary_code(idx, geo.n_vox, F_, fixed)
    got = U.scatter_normalize(samples.cuda(), idx, geo.padded, mean, std, tables=tables)
    assert torch.equal(got, want)
    # a boundary that overlaps cells and another boundary: the later write wins, per channel
    over = fixed + [(torch.from_numpy(geo.cell_idx[:5].copy()), 1, torch.tensor([4.5])), (fixed[0][0][:3], 0, torch.tensor([1.0, 2.0, 3.0]))]
    for bidx, f0, v in over[-2:]:
        xt[..., bidx.cuda(), f0 : f0 + v.numel()] = v.cuda()
    want = torch.addcmul(-m3 / s3, torch.reciprocal(s3), x.view(2, F_, *geo.padded))
    assert torch.equal(U.scatter_normalize(samples.cuda(), idx, geo.padded, mean, std, fixed_values=over), want)


@pytest.mark.parametrize("case", [(2, 9, 7, 6), (2, 40, 20, 50), (1, 30, 12, 62), (3, 21, 11, 8), (1, 194, 6, 4)])
@pytest.mark.parametrize("variant", ["plain", "stats", "all_rows"])
def test_conv3d_bf16_paired_rows_32_to_32(lib, case, variant):
    """tdb_conv3d_bf16_winp: 32 -> 32 channels with two grid rows per 128-byte TMA row (even Z + 2, input pitch 32)."""
    B, X, Y, Z = case
    Cin = Cout = 32
    x = gen(B, Cin, X, Y, Z, seed=1).bfloat16().float()
    w = gen(Cout, Cin, 3, 3, 3, seed=2, scale=1 / math.sqrt(Cin * 27)).bfloat16().float()
    b = gen(Cout, seed=3, scale=0.1)
    wk = w.permute(4, 0, 2, 3, 1).reshape(3 * Cout, 9 * Cin).contiguous().bfloat16()
    out = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda", dtype=torch.bfloat16)
    G = 8
    stats = torch.zeros((B, G, 2), dtype=torch.float64, device="cuda")
    if variant == "all_rows":  # input gradients: zero halo in, every row out
        xin = torch.zeros((B, X + 2, Y + 2, Z + 2, Cin), device="cuda", dtype=torch.bfloat16)
        xin[:, 1:-1, 1:-1, 1:-1] = x.permute(0, 2, 3, 4, 1).bfloat16()
    else:
        xin = to_halo(x, dtype=torch.bfloat16)
    lib.call("tdb_conv3d_bf16_winp", xin.data_ptr(), Cin, wk.data_ptr(), b.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
             stats.data_ptr() if variant == "stats" else None, G, lib.CONV_ALL_ROWS if variant == "all_rows" else 0, lib.stream_ptr())
    torch.cuda.synchronize()
    if variant == "all_rows":
        full = F.conv3d(F.pad(F.pad(x.double().cpu(), (1,) * 6), (1,) * 6), w.double().cpu(), b.double().cpu())
        assert rel_l2(out.permute(0, 4, 1, 2, 3).float(), full) < 4e-3
        return
    want = _conv_ref(x.double().cpu(), w.double().cpu(), b.double().cpu(), 27)
    assert rel_l2(from_halo(out), want) < 4e-3
    # the staged TMA-store epilogue writes whole tiles: rows it does not own a value for (the halo) must come out as zeros
    halo = out.clone()
    halo[:, 1:-1, 1:-1, 1:-1] = 0
    assert not halo.any()
    if variant == "stats":
        wg = want.reshape(B, G, -1)
        np.testing.assert_allclose(stats[..., 0].cpu().numpy(), wg.sum(-1).numpy(), rtol=1e-4, atol=1e-2)
        np.testing.assert_allclose(stats[..., 1].cpu().numpy(), (wg**2).sum(-1).numpy(), rtol=1e-4)
