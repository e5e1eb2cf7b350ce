"""Oracle (test infrastructure, CPU): the TurbDiff 3-D U-Net denoiser as pure functions
of a flat ``state_dict`` whose keys are the reference's parameter names.

Follows ``turbdiff/models/ddpm.py``:
  NyquistFrequencyEmbedding :103-148, Block :154-177, ResnetBlock :180-197,
  Attention :286-308 (+ attention.py:9-15), UNet.forward :351-372,
  DenoisingModel.__init__/forward :398-505.

Nothing here is used by the product path.  dtype follows the inputs (fp32 or fp64).
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetSpec:
    """Hyper-parameters that fix every tensor shape of the denoiser (ddpm.py:399-412)."""

    in_features: int = 4
    out_features: int = 4
    c_local_features: int = 4
    timesteps: int = 500
    dim: int = 32
    u_net_levels: int = 4
    groups: int | None = 8  # None -> one group per channel ("instance"); 1 -> "layer"
    heads: int = 4
    dim_head: int = 32

    def n_groups(self, channels: int) -> int:
        return channels if self.groups is None else self.groups

    # ---- channel plan (ddpm.py:458-469) -------------------------------------------
    @property
    def c_local_dim(self) -> int:
        return self.dim if self.c_local_features > 0 else 0

    def down_channels(self) -> list[tuple[int, int]]:
        d = self.dim
        plan = [(d + self.c_local_dim, 2 * d)]
        plan += [(d * 2**i, d * 2 ** (i + 1)) for i in range(1, self.u_net_levels)]
        return plan

    def up_channels(self) -> list[tuple[int, int]]:
        d = self.dim
        return [(2 * d * 2 ** (i + 1), d * 2**i) for i in reversed(range(self.u_net_levels))]

    @property
    def center_dim(self) -> int:
        return self.dim * 2**self.u_net_levels


def state_dict_layout(spec: UNetSpec) -> list[tuple[str, tuple[int, ...]]]:
    """Names and shapes of the denoiser's parameters, in the reference's registration
    order (ddpm.py:433-475; checkpoint contract)."""

    out: list[tuple[str, tuple[int, ...]]] = []
    d = spec.dim

    def conv(name, cin, cout, k, bias=True):
        out.append((f"{name}.weight", (cout, cin, k, k, k)))
        if bias:
            out.append((f"{name}.bias", (cout,)))

    def linear(name, cin, cout):
        out.append((f"{name}.weight", (cout, cin)))
        out.append((f"{name}.bias", (cout,)))

    def norm(name, c):
        out.append((f"{name}.weight", (c,)))
        out.append((f"{name}.bias", (c,)))

    def resblock(name, cin, cout):
        linear(f"{name}.project_onto_scale_shift", d, 2 * cout)
        conv(f"{name}.block1.conv", cin, cout, 3)
        norm(f"{name}.block1.norm", cout)
        conv(f"{name}.block2.conv", cout, cout, 3)
        norm(f"{name}.block2.norm", cout)
        if cin != cout:
            conv(f"{name}.conv", cin, cout, 1)

    conv("encode_x", spec.in_features, d, 1)
    if spec.c_local_features > 0:
        conv("encode_c_local", spec.c_local_features, d, 1)
    linear("process_c.0", d, 4 * d)
    linear("process_c.2", 4 * d, d)
    resblock("decode.0", d, d)
    conv("decode.1", d, spec.out_features, 1)
    for i, (cin, cout) in enumerate(spec.down_channels()):
        resblock(f"u_net.downsampling_blocks.{i}", cin, cout)
    for i, (cin, cout) in enumerate(spec.up_channels()):
        resblock(f"u_net.upsampling_blocks.{i}", cin, cout)
    cd = spec.center_dim
    hid = spec.heads * spec.dim_head
    resblock("u_net.center_block.0", cd, cd)
    norm("u_net.center_block.1.fn.norm", cd)
    conv("u_net.center_block.1.fn.fn.to_qkv", cd, 3 * hid, 1, bias=False)
    conv("u_net.center_block.1.fn.fn.to_out", hid, cd, 1)
    resblock("u_net.center_block.2", cd, cd)
    return out


def synth_state_dict(spec: UNetSpec, seed: int, dtype=torch.float32) -> dict[str, torch.Tensor]:
    """Deterministic synthetic weights (numpy PCG64, independent of torch's RNG) so the
    golden generator and the tests can rebuild identical weights without storing them.
    Conv/linear weights ~ U(+-1/sqrt(fan_in)) like torch's default init; norm scales are
    perturbed away from 1 and biases away from 0 so that every term is exercised."""

    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for name, shape in state_dict_layout(spec):
        if ".norm." in name or name.endswith("fn.norm.weight") or name.endswith("fn.norm.bias"):
            if name.endswith("weight"):
                v = 1.0 + 0.2 * rng.standard_normal(shape)
            else:
                v = 0.1 * rng.standard_normal(shape)
        elif name.endswith("weight"):
            fan_in = int(np.prod(shape[1:]))
            v = rng.uniform(-1.0, 1.0, shape) / math.sqrt(fan_in)
        else:
            v = 0.1 * rng.uniform(-1.0, 1.0, shape)
        sd[name] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return sd


# ----------------------------------------------------------------------------- layers


def time_embedding(t: torch.Tensor, dim: int, timesteps: int, dtype=torch.float32) -> torch.Tensor:
    """sin(bias + scale*t), geometric frequencies from 1/8 to Nyquist/(2*phi)
    (ddpm.py:127-148).  scale/bias are rounded to fp32 first, as the reference's
    buffers are."""

    k = dim // 2
    phi = (1 + np.sqrt(5)) / 2
    freqs = np.geomspace(1 / 8, (timesteps / 2) / (2 * phi), num=k)
    scale = torch.tensor(np.repeat(2 * np.pi * freqs / timesteps, 2), dtype=torch.float32)
    bias = torch.tensor(np.tile(np.array([0, np.pi / 2]), k), dtype=torch.float32)
    # addcmul like the reference (a fused multiply-add in fp32: the arguments reach
    # ~480 rad, where one fp32 ulp is 3e-5, so the rounding order is visible)
    return torch.addcmul(bias.to(dtype), scale.to(dtype), t[..., None].to(dtype)).sin()


def silu(x):
    return x * torch.sigmoid(x)


def conv3_replicate(x, w, b):
    """3x3x3 cross-correlation over a replicate-padded input (ddpm.py:164)."""
    return F.conv3d(F.pad(x, (1, 1, 1, 1, 1, 1), mode="replicate"), w, b)


def conv1(x, w, b=None):
    return F.conv3d(x, w, b)


def group_norm(x, groups: int, gamma, beta, eps: float = 1e-5):
    """Biased-variance group normalisation over (C/G, X, Y, Z) (ddpm.py:424-431)."""
    B, C = x.shape[:2]
    xg = x.reshape(B, groups, -1)
    mu = xg.mean(dim=2, keepdim=True)
    var = ((xg - mu) ** 2).mean(dim=2, keepdim=True)
    y = ((xg - mu) * torch.rsqrt(var + eps)).reshape(x.shape)
    view = (1, C) + (1,) * (x.ndim - 2)
    return y * gamma.reshape(view) + beta.reshape(view)


def conv_block(x, sd, name, spec: UNetSpec, film=None):
    """conv -> norm -> [shift + (scale+1)*x] -> SiLU (ddpm.py:168-177)."""
    w = sd[f"{name}.conv.weight"]
    h = conv3_replicate(x, w, sd[f"{name}.conv.bias"])
    h = group_norm(h, spec.n_groups(w.shape[0]), sd[f"{name}.norm.weight"], sd[f"{name}.norm.bias"])
    if film is not None:
        scale, shift = film
        h = shift + (scale + 1) * h
    return silu(h)


def resnet_block(x, c, sd, name, spec: UNetSpec):
    """FiLM projection, two conv blocks, residual with optional 1x1 projection
    (ddpm.py:190-197).  Chunk order: scale first, shift second."""
    ss = F.linear(c, sd[f"{name}.project_onto_scale_shift.weight"], sd[f"{name}.project_onto_scale_shift.bias"])
    cout = ss.shape[-1] // 2
    scale = ss[:, :cout, None, None, None]
    shift = ss[:, cout:, None, None, None]
    h = conv_block(x, sd, f"{name}.block1", spec, film=(scale, shift))
    h = conv_block(h, sd, f"{name}.block2", spec)
    if f"{name}.conv.weight" in sd:
        res = conv1(x, sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"])
    else:
        res = x
    return h + res


def attention_block(x, sd, name, spec: UNetSpec):
    """x + to_out(softmax(q k^T / sqrt(d)) v) with q,k,v = to_qkv(GN(x))
    (ddpm.py:472, 295-308; attention.py:9-15).  Channel index = head*dim_head + d,
    tokens row-major over (X,Y,Z)."""
    B, C = x.shape[:2]
    spatial = x.shape[2:]
    h = group_norm(x, spec.n_groups(C), sd[f"{name}.fn.norm.weight"], sd[f"{name}.fn.norm.bias"])
    qkv = conv1(h, sd[f"{name}.fn.fn.to_qkv.weight"])
    hid = spec.heads * spec.dim_head
    S = int(np.prod(spatial))
    q, k, v = (
        qkv[:, i * hid : (i + 1) * hid].reshape(B, spec.heads, spec.dim_head, S).transpose(2, 3)
        for i in range(3)
    )
    att = torch.softmax((q @ k.transpose(2, 3)) / math.sqrt(spec.dim_head), dim=-1)
    o = (att @ v).transpose(2, 3).reshape(B, hid, *spatial)
    o = conv1(o, sd[f"{name}.fn.fn.to_out.weight"], sd[f"{name}.fn.fn.to_out.bias"])
    return o + x


def _axis_lerp_table(n_in: int, n_out: int, dtype):
    """align_corners=True source positions for one axis: src = i*(n_in-1)/(n_out-1),
    computed in fp32 like ATen's area_pixel_compute_source_index."""
    if n_out > 1:
        ratio = torch.tensor((n_in - 1) / (n_out - 1), dtype=torch.float32)
    else:
        ratio = torch.tensor(0.0, dtype=torch.float32)
    src = torch.arange(n_out, dtype=torch.float32) * ratio
    i0 = src.floor().to(torch.long).clamp_(max=n_in - 1)
    i1 = (i0 + 1).clamp_(max=n_in - 1)
    lam = (src - i0.to(torch.float32)).to(dtype)
    return i0, i1, lam


def trilinear_resample(x, size):
    """F.interpolate(mode='trilinear', align_corners=True) written out as three
    separable gathers + lerps (ddpm.py:358-361, 367-369)."""
    for axis, n_out in zip((2, 3, 4), size):
        n_in = x.shape[axis]
        i0, i1, lam = _axis_lerp_table(n_in, int(n_out), x.dtype)
        shape = [1] * x.ndim
        shape[axis] = -1
        lam = lam.reshape(shape)
        x = (1 - lam) * x.index_select(axis, i0) + lam * x.index_select(axis, i1)
    return x


def downsample_size(spatial) -> list[int]:
    """max(int(s/2), 3) per axis (ddpm.py:358)."""
    return [max(int(s * 0.5), 3) for s in spatial]


def level_sizes(spatial, levels: int) -> list[list[int]]:
    sizes = [list(spatial)]
    for _ in range(levels):
        sizes.append(downsample_size(sizes[-1]))
    return sizes


def process_time(t, sd, spec: UNetSpec, dtype):
    """c = SiLU(W2 SiLU(W1 emb(t) + b1) + b2) (ddpm.py:447-452, 483-493)."""
    emb = time_embedding(t, spec.dim, spec.timesteps, dtype)
    c = silu(F.linear(emb, sd["process_c.0.weight"], sd["process_c.0.bias"]))
    return silu(F.linear(c, sd["process_c.2.weight"], sd["process_c.2.bias"]))


def denoiser_forward(sd, spec: UNetSpec, x, t, c_local, taps: dict | None = None):
    """DenoisingModel.forward (ddpm.py:477-505).  ``c_local`` is the unbatched
    (c, X, Y, Z) concatenation of the local conditioning tensors or None.
    ``taps`` (optional dict) receives named intermediate activations."""

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    dtype = x.dtype
    B = x.shape[0]
    c = tap("c", process_time(t, sd, spec, dtype))
    h = conv1(x, sd["encode_x.weight"], sd["encode_x.bias"])
    if c_local is not None:
        e = conv1(c_local[None], sd["encode_c_local.weight"], sd["encode_c_local.bias"])
        h = torch.cat((h, e.expand(B, -1, -1, -1, -1)), dim=1)
    tap("encoded", h)

    skips = []
    for i in range(spec.u_net_levels):
        h = tap(f"down{i}", resnet_block(h, c, sd, f"u_net.downsampling_blocks.{i}", spec))
        skips.append(h)
        h = tap(f"down{i}.pooled", trilinear_resample(h, downsample_size(h.shape[2:])))

    h = tap("center0", resnet_block(h, c, sd, "u_net.center_block.0", spec))
    h = tap("center1", attention_block(h, sd, "u_net.center_block.1", spec))
    h = tap("center2", resnet_block(h, c, sd, "u_net.center_block.2", spec))

    for i in range(spec.u_net_levels):
        skip = skips.pop()
        h = trilinear_resample(h, skip.shape[2:])
        h = tap(f"up{i}", resnet_block(torch.cat((h, skip), dim=1), c, sd, f"u_net.upsampling_blocks.{i}", spec))

    h = tap("decode0", resnet_block(h, c, sd, "decode.0", spec))
    return conv1(h, sd["decode.1.weight"], sd["decode.1.bias"])
