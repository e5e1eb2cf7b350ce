"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(`/root/reference`, importable only in the build container) on deterministic inputs.

    python tests/golden/make_golden.py

The reference imports a few packages that are absent here (pytorch_lightning, h5py,
omegaconf, more_itertools, lightning_utilities); they are stubbed in ``sys.modules``
before ``import turbdiff``.  None of the stubbed code is on the
denoising path.  Inputs and weights come from numpy PCG64 seeds
(``oracle.unet_ref.synth_state_dict``), so the fixtures only hold the reference's OUTPUTS;
``tests/test_oracle_golden.py`` rebuilds the inputs and checks the oracle against them.
"""

import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))


import warnings  # noqa: E402

warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402  (stubs the absent third-party packages, then imports the unmodified reference)

_ns = ref_shim.load()
ref, ofles, ref_utils = _ns.ddpm, _ns.ofles, _ns.utils
CellTypeLearnedEmbedding, Conditioning = _ns.CellTypeLearnedEmbedding, _ns.Conditioning

from oracle import grid_ref  # noqa: E402
from oracle.cases import (CASES, LV_ELBO_WEIGHT, LV_SEEDS, LV_VARIANTS, SHAPES_FWD_T, SHAPES_INPUT_SEED, SHAPES_SEED, SHAPES_T, case_inputs, grad_sample,  # noqa: E402
                          lv_case, shapes_spec, sub3, tap_sample)
from oracle.unet_ref import UNetSpec, state_dict_layout, synth_state_dict  # noqa: E402

NORM_NAME = {8: "group", 1: "layer", None: "instance"}


def build_ref_model(spec: UNetSpec, sd):
    m = ref.DenoisingModel(
        in_features=spec.in_features, out_features=spec.out_features, c_local_features=spec.c_local_features,
        c_global_features=0, timesteps=spec.timesteps, dim=spec.dim, u_net_levels=spec.u_net_levels,
        norm_type=NORM_NAME[spec.groups],
    )
    m.load_state_dict(sd, strict=True)
    return m.eval()


def gen_schedules():
    out = {}
    for name in ("linear", "log-linear", "log-snr-linear", "cosine", "sigmoid"):
        for T in (10, 500, 1000):
            gd = ref.GaussianDiffusion(torch.nn.Identity(), timesteps=T, beta_schedule=name)
            for b, v in gd.named_buffers():
                out[f"{name}/{T}/{b}"] = v.numpy()
    np.savez_compressed(HERE / "schedules.npz", **out)


def gen_time_embedding():
    out = {}
    for dim, T in ((32, 500), (32, 1000), (16, 10), (8, 10)):
        emb = ref.NyquistFrequencyEmbedding(dim, T)
        out[f"{dim}/{T}"] = emb(torch.arange(T)).numpy()
        out[f"{dim}/{T}/scale"] = emb.scale.numpy()
    np.savez_compressed(HERE / "time_embedding.npz", **out)


def gen_layout():
    spec = UNetSpec()  # shapes config
    with torch.device("meta"):
        m = ref.DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0,
                               timesteps=500, dim=32, u_net_levels=4, norm_type="group")
    ref_layout = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    mine = [(k, list(s)) for k, s in state_dict_layout(spec)]
    assert ref_layout == mine, "state-dict layout of the oracle differs from the reference"
    (HERE / "state_dict_layout_shapes.json").write_text(json.dumps(ref_layout))
    print("layout entries", len(ref_layout), "params", sum(int(np.prod(s)) for _, s in ref_layout))


def gen_unet():
    out = {}
    for cname, case in CASES.items():
        spec = case["spec"]
        sd = synth_state_dict(spec, case["seed"])
        m = build_ref_model(spec, sd)
        x, t, c_local, _ = case_inputs(case)
        taps = {}
        hooks = []

        def grab(name):
            def hook(_mod, _inp, outp):
                taps[name] = outp.detach()
            return hook

        un = m.u_net
        for i, blk in enumerate(un.downsampling_blocks):
            hooks.append(blk.register_forward_hook(grab(f"down{i}")))
        for i, blk in enumerate(un.upsampling_blocks):
            hooks.append(blk.register_forward_hook(grab(f"up{i}")))
        for i in range(3):
            hooks.append(un.center_block[i].register_forward_hook(grab(f"center{i}")))
        hooks.append(m.decode[0].register_forward_hook(grab("decode0")))
        hooks.append(m.process_c.register_forward_hook(grab("c")))
        with torch.no_grad():
            y = m(x, t, {Conditioning.Type.CELL_TYPE: c_local})
        for h in hooks:
            h.remove()
        out[f"{cname}/out"] = y.numpy()
        if case.get("save_taps"):
            for k, v in taps.items():
                out[f"{cname}/tap/{k}"] = v.numpy()
        else:  # only cheap checksums of the intermediates
            for k, v in taps.items():
                out[f"{cname}/tapsum/{k}"] = np.array([v.double().sum().item(), v.double().pow(2).sum().item()])
        print(cname, "out", tuple(y.shape), float(y.abs().mean()))
    np.savez_compressed(HERE / "unet.npz", **out)


class _MD:
    def __init__(self, idx):
        self.cell_idx = idx


def gen_diffusion():
    out = {}
    for cname in ("micro", "tiny"):
        case = CASES[cname]
        spec = case["spec"]
        sd = synth_state_dict(spec, case["seed"])
        x, t, c_local, geo = case_inputs(case)
        cell_idx = torch.from_numpy(geo.cell_idx)
        C = {Conditioning.Type.CELL_TYPE: c_local}
        for noise_bcs in (True, False):
            tag = f"{cname}/noise_bcs={int(noise_bcs)}"
            m = build_ref_model(spec, sd)
            gd = ref.GaussianDiffusion(m, timesteps=spec.timesteps, beta_schedule="log-snr-linear",
                                       loss_type="l2", noise_bcs=noise_bcs)
            # sampling, full chain and start_from
            torch.manual_seed(1234)
            out[f"{tag}/sample"] = gd.p_sample_loop(x, C, cell_idx).numpy()
            torch.manual_seed(1234)
            out[f"{tag}/sample_from4"] = gd.p_sample_loop(x, C, cell_idx, start_from=4).numpy()
            # one p_sample at t=3 and t=0
            for tt in (3, 0):
                mean, lv = gd.p_sample(x, tt, C, cell_idx)
                out[f"{tag}/p_sample_mean/{tt}"] = mean.numpy()
                out[f"{tag}/p_sample_logvar/{tt}"] = lv.numpy()
            # training loss + grads
            m.train()
            torch.manual_seed(4321)
            loss, tdraw = gd(x, C, _MD(cell_idx), None)
            loss.backward()
            out[f"{tag}/loss"] = np.array(loss.item())
            out[f"{tag}/t"] = tdraw.numpy()
            for k, p in m.named_parameters():
                g = p.grad
                if g.numel() <= 4096 and cname == "micro":
                    out[f"{tag}/grad/{k}"] = g.numpy()
                out[f"{tag}/gradsum/{k}"] = np.array([g.double().sum().item(), g.double().pow(2).sum().item()])
            if cname == "micro":
                # l1 loss value too
                gd1 = ref.GaussianDiffusion(m, timesteps=spec.timesteps, beta_schedule="log-snr-linear",
                                            loss_type="l1", noise_bcs=noise_bcs)
                torch.manual_seed(4321)
                out[f"{tag}/loss_l1"] = np.array(gd1(x, C, _MD(cell_idx), None)[0].item())
            print(tag, "loss", loss.item())
    np.savez_compressed(HERE / "diffusion.npz", **out)


def gen_diffusion_lv():
    """learned_variances=True (ddpm.py:732-741) with the ELBO term (ddpm.py:853-870): sampling chains, p_sample,
    training loss and every gradient of the unmodified reference."""
    out = {}
    case = lv_case()
    spec = case["spec"]
    sd = synth_state_dict(spec, case["seed"])
    x, t, c_local, geo = case_inputs(case)
    cell_idx = torch.from_numpy(geo.cell_idx)
    C = {Conditioning.Type.CELL_TYPE: c_local}
    for noise_bcs, detach in LV_VARIANTS:
        tag = f"noise_bcs={int(noise_bcs)}/detach={int(detach)}"
        m = build_ref_model(spec, sd)
        gd = ref.GaussianDiffusion(m, timesteps=spec.timesteps, beta_schedule="log-snr-linear", loss_type="l2", noise_bcs=noise_bcs,
                                   learned_variances=True, elbo_weight=LV_ELBO_WEIGHT, detach_elbo_mean=detach)
        if detach:
            # (p_sample_loop itself cannot be pinned: with learned variances the reference raises "only one dimension
            # can be inferred" at ddpm.py:805 - broadcast_right (models/utils.py:11) reshapes the 5-D std with five -1s)
            try:
                gd.p_sample_loop(x, C, cell_idx, start_from=2)
                raise AssertionError("the reference's learned-variance sampling loop unexpectedly works: pin it")
            except RuntimeError as e:
                out[f"{tag}/sample_loop_error"] = np.array(str(e))
            for tt in (3, 0):
                mean, lv = gd.p_sample(x, tt, C, cell_idx)
                out[f"{tag}/p_sample_mean/{tt}"] = mean.numpy()
                out[f"{tag}/p_sample_logvar/{tt}"] = lv.numpy()
        m.train()
        for seed in LV_SEEDS:
            torch.manual_seed(seed)
            loss, tdraw = gd(x, C, _MD(cell_idx), None)
            m.zero_grad()
            loss.backward()
            out[f"{tag}/loss/{seed}"] = np.array(loss.item())
            out[f"{tag}/t/{seed}"] = tdraw.numpy()
            for k, p in m.named_parameters():
                g = p.grad
                if g.numel() <= 4096:
                    out[f"{tag}/grad/{seed}/{k}"] = g.numpy().copy()
                out[f"{tag}/gradsum/{seed}/{k}"] = np.array([g.double().sum().item(), g.double().pow(2).sum().item()])
            print(tag, "seed", seed, "loss", loss.item(), "t", tdraw.tolist())
    np.savez_compressed(HERE / "diffusion_lv.npz", **out)


TKE_STAT_SEEDS = tuple(range(900, 908))  # 8 chains of the tiny configuration x batch 2 = 16 samples


def tke_cubes(samples):
    """The two 16^3 cubes of the tiny configuration's 32x16x16 interior, velocity channels only: (N, 2, 3, 16, 16, 16)
    (the reference cuts cube regions of edge min(Y, Z) out of the channel the same way, metrics.py:424-449)."""
    u = samples[:, :3, 1:-1, 1:-1, 1:-1]
    return torch.stack((u[..., :16, :, :], u[..., 16:, :, :]), dim=1)


def gen_tke():
    """TKE-spectrum statistic through the UNMODIFIED reference classes (models/metrics.py:270-378): known-answer spectra
    and distance matrices on synthetic fields, and the spectra of 16 reference sampling chains of the tiny configuration
    (north_star: "agreement of sample TKE/energy-spectrum statistics")."""
    from oracle import tke_ref

    M = ref_shim.load(with_task=True).metrics
    out = {}
    spec = M.TurbulentKineticEnergySpectrum(n=110)
    out["p110"], out["w110"] = spec.p.numpy(), spec.w.numpy()  # the Lebedev quadrature the fixtures were made with
    for n, seed in ((16, 1), (24, 2)):
        u = torch.from_numpy(tke_ref.synthetic_velocity(3, n, seed))
        um = u.mean(0)
        dist = M.LogTKESpectrumL2Distance(spec, n=16)
        D, la, lb, k = dist(u[:2], u[1:], um)
        out[f"synthetic/{n}/E"] = spec(u - um, k).numpy()
        out[f"synthetic/{n}/D"], out[f"synthetic/{n}/log_a"], out[f"synthetic/{n}/log_b"], out[f"synthetic/{n}/k"] = (
            D.numpy(), la.numpy(), lb.numpy(), k.numpy())
    # the production size: one 48^3 cube, 64 radii, 5810 Lebedev points (only the outputs are stored)
    big = M.LogTKESpectrumL2Distance(M.TurbulentKineticEnergySpectrum(), n=64)
    u = torch.from_numpy(tke_ref.synthetic_velocity(2, 48, 3))
    D, la, lb, k = big(u[:1], u[1:], u.mean(0))
    out["synthetic/48/D"], out["synthetic/48/log_a"], out["synthetic/48/log_b"], out["synthetic/48/k"] = D.numpy(), la.numpy(), lb.numpy(), k.numpy()

    # sample statistics of the tiny configuration: 16 chains of the reference
    case = CASES["tiny"]
    spec_t = case["spec"]
    m = build_ref_model(spec_t, synth_state_dict(spec_t, case["seed"]))
    gd = ref.GaussianDiffusion(m, timesteps=spec_t.timesteps, beta_schedule="log-snr-linear", loss_type="l2", noise_bcs=True)
    x, _, c_local, geo = case_inputs(case)
    C = {Conditioning.Type.CELL_TYPE: c_local}
    idx = torch.from_numpy(geo.cell_idx)
    chains = []
    for seed in TKE_STAT_SEEDS:
        torch.manual_seed(seed)
        chains.append(gd.p_sample_loop(x, C, idx))
    cubes = tke_cubes(torch.cat(chains))            # (16, 2, 3, 16, 16, 16)
    u_mean = cubes.mean(0)
    dist = M.LogTKESpectrumL2Distance(spec, n=16)
    logs = []
    for c in range(2):
        D, la, _, k = dist(cubes[:, c], cubes[:, c], u_mean[c])
        logs.append(la)
        out[f"stat/D/{c}"] = D.numpy()
    out["stat/log_tke"] = torch.stack(logs, dim=1).numpy()   # (16, 2, 16)
    out["stat/u_mean"] = u_mean.numpy()
    out["stat/k"] = k.numpy()
    out["stat/seeds"] = np.array(TKE_STAT_SEEDS)
    np.savez_compressed(HERE / "tke.npz", **out)
    print("tke: D(16^3)", out["synthetic/16/D"].round(3).tolist(), "sample spectra", out["stat/log_tke"].shape,
          "median off-diagonal D", float(np.median(out["stat/D/0"][~np.eye(16, dtype=bool)])))


def gen_grid():
    """grid_embedding / cell types / cell helpers through the reference's own dataclasses
    (SURVEY.md appendix C)."""
    V = ofles.Variable
    BC = ofles.BoundaryCondition
    out = {}
    geo = grid_ref.channel_geometry(cells=(12, 6, 5), hole=((3, 6), (1, 4), (0, 3)), seed=3)
    rng = np.random.Generator(np.random.PCG64(11))
    B, n = 2, len(geo.cell_idx)
    u = rng.standard_normal((B, n, 3)).astype(np.float32)
    p = rng.standard_normal((B, n, 1)).astype(np.float32)
    bnd = {k: {"type": "patch", "idx": torch.from_numpy(v)} for k, v in geo.boundaries.items()}
    bcs = {
        V.U: {"inlets": BC(BC.Type.FIXED_VALUE, torch.tensor([20.0, 0.0, 0.0])),
              "walls": BC(BC.Type.FIXED_VALUE, torch.tensor([0.0, 0.0, 0.0]))},
        V.P: {"outlets": BC(BC.Type.FIXED_VALUE, torch.tensor([0.0]))},
    }
    md = ofles.OpenFOAMMetadata(file=Path("/synthetic/case/data.h5"), nu=1e-5, h=torch.ones(3),
                                cell_counts=np.array(geo.padded), cell_idx=torch.from_numpy(geo.cell_idx),
                                boundaries=bnd, boundary_conditions=bcs, holes=[])
    data = ofles.OpenFOAMData(md, torch.zeros(B), {V.U: torch.from_numpy(u), V.P: torch.from_numpy(p)})
    grid = data.grid_embedding((V.U, V.P))
    out["grid_embedding"] = grid.numpy()
    emb = CellTypeLearnedEmbedding(4)
    table = rng.standard_normal((6, 4)).astype(np.float32)
    with torch.no_grad():
        emb.embedding.weight.copy_(torch.from_numpy(table))
        out["cell_types"] = emb.cell_types(data).numpy()
        out["cell_type_embedding"] = emb(data).numpy()
    out["table"] = table
    other = torch.from_numpy(rng.standard_normal(grid.shape).astype(np.float32))
    out["other"] = other.numpy()
    out["where_cells"] = ref_utils.where_cells(md.cell_idx, grid, other).numpy()
    out["where_cells_zero"] = ref_utils.where_cells(md.cell_idx, other).numpy()
    out["select_cells"] = ref_utils.select_cells(other, md.cell_idx).numpy()
    np.savez_compressed(HERE / "grid.npz", **out)
    print("grid", grid.shape, np.bincount(out["cell_types"].ravel(), minlength=6))


def gen_shapes():
    """The FULL shapes configuration (BASELINE.json configs[1]/[2]: dim 32, 4 levels, 194x50x50 padded grid, u+p,
    55.2 M parameters) through the unmodified reference at B=1: denoiser output + per-block taps (checksums and strided
    sub-samples), one training loss with every parameter gradient (checksums + sub-samples), and a 3-step sampling
    chain.  Inputs: turbdiff_b200.synthetic.synthetic_inputs(1, 100) (numpy-seeded, no model code involved), weights
    oracle.unet_ref.synth_state_dict(spec, 0).  ~25 s of CPU time, ~6 GB of memory."""
    sys.path.insert(0, str(ROOT / "generative-turbulence_b200"))
    from turbdiff_b200.synthetic import synthetic_inputs

    spec = shapes_spec()
    sd = synth_state_dict(spec, SHAPES_SEED)
    geo, x, c_local = synthetic_inputs(1, SHAPES_INPUT_SEED)
    cell_idx = torch.from_numpy(geo.cell_idx)
    C = {Conditioning.Type.CELL_TYPE: c_local}
    t = torch.tensor([SHAPES_FWD_T], dtype=torch.long)
    out = {"t": t.numpy()}
    m = build_ref_model(spec, sd)
    taps, hooks = {}, []

    def grab(name):
        def hook(_mod, _inp, outp):
            taps[name] = outp.detach()
        return hook

    un = m.u_net
    for i, blk in enumerate(un.downsampling_blocks):
        hooks.append(blk.register_forward_hook(grab(f"down{i}")))
    for i, blk in enumerate(un.upsampling_blocks):
        hooks.append(blk.register_forward_hook(grab(f"up{i}")))
    for i in range(3):
        hooks.append(un.center_block[i].register_forward_hook(grab(f"center{i}")))
    hooks.append(m.decode[0].register_forward_hook(grab("decode0")))
    with torch.no_grad():
        y = m(x, t, C)
    for h in hooks:
        h.remove()
    out["out/sub"] = sub3(y).numpy()
    out["out/sum"] = np.array([y.double().sum().item(), y.double().pow(2).sum().item()])
    for k, v in taps.items():
        out[f"tap/{k}/sub"] = tap_sample(v).numpy()
        out[f"tap/{k}/sum"] = np.array([v.double().sum().item(), v.double().pow(2).sum().item()])
    print("shapes forward", tuple(y.shape), float(y.abs().mean()))

    gd = ref.GaussianDiffusion(m, timesteps=SHAPES_T, beta_schedule="log-snr-linear", loss_type="l2", noise_bcs=True)
    torch.manual_seed(77)
    s = gd.p_sample_loop(x, C, cell_idx, start_from=3)
    out["sample_from3/sub"] = sub3(s).numpy()
    out["sample_from3/sum"] = np.array([s.double().sum().item(), s.double().pow(2).sum().item()])
    print("shapes 3-step chain", float(s.abs().mean()))

    m.train()
    torch.manual_seed(4321)
    loss, tdraw = gd(x, C, _MD(cell_idx), None)
    loss.backward()
    out["loss"] = np.array(loss.item())
    out["loss_t"] = tdraw.numpy()
    for k, p in m.named_parameters():
        g = p.grad
        out[f"grad/{k}/sub"] = grad_sample(g).numpy()
        out[f"grad/{k}/sum"] = np.array([g.double().sum().item(), g.double().pow(2).sum().item()])
    print("shapes loss", loss.item(), "t", tdraw.tolist())
    np.savez_compressed(HERE / "shapes.npz", **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    todo = sys.argv[1:] or ["layout", "schedules", "time_embedding", "grid", "unet", "diffusion", "diffusion_lv", "tke", "shapes"]
    for name in todo:
        globals()[f"gen_{name}"]()
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size // 1024, "KiB")
