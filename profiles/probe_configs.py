"""Exploration (not a test): forward + backward of unusual configurations on the GPU against the CPU oracle.
Prints one line per (configuration, precision): worst forward tap error, worst gradient error, or the exception raised."""
import sys, traceback
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "generative-turbulence_b200"), str(ROOT / "tests")]
from oracle.cases import case_inputs  # noqa: E402
from oracle.unet_ref import UNetSpec, denoiser_forward, synth_state_dict  # noqa: E402
from turbdiff_b200 import DenoisingModel  # noqa: E402
from turbdiff_b200.models.conditioning import Conditioning  # noqa: E402
from util import rel_l2  # noqa: E402

NORM = {8: "group", 1: "layer", None: "instance"}
CONFIGS = {
    "dim48-L2": dict(spec=UNetSpec(dim=48, u_net_levels=2, timesteps=10), cells=(14, 8, 8), hole=None, batch=2, seed=1),
    "dim16-L3": dict(spec=UNetSpec(dim=16, u_net_levels=3, timesteps=10), cells=(30, 14, 12), hole=None, batch=1, seed=2),
    "dim32-L1": dict(spec=UNetSpec(dim=32, u_net_levels=1, timesteps=10), cells=(10, 6, 6), hole=None, batch=2, seed=3),
    "dim80-L2": dict(spec=UNetSpec(dim=80, u_net_levels=2, timesteps=10), cells=(14, 8, 8), hole=None, batch=1, seed=4),
    "dim32-L2-smallgrid": dict(spec=UNetSpec(dim=32, u_net_levels=2, timesteps=10), cells=(7, 3, 3), hole=None, batch=2, seed=5),
    "dim32-L2-heads2": dict(spec=UNetSpec(dim=32, u_net_levels=2, timesteps=10, heads=2), cells=(14, 8, 8), hole=None, batch=2, seed=6),
    "dim24-L2-fp32only": dict(spec=UNetSpec(dim=24, u_net_levels=2, timesteps=10), cells=(14, 8, 8), hole=None, batch=2, seed=7),
}


def run(name, case, precision):
    spec = case["spec"]
    m = DenoisingModel(in_features=spec.in_features, out_features=spec.out_features, c_local_features=spec.c_local_features,
                       c_global_features=0, timesteps=spec.timesteps, dim=spec.dim, u_net_levels=spec.u_net_levels,
                       norm_type=NORM[spec.groups], precision=precision)
    if spec.heads != 4:
        return "skipped: heads is not a constructor argument of the reference"
    m.load_state_dict(synth_state_dict(spec, case["seed"]), strict=True)
    m = m.cuda().train()
    x, t, c_local, _ = case_inputs(case)
    G = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))
    sd = {k: v.requires_grad_() for k, v in synth_state_dict(spec, case["seed"], torch.float64).items()}
    clr = c_local.double().requires_grad_()
    taps_ref = {}
    want = denoiser_forward(sd, spec, x.double(), t, clr, taps_ref)
    (want * G.double()).sum().backward()
    cl = c_local.cuda().requires_grad_()
    eps = m(x.cuda(), t.cuda(), {Conditioning.Type.CELL_TYPE: cl})
    (eps * G.cuda()).sum().backward()
    fwd = rel_l2(eps.detach(), want.detach())
    errs = {k: rel_l2(p.grad, sd[k].grad) for k, p in m.named_parameters() if float(sd[k].grad.abs().max()) >= 1e-9}
    errs["c_local"] = rel_l2(cl.grad, clr.grad)
    worst = max(errs.items(), key=lambda kv: kv[1])
    return f"fwd {fwd:.2e}  worst grad {worst[1]:.2e} ({worst[0]})  fallbacks {m.engine().graph_fallbacks}"


for name, case in CONFIGS.items():
    for precision in ("fp32", "bf16"):
        if "fp32only" in name and precision == "bf16":
            pass  # expected to raise: report what it says
        try:
            print(f"{name:22s} {precision}: {run(name, case, precision)}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{name:22s} {precision}: RAISED {type(e).__name__}: {str(e)[:300]}", flush=True)
        torch.cuda.synchronize()
