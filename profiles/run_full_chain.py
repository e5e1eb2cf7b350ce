"""One COMPLETE ancestral-sampling chain of the bench workload (T = 1000, B = 8) through the public API with pinned host buffers:
the un-extrapolated end-to-end number behind bench.py's `e2e` (which times 16-step chains and scales by T/16).

    python profiles/run_full_chain.py [--batch 8] [--timesteps 1000]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "generative-turbulence_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
from turbdiff_b200 import DenoisingModel, GaussianDiffusion  # noqa: E402
from turbdiff_b200.models.conditioning import Conditioning  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--timesteps", type=int, default=1000)
a = ap.parse_args()
T, B = a.timesteps, a.batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=T, dim=32, u_net_levels=4,
                   norm_type="group", precision="bf16").to(dev).eval()
gd = GaussianDiffusion(m, timesteps=T, beta_schedule="log-snr-linear", noise_bcs=True).to(dev)
geo, x, c_local = bench.synthetic_inputs(B, 100)
x_pinned = x.pin_memory()
out_pinned = torch.empty_like(x).pin_memory()
C = {Conditioning.Type.CELL_TYPE: c_local.to(dev)}
idx = torch.from_numpy(geo.cell_idx).to(dev)
gd.p_sample_loop(x_pinned.to(dev), C, idx, start_from=4)  # warm-up: plan, weights, graph
torch.cuda.synchronize()
mem0 = torch.cuda.memory_allocated()
t0 = time.perf_counter()
s = gd.p_sample_loop(x_pinned.to(dev, non_blocking=True), C, idx)
out_pinned.copy_(s, non_blocking=True)
torch.cuda.synchronize()
sec = time.perf_counter() - t0
ok = bool(torch.isfinite(out_pinned).all())
print(json.dumps({"what": "full chain through GaussianDiffusion.p_sample_loop, pinned host in/out", "T": T, "batch": B, "seconds": sec,
                  "samples_per_sec": B / sec, "ms_per_step": sec * 1e3 / T, "finite": ok, "sample_abs_max": float(out_pinned.abs().max()),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "mem_growth_mb": (torch.cuda.memory_allocated() - mem0) / 1e6}))
