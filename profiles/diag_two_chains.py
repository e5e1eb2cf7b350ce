"""Experiment: ONE chain of B samples vs TWO independent chains of B/2 samples replayed on two streams (the
bandwidth-bound kernels of one chain can then overlap the tensor-bound convolutions of the other).

  python profiles/diag_two_chains.py [--batch 8] [--steps 20]
"""

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "generative-turbulence_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from turbdiff_b200 import DenoisingModel, GaussianDiffusion, _lib  # noqa: E402
from turbdiff_b200.models.utils import inside_mask  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
T = 1000
dev = torch.device("cuda", 0)


class Chain:
    def __init__(self, B, seed):
        torch.manual_seed(0)
        m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=T, dim=32, u_net_levels=4,
                           norm_type="group", precision="bf16").to(dev).eval()
        self.gd = GaussianDiffusion(m, timesteps=T, beta_schedule="log-snr-linear", noise_bcs=True).to(dev)
        geo, x, c_local = bench.synthetic_inputs(B, seed)
        self.x_bcs, cl = x.to(dev), c_local.to(dev)
        idx = torch.from_numpy(geo.cell_idx).to(dev)
        self.nvox = int(np.prod(geo.padded))
        self.mask, self.coef, self.eng = inside_mask(idx, self.nvox), self.gd._coef_table(dev), m.engine()
        self.st = self.eng.sampler_state(B, tuple(geo.padded), dev, cl)
        self.st["x_t"].copy_(torch.randn_like(self.x_bcs))
        self.st["t_dev"].fill_(500)
        self.st["t_vec"].fill_(500)
        self.state = [self.st["x_t"], self.st["x_t2"]]
        self.eng.encode_state(self.st, self.state[0])
        self.stream = torch.cuda.Stream(device=dev)

    def step(self):
        self.eng.forward_graphed(self.st, tail=True)
        z, zb = torch.randn_like(self.x_bcs), torch.randn_like(self.x_bcs)
        self.eng.step_tail(self.st, self.state[0], self.state[1], z, zb, self.x_bcs, self.mask, self.coef, self.st["t_dev"], _lib.STEP_NOISE_BCS)
        self.state.reverse()


def timed(chains, steps):
    main = torch.cuda.current_stream()
    for c in chains:
        for _ in range(3):
            with torch.cuda.stream(c.stream):
                c.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for c in chains:
        c.stream.wait_event(e0)
    for _ in range(steps):
        for c in chains:
            with torch.cuda.stream(c.stream):
                c.step()
    for c in chains:
        main.wait_stream(c.stream)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


one = Chain(a.batch, 100)
print(f"one chain  B={a.batch}: {timed([one], a.steps):.3f} ms per step of {a.batch} samples")
del one
torch.cuda.empty_cache()
two = [Chain(a.batch // 2, 100), Chain(a.batch // 2, 101)]
print(f"two chains B={a.batch // 2}+{a.batch // 2}: {timed(two, a.steps):.3f} ms per step of {a.batch} samples")
print(f"one of them alone: {timed(two[:1], a.steps):.3f} ms per step of {a.batch // 2} samples")
