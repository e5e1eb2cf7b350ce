"""One sampling step of the bench workload between cudaProfilerStart/Stop, for ncu - issued exactly like
GaussianDiffusion.p_sample_loop issues it (engine sampler state, eager launch program, fused step tail):

  ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/run_step.py [--batch 8] [--precision bf16]
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
ap.add_argument("--precision", default="bf16")
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
T, B = 1000, a.batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=T, dim=32, u_net_levels=4,
                   norm_type="group", precision=a.precision)
m = m.to(dev).eval()
gd = GaussianDiffusion(m, timesteps=T, beta_schedule="log-snr-linear", noise_bcs=True).to(dev)
geo, x, c_local = bench.synthetic_inputs(B, 100)
x_bcs, cl = x.to(dev), c_local.to(dev)
idx = torch.from_numpy(geo.cell_idx).to(dev)
nvox = int(np.prod(geo.padded))
mask, coef, eng = inside_mask(idx, nvox), gd._coef_table(dev), m.engine()
eng.use_graph = False
st = eng.sampler_state(B, tuple(geo.padded), dev, cl)
st["x_t"].copy_(torch.randn_like(x_bcs))
st["t_dev"].fill_(500)
st["t_vec"].fill_(500)
fused = eng.can_fuse_tail()
state = [st["x_t"], st["x_t2"]]
if fused:
    eng.encode_state(st, state[0])


def step():
    eps = eng.forward_graphed(st, tail=fused)
    z, zb = torch.randn_like(x_bcs), torch.randn_like(x_bcs)
    if fused:
        eng.step_tail(st, state[0], state[1], z, zb, x_bcs, mask, coef, st["t_dev"], _lib.STEP_NOISE_BCS)
        state.reverse()
    else:
        _lib.call("tdb_ddpm_step", state[0].data_ptr(), eps.data_ptr(), z.data_ptr(), zb.data_ptr(), x_bcs.data_ptr(), mask.data_ptr(),
                  coef.data_ptr(), st["t_dev"].data_ptr(), state[0].data_ptr(), B, 4, nvox, _lib.STEP_NOISE_BCS, _lib.stream_ptr())


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
