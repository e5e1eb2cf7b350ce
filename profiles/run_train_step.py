"""One training step of the bench workload (B = 4: q_sample + U-Net forward + backward + clip + RAdam) between
cudaProfilerStart/Stop, on the eager launch programs, for ncu:

  TURBDIFF_B200_TRAIN_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/train_launches.csv python profiles/run_train_step.py
"""
import os
import sys
from pathlib import Path

os.environ.setdefault("TURBDIFF_B200_TRAIN_GRAPH", "0")
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "generative-turbulence_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
from turbdiff_b200 import DenoisingModel, GaussianDiffusion  # noqa: E402
from turbdiff_b200.models.conditioning import Conditioning  # noqa: E402
from turbdiff_b200.optim import FusedRAdam  # noqa: E402

T, B = 1000, int(os.environ.get("B", 4))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=T, dim=32, u_net_levels=4,
                   norm_type="group", precision="bf16").to(dev).train()
gd = GaussianDiffusion(m, timesteps=T, beta_schedule="log-snr-linear", noise_bcs=True).to(dev)
geo, x, c_local = bench.synthetic_inputs(B, 100)
x = x.to(dev)
C = {Conditioning.Type.CELL_TYPE: c_local.to(dev)}


class MD:
    cell_idx = torch.from_numpy(geo.cell_idx).to(dev)


opt = FusedRAdam(m.parameters(), lr=1e-4, max_grad_norm=0.1)


def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = gd(x, C, MD, None)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
