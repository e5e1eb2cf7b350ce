"""Stand-alone launches of the streaming backward kernels at the training workload's level-0 / level-1 shapes (B = 4),
for ncu and for CUDA-event timing:

  python profiles/run_bwd_kernels.py [--reps 5]            # prints ms and GB/s per kernel
  ncu --set full --clock-control none --import-source on -k regex:trilinear_bwd -o gpurun_out/prof python profiles/run_bwd_kernels.py --reps 1
"""

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "generative-turbulence_b200")]

import torch  # noqa: E402

from turbdiff_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--only", default="")
a = ap.parse_args()
B = a.batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
_lib.load()
s = _lib.stream_ptr


def grid(size, C):
    X, Y, Z = size
    return (torch.randn((B, X + 2, Y + 2, Z + 2, C), device=dev) * 0.5).to(torch.bfloat16)


L0, L1 = (194, 50, 50), (97, 25, 25)
g0_128, g1_64, g1_128, g0_64 = grid(L0, 128), grid(L1, 64), grid(L1, 128), grid(L0, 64)
raw0, d0 = grid(L0, 64), grid(L0, 64)
stats = torch.zeros((B, 8, 2), dtype=torch.float64, device=dev)
_lib.call("tdb_gn_stats", raw0.data_ptr(), 64, stats.data_ptr(), B, *L0, 64, 8, 1, s())
gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
film = torch.randn((B, 128), device=dev) * 0.1
red = torch.zeros((B, 64, 4), dtype=torch.float64, device=dev)
grp = torch.zeros((B, 8, 2), dtype=torch.float32, device=dev)


def rows(size):
    return B * (size[0] + 2) * (size[1] + 2) * (size[2] + 2)


cases = {
    # name: (callable, algorithmic bytes)
    "trilinear_bwd up-adjoint 64ch L0->L1": (lambda: _lib.call("tdb_trilinear_bwd", g0_128.data_ptr(), 128, *L0, g1_64.data_ptr(), 64, *L1, B, 64, 1, 0, s()),
                                             rows(L0) * 128 + rows(L1) * 128),
    "trilinear_bwd down-adjoint+acc 64ch L1->L0": (lambda: _lib.call("tdb_trilinear_bwd", g1_64.data_ptr(), 64, *L1, g0_128.data_ptr() + 128, 128, *L0, B, 64, 1, 1, s()),
                                                   rows(L1) * 128 + 2 * rows(L0) * 128),
    "pw_bwd_reduce 64ch L0": (lambda: _lib.call("tdb_pointwise_bwd_reduce", g0_64.data_ptr(), 64, raw0.data_ptr(), 64, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                 film.data_ptr(), 128, red.data_ptr(), B, *L0, 64, 8, 1e-5, 1, 1, s()), 2 * rows(L0) * 128),
    "pw_bwd_apply 64ch L0": (lambda: _lib.call("tdb_pointwise_bwd_apply", g0_64.data_ptr(), 64, raw0.data_ptr(), 64, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                film.data_ptr(), 128, grp.data_ptr(), d0.data_ptr(), 64, B, *L0, 64, 8, 1e-5, 1, 1, s()), 3 * rows(L0) * 128),
    "halo_fold 128ch L0": (lambda: _lib.call("tdb_halo_fold", g0_128.data_ptr(), 128, B, *L0, 128, 1, s()), 0),
}
for name, (fn, nbytes) in cases.items():
    if a.only and a.only not in name:
        continue
    fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
    for e0, e1 in ev:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[len(ev) // 2]
    print(f"{name:48s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.0f} GB/s (algorithmic)")
