"""Diagnostic: per-tap error of the tcgen05 weight-gradient modes on a small case."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-turbulence_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from turbdiff_b200 import _lib
from util import to_halo
import torch.nn.functional as F
_lib.load()
for (B, X, Y, Z, Cin, Cout) in [(2, 12, 6, 5, 64, 64), (2, 9, 7, 6, 32, 32), (1, 10, 6, 6, 128, 32)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, X, Y, Z, generator=g).cuda().bfloat16().float()
    dy = torch.randn(B, Cout, X, Y, Z, generator=g).cuda().bfloat16().float()
    xin = to_halo(x, dtype=torch.bfloat16)
    dyh = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda", dtype=torch.bfloat16)
    dyh[:, 1:-1, 1:-1, 1:-1, :] = dy.permute(0, 2, 3, 4, 1).bfloat16()
    w = torch.zeros(Cout, Cin, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv3d(F.pad(x.double().cpu(), (1,) * 6, mode="replicate"), w), w, dy.double().cpu())
    ref = ref.permute(2, 3, 4, 1, 0).reshape(27, Cin, Cout)
    for mode in (0, 1):
        dw = torch.zeros((27, Cin, Cout), dtype=torch.float32, device="cuda")
        _lib.call("tdb_conv3d_wgrad_tc", xin.data_ptr(), Cin, dyh.data_ptr(), Cout, dw.data_ptr(), B, X, Y, Z, Cin, Cout, 27, mode,
                  _lib.stream_ptr())
        torch.cuda.synchronize()
        d = dw.double().cpu()
        errs = [float((d[t] - ref[t]).norm() / ref[t].norm()) for t in range(27)]
        print(f"{Cin}->{Cout} mode {mode}: per-tap rel err (kz fastest):", " ".join(f"{e:.1e}" for e in errs), flush=True)
        if mode != 0:
            # which tap does each computed tap resemble most?
            for t in (0, 1, 2):
                best = min(range(27), key=lambda u: float((d[t] - ref[u]).norm()))
                print(f"   tap {t} closest to ref tap {best} (err {float((d[t]-ref[best]).norm()/ref[best].norm()):.2e})")
            # per-ci error pattern for tap 1
            e_ci = ((d[1] - ref[1]).norm(dim=1) / ref[1].norm(dim=1))
            print("   tap1 per-ci err:", " ".join(f"{float(v):.0e}" for v in e_ci[:16]))
