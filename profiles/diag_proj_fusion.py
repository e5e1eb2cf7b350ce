"""Is the fused 1x1 residual projection worth it on the streamed row-window kernel?  Per layer (B = 8): the 3x3x3 convolution
alone, with the fused projection, and the stand-alone 1x1 projection (per-tap kernel).  python profiles/diag_proj_fusion.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-turbulence_b200"))
from turbdiff_b200 import _lib
_lib.load()
B = int(os.environ.get("B", 8))
s = _lib.stream_ptr
for (X, Y, Z, Cin, Cout) in [(97, 25, 25, 64, 128), (48, 12, 12, 128, 256), (24, 6, 6, 256, 512), (24, 6, 6, 1024, 256), (48, 12, 12, 512, 128)]:
    rows = B * (X + 2) * (Y + 2) * (Z + 2)
    xin = (torch.randn(rows, Cin, device="cuda") * 0.5).bfloat16()
    out = torch.zeros((rows, Cout), device="cuda", dtype=torch.bfloat16)
    outp = torch.zeros((rows, Cout), device="cuda", dtype=torch.bfloat16)
    wk = (torch.randn(Cout, 27 * Cin, device="cuda") * 0.02).bfloat16()
    wp = (torch.randn(Cout, Cin, device="cuda") * 0.02).bfloat16()
    bias = torch.zeros(Cout, device="cuda")
    stats = torch.zeros((B, 8, 2), dtype=torch.float64, device="cuda")
    def plain():
        _lib.call("tdb_conv3d_bf16_win", xin.data_ptr(), Cin, wk.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
                  stats.data_ptr(), 8, 0, None, None, None, 0, s())
    def fused():
        _lib.call("tdb_conv3d_bf16_win", xin.data_ptr(), Cin, wk.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
                  stats.data_ptr(), 8, 0, wp.data_ptr(), bias.data_ptr(), outp.data_ptr(), Cout, s())
    def proj():
        _lib.call("tdb_conv3d_bf16", xin.data_ptr(), Cin, wp.data_ptr(), bias.data_ptr(), outp.data_ptr(), Cout, B, X, Y, Z, Cin, Cout, 1,
                  None, 8, 0, None, s())
    row = {"layer": f"{Cin}->{Cout} @{X}x{Y}x{Z} B={B}"}
    for name, fn in (("plain", plain), ("fused", fused), ("proj_alone", proj)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        row[name + "_ms"] = round(e0.elapsed_time(e1) / 10, 4)
    print(json.dumps(row), flush=True)
