"""Diagnostic: the streamed-weight row-window kernel (tdb_conv3d_bf16_win) over grid geometries, to separate the
effect of the grid (Z + 2 odd/even, rows per plane, halo fraction) from the channel configuration.

    python profiles/diag_win.py            -> one JSON line per case
"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-turbulence_b200"))
from turbdiff_b200 import _lib  # noqa: E402
_lib.load()

CASES = [
    # (B, X, Y, Z, Cin, Cout, proj)
    (8, 97, 25, 25, 64, 128, 0), (8, 97, 25, 25, 64, 128, 1), (1, 194, 50, 50, 64, 128, 0), (8, 96, 24, 24, 64, 128, 0),
    (8, 97, 25, 26, 64, 128, 0), (8, 97, 25, 30, 64, 128, 0), (2, 194, 50, 50, 64, 128, 0),
    (8, 97, 25, 25, 128, 128, 0), (8, 48, 12, 12, 128, 256, 0), (8, 48, 12, 12, 128, 256, 1), (8, 48, 12, 12, 256, 256, 0),
    (8, 24, 6, 6, 256, 512, 0), (8, 24, 6, 6, 512, 512, 0), (8, 12, 3, 3, 512, 512, 0), (8, 24, 6, 6, 1024, 256, 1), (8, 24, 6, 6, 256, 256, 0),
]
only = os.environ.get("ONLY")
for i, (B, X, Y, Z, Cin, Cout, proj) in enumerate(CASES):
    if only and str(i) not in only.split(","):
        continue
    pad = (Y + 2) * (Z + 2) + 2 * (Z + 2) + 256
    rows = B * (X + 2) * (Y + 2) * (Z + 2)
    buf = torch.zeros((rows + 2 * pad, Cin), device="cuda", dtype=torch.bfloat16)
    buf[pad:pad + rows] = (torch.randn(rows, Cin, device="cuda") * 0.5).bfloat16()
    xin = buf[pad:pad + rows]
    out = torch.zeros((rows, Cout), device="cuda", dtype=torch.bfloat16)
    outp = torch.zeros((rows, Cout), device="cuda", dtype=torch.bfloat16)
    wk = (torch.randn(Cout, 27 * Cin, device="cuda") * 0.02).bfloat16()
    wp = (torch.randn(Cout, Cin, device="cuda") * 0.02).bfloat16()
    bias = torch.zeros(Cout, device="cuda")
    stats = torch.zeros((B, 8, 2), dtype=torch.float64, device="cuda")

    def run():
        _lib.call("tdb_conv3d_bf16_win", xin.data_ptr(), Cin, wk.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
                  stats.data_ptr(), 8, 0, wp.data_ptr() if proj else None, bias.data_ptr() if proj else None,
                  outp.data_ptr() if proj else None, Cout, _lib.stream_ptr())

    run(); torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            run()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gflop = 2 * 27 * Cin * Cout * B * X * Y * Z / 1e9
    gflop_rows = 2 * 27 * Cin * Cout * rows / 1e9
    print(json.dumps({"case": i, "B": B, "grid": [X, Y, Z], "cin": Cin, "cout": Cout, "proj": proj, "us": round(ms * 1e3, 1),
                      "tflops_alg": round(gflop / ms, 1), "tflops_rows": round(gflop_rows / ms, 1), "row_eff": round(gflop / gflop_rows, 3)}), flush=True)
    del buf, out, outp
    torch.cuda.empty_cache()
