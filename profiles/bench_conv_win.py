"""Times the level-0 convolutions (B=4) on the kz-folded kernels vs the row-window kernel."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-turbulence_b200"))
from turbdiff_b200 import _lib
_lib.load()
B = int(os.environ.get("B", 4))
res = []
LAYERS = [(194, 50, 50, 64, 64, "fold2"), (194, 50, 50, 128, 32, "fold2"), (194, 50, 50, 32, 32, "fold"), (194, 50, 50, 32, 128, "v1"),
          (97, 25, 25, 64, 128, "fold2"), (97, 25, 25, 128, 128, "fold2"), (97, 25, 25, 64, 64, "fold2"),
          (48, 12, 12, 128, 256, "fold2"), (48, 12, 12, 256, 256, "fold2"), (48, 12, 12, 512, 128, "fold2"),
          (24, 6, 6, 256, 512, "fold2"), (24, 6, 6, 512, 512, "fold2"), (24, 6, 6, 1024, 256, "fold2")]
ONLY = os.environ.get("ONLY")  # e.g. ONLY=32-32 restricts the run to one layer shape (for ncu)
for (X, Y, Z, Cin, Cout, old) in LAYERS:
    if ONLY and ONLY != f"{Cin}-{Cout}":
        continue
    pad = (Y + 2) * (Z + 2) + 2 * (Z + 2) + 256
    rows = B * (X + 2) * (Y + 2) * (Z + 2)
    buf = torch.zeros((rows + 2 * pad, Cin), device="cuda", dtype=torch.bfloat16)
    buf[pad:pad + rows] = (torch.randn(rows, Cin, device="cuda") * 0.5).bfloat16()
    xin = buf[pad:pad + rows]
    out = torch.zeros((rows, Cout), device="cuda", dtype=torch.bfloat16)
    wf = (torch.randn(3 * Cout, 9 * Cin, device="cuda") * 0.02).bfloat16()
    wk = (torch.randn(Cout, 27 * Cin, device="cuda") * 0.02).bfloat16()
    bias = torch.zeros(Cout, device="cuda")
    stats = torch.zeros((B, 8, 2), dtype=torch.float64, device="cuda")
    s = _lib.stream_ptr
    def run_old():
        if old == "fold2":
            _lib.call("tdb_conv3d_bf16_fold2", xin.data_ptr(), Cin, pad, wf.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z,
                      Cin, Cout, stats.data_ptr(), 8, 0, None, None, None, 0, s())
        elif old == "fold":
            _lib.call("tdb_conv3d_bf16_fold", xin.data_ptr(), Cin, pad, wf.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z,
                      Cin, Cout, stats.data_ptr(), 8, 0, s())
        else:
            _lib.call("tdb_conv3d_bf16", xin.data_ptr(), Cin, wk.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
                      27, None, 8, 1, None, s())
    def run_win():
        _lib.call("tdb_conv3d_bf16_win", xin.data_ptr(), Cin, wk.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
                  None if old == "v1" else stats.data_ptr(), 8, 1 if old == "v1" else 0, None, None, None, 0, s())
    def run_winz():
        _lib.call("tdb_conv3d_bf16_winz", xin.data_ptr(), Cin, wf.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
                  None if old == "v1" else stats.data_ptr(), 8, 1 if old == "v1" else 0, None, None, None, 0, s())
    row = {"layer": f"{Cin}->{Cout} @{X}", "gflop": 2 * 27 * Cin * Cout * B * X * Y * Z / 1e9, "old_kernel": old}
    def run_winp():
        _lib.call("tdb_conv3d_bf16_winp", xin.data_ptr(), Cin, wf.data_ptr(), bias.data_ptr(), out.data_ptr(), Cout, B, X, Y, Z, Cin, Cout,
                  stats.data_ptr(), 8, 0, s())
    cands = [("old", run_old), ("win", run_win)]
    if Cin == 32 and Cout == 32:
        cands.append(("winp", run_winp))
    if Cout in (32, 64) and 27 * Cin * Cout <= 116 * 1024:
        cands.append(("winz", run_winz))
    for name, fn in cands:
        try:
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            row[name + "_ms"] = round(ms, 4); row[name + "_tflops"] = round(row["gflop"] / ms, 1)
        except Exception as ex:  # noqa
            row[name + "_err"] = str(ex)[:300]
            continue
    print(json.dumps(row), flush=True)
    res.append(row)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_conv_win.json"), "w"), indent=1)
