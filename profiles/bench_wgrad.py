"""Times the weight-gradient kernels on the layer shapes of the shapes config (B=4): old dispatch
(tdb_conv3d_wgrad) vs the tcgen05 kernel in its modes.  Usage: python profiles/bench_wgrad.py"""
import os, sys, json
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-turbulence_b200"))
from turbdiff_b200 import _lib

_lib.load()
B = 4
LAYERS = [  # X, Y, Z, Cin, Cout
    (194, 50, 50, 64, 64), (194, 50, 50, 128, 32), (194, 50, 50, 32, 32),
    (97, 25, 25, 64, 128), (97, 25, 25, 128, 128), (97, 25, 25, 256, 64),
    (49, 13, 13, 128, 256), (49, 13, 13, 256, 256), (49, 13, 13, 512, 128),
    (25, 7, 7, 256, 512), (25, 7, 7, 512, 512), (25, 7, 7, 1024, 256),
    (13, 4, 4, 512, 512),
]
if os.environ.get("WGRAD_LAYERS"):
    LAYERS = [LAYERS[int(i)] for i in os.environ["WGRAD_LAYERS"].split(",")]
MODES = [("tc0", 0), ("tc1", 1), ("tc3", 3)]
if os.environ.get("WGRAD_MODES"):
    MODES = [(f"tc{m}", int(m)) for m in os.environ["WGRAD_MODES"].split(",")]
res = []
for (X, Y, Z, Cin, Cout) in LAYERS:
    rows = B * (X + 2) * (Y + 2) * (Z + 2)
    x = (torch.randn(rows, Cin, device="cuda") * 0.5).bfloat16()
    dy = torch.zeros((B, X + 2, Y + 2, Z + 2, Cout), device="cuda", dtype=torch.bfloat16)
    dy[:, 1:-1, 1:-1, 1:-1] = torch.randn(B, X, Y, Z, Cout, device="cuda").bfloat16()
    dws = {}
    row = {"layer": f"{Cin}->{Cout} @{X}x{Y}x{Z}", "gflop": 2 * 27 * Cin * Cout * B * X * Y * Z / 1e9}
    for name, mode in MODES:
        dw = torch.zeros((27, Cin, Cout), dtype=torch.float32, device="cuda")
        def run():
            if mode is None:
                _lib.call("tdb_conv3d_wgrad", x.data_ptr(), Cin, dy.data_ptr(), Cout, dw.data_ptr(), B, X, Y, Z, Cin, Cout, 27, 1,
                          _lib.WGRAD_ZERO_HALO, _lib.stream_ptr())
            else:
                _lib.call("tdb_conv3d_wgrad_tc", x.data_ptr(), Cin, dy.data_ptr(), Cout, dw.data_ptr(), B, X, Y, Z, Cin, Cout, 27, mode,
                          _lib.stream_ptr())
        try:
            run()
            torch.cuda.synchronize()
            dws[name] = dw.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            row[name + "_ms"] = round(ms, 4)
            row[name + "_tflops"] = round(row["gflop"] / ms, 1)
        except Exception as ex:  # noqa
            row[name + "_err"] = str(ex)[:200]
    ref = dws.get("tc0")
    for k, v in dws.items():
        if ref is not None and k != "tc0":
            row[k + "_rel_vs_tc0"] = float((v - ref).norm() / ref.norm())
    print(json.dumps(row), flush=True)
    res.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_wgrad.json"), "w"), indent=1)
