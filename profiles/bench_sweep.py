"""BASELINE.json configs[4]: 3x3x3 replicate-padded convolution sweep over (Cin, Cout) in {64,128,256,512}^2 plus the 22
model shapes, on the grids 24^3, 48x12x12, 97x25x25 and 194x50x50, and the bottleneck attention (B*4 heads, S in
{108, 128}, d = 32), each against the tensor-core roofline (MEASURED_PEAKS.json, burst figure: kernels timed alone).

Every convolution goes through the product's own dispatch (DenoiserEngine.fold_kind / _conv / pack_conv), so the
table shows the kernel a model of that shape would actually run.  One JSON line per case; the whole table is also
written to gpurun_out/sweep.json.

    python profiles/bench_sweep.py [--batch 4] [--quick]
"""

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "generative-turbulence_b200")]

import torch  # noqa: E402

from turbdiff_b200 import _lib  # noqa: E402
from turbdiff_b200.engine import DenoiserEngine, View  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--quick", action="store_true")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
B = a.batch
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"bf16_tflops": 1590.0}
PEAK = peaks["bf16_tflops"]

GRIDS = {"24^3": (24, 24, 24), "48x12x12": (48, 12, 12), "97x25x25": (97, 25, 25), "194x50x50": (194, 50, 50)}
MODEL = [(64, 64, "194x50x50"), (64, 128, "97x25x25"), (128, 128, "97x25x25"), (128, 256, "48x12x12"), (256, 256, "48x12x12"),
         (256, 512, "24x6x6"), (512, 512, "24x6x6"), (512, 512, "12x3x3"), (1024, 256, "24x6x6"), (256, 256, "24x6x6"),
         (512, 128, "48x12x12"), (128, 128, "48x12x12"), (256, 64, "97x25x25"), (64, 64, "97x25x25"), (128, 32, "194x50x50"),
         (32, 32, "194x50x50")]
GRIDS.update({"24x6x6": (24, 6, 6), "12x3x3": (12, 3, 3)})
SQUARE = [(ci, co) for ci in (64, 128, 256, 512) for co in (64, 128, 256, 512)]


class _FakeModel:
    """Just enough of DenoisingModel for DenoiserEngine's kernel selection (no parameters are touched)."""

    u_net_levels = 4


def engine():
    eng = DenoiserEngine.__new__(DenoiserEngine)
    eng.model = _FakeModel()
    eng.precision, eng.dt, eng.tdtype = "bf16", _lib.BF16, torch.bfloat16
    eng.fold = eng.fold2 = eng.win = eng.fold_wide = eng.fuse_proj = eng.win_center = True
    eng._level_zp = {}
    return eng


def time_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_case(eng, cin, cout, gname, level=1):
    X, Y, Z = GRIDS[gname]
    eng._level_zp = {level: Z + 2}
    rows = B * (X + 2) * (Y + 2) * (Z + 2)
    if rows * max(cin, cout) * 2 > 6e9:
        return None
    pad = eng.pad_rows((X, Y, Z))

    def grid(C):
        flat = torch.zeros((rows + 2 * pad, C), dtype=torch.bfloat16, device="cuda")
        flat[pad : pad + rows] = (torch.randn(rows, C, device="cuda") * 0.5).bfloat16()
        return View(flat[pad : pad + rows].view(B, X + 2, Y + 2, Z + 2, C), 0, C, level)

    x, out = grid(cin), grid(cout)
    w = eng.pack_conv(torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.02, level)
    bias = torch.zeros(cout, device="cuda")
    stats = torch.zeros(B * 8 * 2, dtype=torch.float64, device="cuda")
    p = {"B": B, "sizes": {level: (X, Y, Z)}, "splitk": torch.zeros(rows * min(cout, 1024), dtype=torch.float32, device="cuda")}
    kind = eng.fold_kind(27, cin, cout, level)
    if kind == "fold" and cin == 32 and cout == 32 and (Z + 2) % 2 == 0:
        kind = "winp"
    fused = kind is not None and (cout // 8) % 2 == 0
    fn = lambda: eng._conv(p, x, w, bias, out, 27, stats if fused else None, 8)  # noqa: E731
    ms = time_ms(fn, a.reps)
    gflop = 2.0 * 27 * cin * cout * B * X * Y * Z / 1e9
    return {"op": "conv3x3x3", "cin": cin, "cout": cout, "grid": gname, "batch": B, "kernel": kind or "tc (per tap)", "ms": round(ms, 4),
            "tflops": round(gflop / ms, 1), "frac_of_measured_burst_peak": round(gflop / ms / PEAK, 3)}


def attention_case(S_shape, Bh):
    X, Y, Z = S_shape
    heads, dh = 4, 32
    hid = heads * dh
    qkv = (torch.randn(Bh, X + 2, Y + 2, Z + 2, 3 * hid, device="cuda")).bfloat16()
    out = torch.zeros(Bh, X + 2, Y + 2, Z + 2, hid, device="cuda", dtype=torch.bfloat16)
    fn = lambda: _lib.call("tdb_attention", qkv.data_ptr(), 3 * hid, out.data_ptr(), hid, Bh, X, Y, Z, heads, dh, _lib.BF16, _lib.stream_ptr())  # noqa: E731
    ms = time_ms(fn, 20)
    S = X * Y * Z
    gflop = 4.0 * Bh * heads * S * S * dh / 1e9
    return {"op": "attention", "S": S, "batch": Bh, "heads": heads, "dh": dh, "kernel": "tcgen05" if S <= 128 else "simt", "ms": round(ms, 4),
            "tflops": round(gflop / ms, 3), "note": "latency-bound: B*heads CTAs of one 128x128x32 + one 128x32x128 MMA chain"}


def main():
    _lib.load()
    eng = engine()
    res = []
    cases = [(ci, co, g) for (ci, co, g) in MODEL]
    if not a.quick:
        for gname in ("24^3", "48x12x12", "97x25x25", "194x50x50"):
            cases += [(ci, co, gname) for ci, co in SQUARE]
    seen = set()
    for ci, co, g in cases:
        if (ci, co, g) in seen:
            continue
        seen.add((ci, co, g))
        try:
            row = conv_case(eng, ci, co, g, level=4 if g == "12x3x3" else 1)
        except Exception as e:  # a shape the dispatch cannot take is a table entry, not a crash
            row = {"op": "conv3x3x3", "cin": ci, "cout": co, "grid": g, "error": str(e)[:160]}
        if row is None:
            continue
        print(json.dumps(row), flush=True)
        res.append(row)
        torch.cuda.empty_cache()
    for shape, Bh in (((12, 3, 3), 8), ((12, 3, 3), 64), ((8, 4, 4), 8), ((8, 4, 4), 64), ((8, 8, 8), 8)):
        row = attention_case(shape, Bh)
        print(json.dumps(row), flush=True)
        res.append(row)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "sweep.json").write_text(json.dumps({"peak_tflops_burst": PEAK, "rows": res}, indent=1))


if __name__ == "__main__":
    main()
