"""ctypes binding of libturbdiff_b200.so (the C ABI declared in include/turbdiff_b200.h).

The library is built in-tree by ``generative-turbulence_b200/build.py``.  A missing library
is a hard error: the product path has no fallback."""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "libturbdiff_b200.so"

F32, BF16 = 0, 1
PW_SILU, PW_NOHALO = 1, 2
STEP_NOISE_BCS, STEP_CLIP, STEP_FINAL, STEP_LEARNED_VAR = 1, 2, 4, 8
CONV_ALL_ROWS = 1
TRILINEAR_LINE = 0x100  # tdb_trilinear: dtype | TRILINEAR_LINE forces the two-stage up-sampling kernel (tests)
CONV_CLUSTER_MC = 2
WGRAD_ZERO_HALO = 1
TRIBWD_ACCUMULATE = 1

_p, _i, _l, _u, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_float, C.c_double

class PackJob(C.Structure):
    """TdbPackJob of include/turbdiff_b200.h (tdb_pack_conv_weights_batch)."""

    _fields_ = [("w", _p), ("dst", _p), ("Cout", _i), ("Cin", _i), ("taps", _i), ("folded", _i), ("tile_n", _i), ("transpose", _i)]


# name -> argument types (all return int unless noted)
SIGNATURES = {
    "tdb_encode_input": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_decode_output": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_conv3d_f32": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_conv3d_bf16": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _u, _p, _p],
    "tdb_conv3d_bf16_fold": [_p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _u, _p],
    "tdb_conv3d_bf16_fold2": [_p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _u, _p, _p, _p, _i, _p],
    "tdb_conv3d_bf16_win": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _u, _p, _p, _p, _i, _p],
    "tdb_conv3d_bf16_win_add1x1": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _u, _p, _i, _p, _p],
    "tdb_conv3d_bf16_winz": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _u, _p, _p, _p, _i, _p],
    "tdb_conv3d_bf16_winp": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _u, _p],
    "tdb_pack_conv_weights": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "tdb_pack_conv_weights_batch": [_p, _i, _p],
    "tdb_unpack_wgrad": [_p, _p, _i, _i, _i, _p],
    "tdb_gn_stats": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_pointwise": [_p, _i, _p, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _f, _u, _i, _p],
    "tdb_trilinear": [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_attention": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_time_film": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "tdb_time_film_bwd": [_p] * 17 + [_i, _i, _i, _p],
    "tdb_ddpm_step": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _u, _p],
    "tdb_step_tail": [_p, _i, _p, _p, _p, _p, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _u, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p],
    "tdb_q_sample": [_p, _p, _p, _p, _p, _p, _i, _i, _l, _i, _p],
    "tdb_masked_loss": [_p, _p, _p, _p, _p, _i, _i, _l, _l, _i, _p],
    "tdb_halo_fold": [_p, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_pointwise_bwd_reduce": [_p, _i, _p, _i, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _f, _u, _i, _p],
    "tdb_pointwise_bwd_finalize": [_p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p],
    "tdb_pointwise_bwd_apply": [_p, _i, _p, _i, _p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _u, _i, _p],
    "tdb_pointwise_bwd_apply_fused": [_p, _i, _p, _i, _p, _p, _p, _p, _i, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _u, _i, _p],
    "tdb_conv3d_wgrad": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _u, _p],
    "tdb_conv3d_wgrad_tc": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _u, _p],
    "tdb_trilinear_bwd": [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _u, _p],
    "tdb_attention_bwd": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_grad_sqnorm": [_p, _p, _p, _p, _i, _i, _p, _p],
    "tdb_radam_step": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _f, _f, _d, _d, _f, _f, _i, _p],
    "tdb_cl_nc_outer": [_p, _i, _p, _l, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "tdb_scatter_normalize": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _p],
    "tdb_gather_denormalize": [_p, _p, _p, _p, _p, _i, _i, _l, _l, _p],
    "tdb_where_cells": [_p, _p, _p, _p, _l, _l, _p],
    "tdb_select_cells": [_p, _p, _p, _l, _l, _l, _p],
    "tdb_scatter_cells": [_p, _p, _p, _i, _i, _l, _l, _p],
    "tdb_build_mask": [_p, _p, _l, _l, _p],
    "tdb_tke_spectrum": [_p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _i, _p, _p, _p],
}
OTHER = {"tdb_last_error": ([], C.c_char_p), "tdb_version": ([], _i), "tdb_launch_count": ([], _l)}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python generative-turbulence_b200/build.py` "
                "(turbdiff_b200 has no CPU or PyTorch fallback)"
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = _i
        for name, (args, res) in OTHER.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().tdb_last_error().decode(errors="replace")
        raise RuntimeError(f"libturbdiff_b200 {what} failed (code {rc}): {msg}")


# Optional per-kernel timing (bench.py): name -> list of (start, end) CUDA events recorded on the
# launching stream around each C-ABI call.  None = off (the normal case: zero overhead).
PROFILE: dict | None = None

# TURBDIFF_B200_DEBUG_CAPTURE=1: after every C-ABI call made while a CUDA graph is being captured, ask the runtime
# whether the capture is still valid and name the first call after which it is not (diagnostic).
import os as _os

_DEBUG_CAPTURE = _os.environ.get("TURBDIFF_B200_DEBUG_CAPTURE", "0") == "1"
_capture_was = [False]


def _capture_probe(name):
    try:
        now = torch.cuda.is_current_stream_capturing()
    except Exception as e:  # invalidated captures raise here
        print(f"[capture-probe] capture INVALID after {name}: {str(e)[:120]}", flush=True)
        raise
    if _capture_was[0] and not now:
        print(f"[capture-probe] capture ended/invalid after {name}", flush=True)
    _capture_was[0] = now


def call(name: str, *args) -> None:
    if PROFILE is None:
        check(getattr(load(), name)(*args), name)
        if _DEBUG_CAPTURE:
            _capture_probe(name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(getattr(load(), name)(*args), name)
    b.record()
    PROFILE.setdefault(name, []).append((a, b, tuple(x for x in args if isinstance(x, int) and 0 <= x < 100000)))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    return int(load().tdb_launch_count())


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"turbdiff_b200: {what} must live on a CUDA device (got {t.device}); there is no CPU path")
