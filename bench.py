#!/usr/bin/env python
"""Benchmark of the TurbDiff denoising hot path (BASELINE.json metric: DDPM samples/sec at
192x48x48 on 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W                # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W        # the UNMODIFIED reference on the host cores (CPU)
    python bench.py --impl reference-gpu --steps K --warmup W    # the unmodified reference on one B200 (torch eager)

One "step" = one ancestral-sampling step of the shapes-config model (dim 32, 4 levels, padded grid
194x50x50, u+p) over a batch of B samples per GPU: U-Net forward, two Gaussian draws and the fused
posterior update.  A full sample is T such steps, so samples/s = B*N / (T * t_step).  Samples are
independent: ranks share nothing (weak scaling, no data-path collective).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "generative-turbulence_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "ddpm_samples_per_sec"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per sampling step (configs[2]: 64 samples over 8 GPUs)")
    ap.add_argument("--train-batch", type=int, default=4, help="samples per GPU per training step (configs[1]: batch 4)")
    ap.add_argument("--timesteps", type=int, default=1000, help="T of the sampled chain (BASELINE configs[2]: 1000)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--e2e-steps", type=int, default=0,
                    help="chain length of one end-to-end public-API call; 0 (default) = ONE COMPLETE chain of T steps, nothing extrapolated")
    ap.add_argument("--train-steps", type=int, default=20, help="timed training steps (configs[1]/[3]: fwd+bwd+all-reduce+optimizer); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the gpu_reference block (the reference under torch eager on this GPU)")
    ap.add_argument("--ref-mode", default="tf32", choices=["tf32", "bf16", "fp32"], help="--impl reference-gpu: TF32 (the reference's "
                    "shapes experiment, matmul_precision=medium), bf16 autocast, or strict fp32")
    ap.add_argument("--no-graph", action="store_true")
    return ap.parse_args()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def shapes_spec(T):
    from oracle.unet_ref import UNetSpec

    return UNetSpec(in_features=4, out_features=4, c_local_features=4, timesteps=T, dim=32, u_net_levels=4, groups=8)


def synthetic_inputs(B, seed):
    """x_bcs ~ N(0,1) on the padded grid and the cell-type conditioning (turbdiff_b200.synthetic: no oracle involved)."""
    from turbdiff_b200.synthetic import synthetic_inputs as make

    return make(B, seed)


def conv_flops_per_sample(spatial):
    """Algorithmic FLOPs of one denoiser forward (666.2 GFLOP at the shapes config)."""
    from turbdiff_b200.synthetic import conv_flops_per_sample as flops

    return flops(spatial)


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples of one GPU during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # power_w: the board power during the timed region - SM clocks below max at ~1 kW are the power limiter at work
        # even when the 20 ms sampling misses the sw_power_cap flag
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "power_w": statistics.median(pw) if pw else None}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload(T):
    return f"configs[2] DDPM ancestral sampling, shapes config 194x50x50 u+p (dim 32, 4 levels, 55.2M params), T={T}"


def build_reference(T, device, seed=0):
    """The unmodified reference's DenoisingModel + GaussianDiffusion (oracle/ref_shim.py) at the shapes configuration
    (config/model/diffusion.yaml of the reference), random init under torch.manual_seed(seed)."""
    from oracle import ref_shim

    ns = ref_shim.load()
    torch.manual_seed(seed)
    m = ns.ddpm.DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=T, dim=32,
                               u_net_levels=4, norm_type="group")
    gd = ns.ddpm.GaussianDiffusion(m, timesteps=T, beta_schedule="log-snr-linear", loss_type="l2", noise_bcs=True)
    return ns, gd.to(device).eval()


def time_cpu_reference(T, n_steps, warm, threads=None):
    """The reference's own CPU path of the hot loop on the host cores: `GaussianDiffusion.p_sample_loop` (ddpm.py:767-816)
    of the UNMODIFIED reference (kind "reference": /root/reference or the oracle/_ref install) over `n_steps` ancestral
    steps of ONE sample of the shapes config; the CPU oracle port (kind "port") only if the reference is not installed.
    torchrun exports OMP_NUM_THREADS=1, so the thread count is set explicitly to every host core this process may use.
    Returns (seconds per step, threads, kind)."""
    from oracle import ref_shim

    threads = threads or host_cores()
    torch.set_num_threads(threads)
    geo, x, c_local = synthetic_inputs(1, 1)
    idx = torch.from_numpy(geo.cell_idx)
    if ref_shim.available():
        ns, gd = build_reference(T, "cpu")
        C = {ns.Conditioning.Type.CELL_TYPE: c_local}
        with torch.no_grad():
            if warm > 0:
                gd.p_sample_loop(x, C, idx, start_from=warm)
            t0 = time.perf_counter()
            gd.p_sample_loop(x, C, idx, start_from=n_steps)
            sec = (time.perf_counter() - t0) / n_steps
        return sec, torch.get_num_threads(), "reference"
    from oracle.diffusion_ref import DiffusionRef
    from oracle.unet_ref import denoiser_forward, synth_state_dict

    spec = shapes_spec(T)
    sd = synth_state_dict(spec, 0)
    d = DiffusionRef(lambda xt, tt: denoiser_forward(sd, spec, xt, tt, c_local), timesteps=T, beta_schedule="log-snr-linear", noise_bcs=True)
    with torch.no_grad():
        if warm > 0:
            d.sample_loop(x, idx, start_from=warm)
        t0 = time.perf_counter()
        d.sample_loop(x, idx, start_from=n_steps)
        sec = (time.perf_counter() - t0) / n_steps
    return sec, torch.get_num_threads(), "port"


def time_gpu_reference(T, B, n_steps, warm, mode, dev):
    """the same reference code under this image's torch (eager, cuDNN) on this B200, timed with the
    protocol of scripts/evaluate-runtime.py:64-82 (synchronize around the sampling call) on `p_sample_loop` chains of
    `n_steps` steps at batch B.  mode "tf32" = matmul_precision medium (train.py:144-150, config/shapes_experiment.yaml:59),
    "bf16" = the same under torch.autocast(bfloat16), "fp32" = matmul_precision highest."""
    ns, gd = build_reference(T, dev)
    geo, x, c_local = synthetic_inputs(B, 100)
    x = x.to(dev)
    C = {ns.Conditioning.Type.CELL_TYPE: c_local.to(dev)}
    idx = torch.from_numpy(geo.cell_idx).to(dev)
    old = (torch.get_float32_matmul_precision(), torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        torch.set_float32_matmul_precision("highest" if mode == "fp32" else "medium")
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16" else torch.autocast("cuda", enabled=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad(), ctx:
            gd.p_sample_loop(x, C, idx, start_from=max(1, warm))
            torch.cuda.synchronize()
            e0.record()
            gd.p_sample_loop(x, C, idx, start_from=n_steps)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_steps
    finally:
        torch.set_float32_matmul_precision(old[0])
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old[1], old[2]
        del gd
        torch.cuda.empty_cache()
    return ms


def gpu_reference_block(T, B, dev, ours_ms_per_step):
    """`gpu_reference` of the bench line: the unmodified reference on this GPU in the two precisions it is run at
    (TF32 = its shapes experiment; bf16 autocast), beside this repo's step time.  A baseline leg like cpu_baseline:
    reported, never part of the measured path."""
    from oracle import ref_shim

    if not ref_shim.available():
        return {"unavailable": "reference package not installed (oracle/install_ref.py)"}
    torch.cuda.empty_cache()
    out = {"batch_per_gpu": B, "how": "unmodified reference GaussianDiffusion.p_sample_loop, torch eager + cuDNN, 6 timed + 2 warm-up steps"}
    for mode in ("tf32", "bf16"):
        try:
            ms_ref = time_gpu_reference(T, B, 6, 2, mode, dev)
            out[mode] = {"ms_per_step": ms_ref, "value": B / (T * ms_ref * 1e-3), "unit": UNIT, "speedup_of_this_repo": ms_ref / ours_ms_per_step}
        except Exception as e:  # the bar is reported, never allowed to break the bench line
            out[mode] = {"error": str(e)[:200]}
    return out


def run_reference(args):
    """Reference arm.  Under torchrun only rank 0 works (and uses every host core); the other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = args.timesteps
    steps, warm = max(1, args.steps), max(0, args.warmup)
    if args.impl == "reference-gpu":
        if not torch.cuda.is_available():
            _emit({"impl": "reference-gpu", "unavailable": "no CUDA device"})
            return
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        ms = time_gpu_reference(T, args.batch, steps, warm, args.ref_mode, dev)
        value = args.batch / (T * ms * 1e-3)
        _emit({"impl": "reference-gpu", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warm,
               "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.ref_mode, "data": "synthetic",
               "config": {"workload": workload(T), "batch_per_gpu": args.batch, "timesteps": T},
               "how": "unmodified reference GaussianDiffusion.p_sample_loop under torch eager on this GPU (scripts/evaluate-runtime.py protocol)"})
        return
    sec, cores, kind = time_cpu_reference(T, steps, warm)
    value = 1.0 / (sec * T)
    sample = (f"{steps} timed + {warm} warm-up ancestral steps of ONE of the batch's {args.batch} samples through "
              f"{'the unmodified reference GaussianDiffusion.p_sample_loop' if kind == 'reference' else 'the CPU oracle port'} "
              f"on {cores} host threads ({sec:.2f} s/step); samples/s = 1 / (s_per_step * T) - per-sample cost, batch independent")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(T), "batch_per_gpu": args.batch, "timesteps": T},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def run_ours(args):
    import torch.distributed as dist

    from turbdiff_b200 import DenoisingModel, GaussianDiffusion, _lib
    from turbdiff_b200.models.conditioning import Conditioning
    from turbdiff_b200.models.utils import inside_mask

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, B = args.timesteps, args.batch
    torch.manual_seed(0)  # random-init weights of the architecture (the reference's own initialisers, same order)
    model = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=T, dim=32, u_net_levels=4,
                           norm_type="group", precision=args.precision)
    model = model.to(dev).eval()
    gd = GaussianDiffusion(model, timesteps=T, beta_schedule="log-snr-linear", loss_type="l2", noise_bcs=True).to(dev)
    geo, x_host, c_local = synthetic_inputs(B, 100 + rank)
    x_pinned = x_host.pin_memory()
    out_pinned = torch.empty_like(x_host).pin_memory()
    x_bcs = x_pinned.to(dev, non_blocking=True)
    C = {Conditioning.Type.CELL_TYPE: c_local.to(dev)}
    cl = C[Conditioning.Type.CELL_TYPE]
    cell_idx = torch.from_numpy(geo.cell_idx).to(dev)
    nvox = int(np.prod(geo.padded))
    mask = inside_mask(cell_idx, nvox)
    coef = gd._coef_table(dev)
    eng = model.engine()
    flags = _lib.STEP_NOISE_BCS

    # The step is issued exactly as GaussianDiffusion.p_sample_loop issues it: the engine's persistent sampler state, the
    # denoiser launch program replayed from a CUDA graph, the two Gaussian draws on a side stream, and the fused step
    # tail (decoder block tail + decode.1 + posterior update + next step's encode_x in one kernel).
    st = eng.sampler_state(B, tuple(geo.padded), dev, cl)
    x_t, t_dev, t_vec = st["x_t"], st["t_dev"], st["t_vec"]
    x_t.copy_(torch.randn_like(x_bcs))
    eng.use_graph = not args.no_graph
    fused = eng.can_fuse_tail()
    state = {"cur": x_t, "nxt": st["x_t2"]}
    if fused:
        eng.encode_state(st, x_t)
    step_no = [0]
    rng_stream = torch.cuda.Stream(device=dev)
    side_rng = os.environ.get("TURBDIFF_B200_RNG_STREAM", "1") != "0"

    pending = []

    def one_step():
        # (the public sampling loop throttles the host the same way: at most 3 steps ahead of the device)
        if len(pending) >= 3:
            pending.pop(0).synchronize()
        t = T - 1 - (step_no[0] % (T - 1))
        step_no[0] += 1
        t_dev.fill_(t)
        t_vec.fill_(t)
        main = torch.cuda.current_stream()
        if side_rng:
            rng_stream.wait_stream(main)
            with torch.cuda.stream(rng_stream):
                z = torch.randn_like(x_t)
                z_bc = torch.randn_like(x_bcs)
            z.record_stream(main)
            z_bc.record_stream(main)
            eps = eng.forward_graphed(st, tail=fused)
            main.wait_stream(rng_stream)
        else:
            eps = eng.forward_graphed(st, tail=fused)
            z = torch.randn_like(x_t)
            z_bc = torch.randn_like(x_bcs)
        cur, nxt = state["cur"], state["nxt"]
        if fused:
            eng.step_tail(st, cur, nxt, z, z_bc, x_bcs, mask, coef, t_dev, flags)
            state["cur"], state["nxt"] = nxt, cur
        else:
            _lib.call("tdb_ddpm_step", cur.data_ptr(), eps.data_ptr(), z.data_ptr(), z_bc.data_ptr(), x_bcs.data_ptr(), mask.data_ptr(),
                      coef.data_ptr(), t_dev.data_ptr(), cur.data_ptr(), B, 4, nvox, flags, _lib.stream_ptr())
        ev = torch.cuda.Event()
        ev.record()
        pending.append(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: builds the plan and the packed-weight cache and captures the denoiser graph (engine.forward_graphed)
    one_step()
    torch.cuda.synchronize()
    used_graph = eng.use_graph
    launches_per_unet = None
    if used_graph:
        # C-ABI launches the graph re-issues per replay: counted once from an eager pass of the same program
        n0 = _lib.launch_count()
        eng.use_graph = False
        eng.forward_graphed(st, tail=fused)
        eng.use_graph = True
        launches_per_unet = _lib.launch_count() - n0
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        one_step()

    # ---- timed region: device-resident inputs ---------------------------------------------------
    clocks = ClockSampler(local)
    barrier()
    n_l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n_l0
    if used_graph:
        launches += launches_per_unet * args.steps  # graph replays re-issue the captured C-ABI launches
    clk = clocks.stop()
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = world * B / (T * ms_per_step * 1e-3)

    # ---- end to end through the public API: host buffers in, host buffers out -----------------------
    # default: one COMPLETE chain (start_from=None: x_T ~ N(0, I), T ancestral steps) per call, so nothing is
    # extrapolated; shorter chains (--e2e-steps S) are scaled by T/S and over-count the per-call copies T/S times
    S = T if args.e2e_steps <= 0 else max(2, min(T, args.e2e_steps))
    def e2e_call(n):
        xb = x_pinned.to(dev, non_blocking=True)
        s = gd.p_sample_loop(xb, C, cell_idx, start_from=None if n == T else n)
        out_pinned.copy_(s, non_blocking=True)

    e2e_call(min(S, 8))  # warm-up: sampler state, graph and the allocator's noise blocks exist afterwards
    barrier()
    reps = 1 if S >= 256 else 2
    w0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        e2e_call(S)
    e1.record()
    barrier()
    e2e_wall_s = (time.perf_counter() - w0) / reps
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(e2e_ms.item()) * 1e-3 * T / S)
    if fused:
        eng.encode_state(st, state["cur"])  # the public-API chains reused the plan's input buffer

    # ---- per-kernel device times (CUDA events around every C-ABI launch, eager, after the timed region)
    roof = None
    if rank == 0:
        _lib.PROFILE = {}
        eng.use_graph = False  # eager launch program: one pair of CUDA events around every C-ABI call
        for _ in range(3):
            one_step()
        torch.cuda.synchronize()
        prof = {k: sum(a.elapsed_time(b) for a, b, _ in v) / 3 for k, v in _lib.PROFILE.items()}
        _lib.PROFILE = None
        eng.use_graph = used_graph
        pk = peaks()
        conv_ms = sum(v for k, v in prof.items() if k.startswith("tdb_conv3d"))  # all convolution kernels of one step
        flops = conv_flops_per_sample(geo.padded) * B
        ach = flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        # DRAM traffic of the top kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
        # capture of THIS workload, committed under profiles/ (bench.py cannot run under ncu itself); null when the
        # capture does not match the run's batch / precision
        traffic, traffic_of = None, None
        tf = ROOT / "profiles" / "ncu_traffic.json"
        if tf.exists():
            td = json.loads(tf.read_text())
            if td.get("batch") == B and td.get("precision") == args.precision:
                traffic, traffic_of = td["dram_bytes_per_launch"], f"{td['kernel']} ({td['source']})"
        from turbdiff_b200.synthetic import conv_flops_per_sample as _flops

        halo_factor = _flops(geo.padded, haloed=True) / _flops(geo.padded)
        roof = {"bound": "tensor", "kernel": "tcgen05 convolution family: conv3d_bf16_winz / _win / _winp / _fold2 / _tc kernels (all conv launches of one step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_of": traffic_of,
                # informative: the same time against the rows the implicit GEMMs process (halo rows of the halo-grid layout are
                # computed and discarded); `achieved` / `frac` above stay on the ALGORITHMIC FLOPs
                "processed_rows": {"flops_factor": halo_factor, "achieved": ach * halo_factor, "frac": ach * halo_factor / peak},
                "peak_src": f"{pk['src']} bf16 sustained (kernel timed inside a long step)",
                "conv_ms_per_step": conv_ms, "kernel_ms_per_step": prof}
        # the bandwidth-bound update kernel against the HBM roofline: 6 tensors x 4 B per element
        tl_ms = prof.get("tdb_step_tail", 0.0)
        if tl_ms > 0:
            # fused step tail: reads raw2 + res (bf16 halo grids, dim channels), x_t / z / z' / x_bcs, writes x' and the next step's
            # encoded input rows (bf16 halo grid, dim channels)
            rows_p = B * int(np.prod([d + 2 for d in geo.padded]))
            esz = 2 if args.precision == "bf16" else 4
            bytes_tail = 3 * rows_p * 32 * esz + 5 * 4 * B * 4 * nvox + nvox
            roof["step_tail"] = {"bound": "hbm", "achieved": bytes_tail / (tl_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": bytes_tail / (tl_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": bytes_tail}
        st_ms = prof.get("tdb_ddpm_step", 0.0)
        if st_ms > 0:
            bytes_step = 6 * 4 * B * 4 * nvox + nvox
            roof["ddpm_step"] = {"bound": "hbm", "achieved": bytes_step / (st_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": bytes_step / (st_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": bytes_step}

    # ---- training step (BASELINE configs[1] / [3]): q_sample + U-Net fwd + bwd (+ NCCL grad all-reduce) + clip + RAdam
    train = None
    if args.train_steps > 0:
        from turbdiff_b200.parallel import GradientAllReduce

        torch.cuda.empty_cache()
        model.train()
        from turbdiff_b200.optim import FusedRAdam

        # the reference's optimiser (torch.optim.RAdam, diffusion.py:216) with Lightning's gradient_clip_val 0.1 (norm):
        # both in the fused two-launch step
        opt = FusedRAdam(model.parameters(), lr=1e-4, max_grad_norm=0.1)
        reducer = GradientAllReduce(model.parameters()).attach(model)  # all-reduce on the backward program's flat gradient buffer

        class MD:
            pass

        MD.cell_idx = cell_idx

        TB = min(args.train_batch, B)
        x_train = x_bcs[:TB].contiguous()

        train_pending, train_wait = [], [0.0]

        def train_step():
            # the host stays at most 2 steps ahead of the device (a real loop reads the loss every step; unthrottled, all
            # timed steps are enqueued within a few ms and their transient tensors - 0.35 GB per step - pile up in the allocator)
            if len(train_pending) >= 2:
                tw = time.perf_counter()
                train_pending.pop(0).synchronize()
                train_wait[0] += time.perf_counter() - tw
            opt.zero_grad(set_to_none=True)
            loss, _ = gd(x_train, C, MD, None)
            loss.backward()
            reducer()
            opt.step()  # clip_grad_norm_(0.1) + RAdam
            ev = torch.cuda.Event()
            ev.record()
            train_pending.append(ev)
            return loss

        for _ in range(3):  # the first call captures the two CUDA graphs, the second uploads them
            train_step()
        barrier()
        n_t0 = _lib.launch_count() + model.engine().replayed_launches
        e0.record()
        th0, wait0 = time.perf_counter(), train_wait[0]
        for _ in range(args.train_steps):
            loss = train_step()
        host_ms = (time.perf_counter() - th0 - (train_wait[0] - wait0)) * 1e3 / args.train_steps  # CPU time to enqueue one step (throttle waits excluded)
        n_train_launches = (_lib.launch_count() + model.engine().replayed_launches - n_t0) // args.train_steps
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1) / args.train_steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # one more step with per-kernel events (every rank takes part: the step contains a collective)
        # (ALL ranks switch to the per-call profiling mode: it runs the eager launch programs, whose gradient exchange is the
        # bucketed generic path - a rank left in graph mode would issue the flat-buffer collectives instead and the two
        # sequences would never match)
        train_prof = None
        _lib.PROFILE = {}
        train_step()
        torch.cuda.synchronize()
        if rank != 0:
            _lib.PROFILE = None
        if rank == 0:
            train_prof = {k: sum(a.elapsed_time(b) for a, b, _ in v) for k, v in _lib.PROFILE.items()}
            if os.environ.get("TDB_PROFILE_CALLS"):  # per-call dump (name, ms, small integer arguments) for kernel work
                with open(os.environ["TDB_PROFILE_CALLS"], "w") as fh:
                    for k, v in _lib.PROFILE.items():
                        for a, b, ints in v:
                            fh.write(f"{k}\t{a.elapsed_time(b):.4f}\t{ints}\n")
            _lib.PROFILE = None
        train = {"steps_per_sec": 1e3 / float(tt.item()), "ms_per_step": float(tt.item()), "batch_per_gpu": TB, "global_batch": TB * world,
                 "kernel_ms_per_step": train_prof, "host_enqueue_ms_per_step": host_ms,
                 "loss": float(loss.item()), "kernel_launches_per_step": n_train_launches,
                 "includes": "q_sample + U-Net forward + backward + NCCL gradient all-reduce on the flat gradient buffer (N>1) + fused clip_grad_norm(0.1) + RAdam step (turbdiff_b200.optim.FusedRAdam)",
                 "train_flops_per_step": 3 * conv_flops_per_sample(geo.padded) * TB}
        train["tflops"] = train["train_flops_per_step"] / (train["ms_per_step"] * 1e-3) / 1e12
        model.eval()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores, kind = time_cpu_reference(T, 4, 1)
        cpu = {"value": 1.0 / (sec * T), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"4 timed + 1 warm-up ancestral steps of ONE sample of the same workload through "
                         f"{'the unmodified reference p_sample_loop' if kind == 'reference' else 'the CPU oracle port'} on {cores} host "
                         f"threads ({sec:.2f} s/step), samples/s = 1 / (s_per_step * T)"}

    # the reference's own code on THIS GPU (torch eager + cuDNN): the throughput bar of SURVEY 8(d)(ii)
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        gpu_ref = gpu_reference_block(T, B, dev, ms_per_step)

    if rank == 0:
        h2d = x_pinned.numel() * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload(T), "batch_per_gpu": B, "timesteps": T, "step": "one denoise step of the whole batch (U-Net forward + 2 randn + fused update)",
                       "l2": "activations exceed L2 (>= 270 MB per level-0 tensor), no flush needed", "cuda_graph": used_graph,
                       "samples_per_sec_T500": value * T / 500},
            "clocks": clk, "gpu_launches": int(launches),
            # the copies happen once per chain (the public call takes x_bcs and returns the samples): per denoise step that is
            # bytes / S; *_per_chain are the bytes of one call
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / S, "d2h_bytes_per_step": h2d / S,
                    "h2d_bytes_per_chain": h2d, "d2h_bytes_per_chain": h2d, "chain_steps": S, "chains_timed": reps,
                    "device_s_per_chain": float(e2e_ms.item()) * 1e-3, "wall_s_per_chain": e2e_wall_s,
                    "how": (f"ONE complete chain: GaussianDiffusion.p_sample_loop(x_bcs, C, cell_idx) with T={T} on pinned host x_bcs -> "
                            "pinned host sample, nothing extrapolated" if S == T else
                            f"GaussianDiffusion.p_sample_loop(start_from={S}) on pinned host x_bcs -> pinned host sample, scaled by T/{S}")},
            "roofline": roof, "cpu_baseline": cpu, "gpu_reference": gpu_ref,
            "train_steps_per_sec": None if train is None else train["steps_per_sec"],
            "train_ms_per_step": None if train is None else train["ms_per_step"],
            "train_samples_per_sec": None if train is None else train["steps_per_sec"] * train["global_batch"],
            "train": train,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # C-level writers to fd 1 (e.g. the "NCCL version" banner) must not pollute the JSON line
    args = parse()
    if args.impl in ("reference", "reference-gpu"):
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (turbdiff_b200 has no CPU path); use --impl reference for the host baseline")
        run_ours(args)


if __name__ == "__main__":
    main()
