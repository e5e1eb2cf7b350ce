#!/usr/bin/env python
"""Benchmark of the TurbDiff denoising hot path (BASELINE.json metric: DDPM samples/sec at
192x48x48 on 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per sampling step (configs[2]: 64 samples over 8 GPUs)")
    ap.add_argument("--train-batch", type=int, default=4, help="samples per GPU per training step (configs[1]: batch 4)")
    ap.add_argument("--timesteps", type=int, default=1000, help="T of the sampled chain (BASELINE configs[2]: 1000)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--e2e-steps", type=int, default=16, help="chain length of one end-to-end public-API call")
    ap.add_argument("--train-steps", type=int, default=5, help="timed training steps (configs[1]/[3]: fwd+bwd+all-reduce+optimizer); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def time_cpu_reference(T, n_steps, warm, threads=None):
    """The reference algorithm (CPU oracle port: oracle/unet_ref.py + oracle/diffusion_ref.py) on the host
    cores: B=1 denoise steps of the shapes config.  Returns (seconds per step, cores)."""
    from oracle.diffusion_ref import DiffusionRef
    from oracle.unet_ref import denoiser_forward, synth_state_dict

    if threads:
        torch.set_num_threads(threads)
    spec = shapes_spec(T)
    sd = synth_state_dict(spec, 0)
    geo, x, c_local = synthetic_inputs(1, 1)
    idx = torch.from_numpy(geo.cell_idx)
    d = DiffusionRef(lambda xt, tt: denoiser_forward(sd, spec, xt, tt, c_local), timesteps=T, beta_schedule="log-snr-linear", noise_bcs=True)
    times = []
    with torch.no_grad():
        xt = torch.randn_like(x)
        for i in range(warm + n_steps):
            t0 = time.perf_counter()
            tt = torch.full((1,), T - 1 - i, dtype=torch.long)
            _, _, mean, log_var = d.predictions(xt, tt, idx)
            z = torch.randn_like(xt)
            xt = mean + (log_var / 2).exp() * z
            from oracle.diffusion_ref import where_cells

            xt = where_cells(idx, xt, d.q_sample(x, tt, torch.randn_like(x)))
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = args.timesteps
    sec, cores = time_cpu_reference(T, max(1, args.steps), max(0, args.warmup))
    value = 1.0 / (sec * T)
    sample = f"B=1 single denoise steps of the shapes config ({args.steps} timed, {args.warmup} warm-up), extrapolated to T={T} steps per sample"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[2] DDPM ancestral sampling, shapes config 194x50x50 u+p, T={T}", "batch_per_step": 1, "timesteps": T},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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

    x_t = torch.randn_like(x_bcs)
    t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    t_vec = torch.zeros(B, dtype=torch.int64, device=dev)
    step_no = [0]

    graph = None

    def unet():
        if graph is not None:
            graph.replay()
            return eng.plan(B, geo.padded, dev)["eps"]
        return eng.forward(x_t, t_vec, cl, c_static=True)  # inside a sampling chain C is constant (as in p_sample_loop)

    def one_step():
        t = T - 1 - (step_no[0] % (T - 1))
        step_no[0] += 1
        t_dev.fill_(t)
        t_vec.fill_(t)
        eps = unet()
        z = torch.randn_like(x_t)
        z_bc = torch.randn_like(x_bcs)
        _lib.call("tdb_ddpm_step", x_t.data_ptr(), eps.data_ptr(), z.data_ptr(), z_bc.data_ptr(), x_bcs.data_ptr(), mask.data_ptr(),
                  coef.data_ptr(), t_dev.data_ptr(), x_t.data_ptr(), B, 4, nvox, flags, _lib.stream_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also builds the plan and the packed-weight cache), then capture the U-Net launch program
    one_step()
    torch.cuda.synchronize()
    launches_per_unet = None
    if not args.no_graph:
        n0 = _lib.launch_count()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eng.forward(x_t, t_vec, cl, c_static=True)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eng.forward(x_t, t_vec, cl, c_static=True)
        launches_per_unet = (_lib.launch_count() - n0) // 2
        graph = g
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
    if graph is not None:
        launches += launches_per_unet * args.steps  # graph replays re-issue the captured C-ABI launches
    clk = clocks.stop()
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = world * B / (T * ms_per_step * 1e-3)

    # ---- end to end through the public API: host buffers in, host buffers out -----------------------
    S = max(2, args.e2e_steps)
    graph_saved, graph = graph, None

    def e2e_call():
        xb = x_pinned.to(dev, non_blocking=True)
        s = gd.p_sample_loop(xb, C, cell_idx, start_from=S)
        out_pinned.copy_(s, non_blocking=True)

    e2e_call()
    barrier()
    reps = 2
    e0.record()
    for _ in range(reps):
        e2e_call()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(e2e_ms.item()) * 1e-3 * T / S)
    graph = graph_saved

    # ---- per-kernel device times (CUDA events around every C-ABI launch, eager, after the timed region)
    roof = None
    if rank == 0:
        _lib.PROFILE = {}
        graph_saved, graph = graph, None
        for _ in range(3):
            one_step()
        torch.cuda.synchronize()
        prof = {k: sum(a.elapsed_time(b) for a, b, _ in v) / 3 for k, v in _lib.PROFILE.items()}
        _lib.PROFILE = None
        graph = graph_saved
        pk = peaks()
        conv_ms = sum(v for k, v in prof.items() if k.startswith("tdb_conv3d"))  # all convolution kernels of one step
        flops = conv_flops_per_sample(geo.padded) * B
        ach = flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        # ncu (profiles/r01_launches_v52_b8.csv, r01_ncu_full_win_v40.json): the top kernel (row-window conv, 64->64
        # @194x50x50, B=8) moves 543 MB + 456 MB of DRAM traffic per launch = its algorithmic bytes (input read once, output written once)
        traffic = 1.0015e9 if (B == 8 and args.precision == "bf16") else None
        roof = {"bound": "tensor", "kernel": "tcgen05 convolution family: conv3d_bf16_winz / _win / _winp / _fold2 / _tc kernels (all conv launches of one step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_of": "conv3d_bf16_winz_kernel<64,64> per launch (ncu dram__bytes_read+write, round-1 capture)" if traffic else None,
                "peak_src": f"{pk['src']} bf16 sustained (kernel timed inside a long step)",
                "conv_ms_per_step": conv_ms, "kernel_ms_per_step": prof}
        # the bandwidth-bound update kernel against the HBM roofline: 6 tensors x 4 B per element
        st_ms = prof.get("tdb_ddpm_step", 0.0)
        if st_ms > 0:
            bytes_step = 6 * 4 * B * 4 * nvox + nvox
            roof["ddpm_step"] = {"bound": "hbm", "achieved": bytes_step / (st_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": bytes_step / (st_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": bytes_step}

    # ---- training step (BASELINE configs[1] / [3]): q_sample + U-Net fwd + bwd (+ NCCL grad all-reduce) + clip + RAdam
    train = None
    if args.train_steps > 0:
        from turbdiff_b200.parallel import GradientAllReduce

        graph = None
        torch.cuda.empty_cache()
        model.train()
        from turbdiff_b200.optim import FusedRAdam

        # the reference's optimiser (torch.optim.RAdam, diffusion.py:216) with Lightning's gradient_clip_val 0.1 (norm):
        # both in the fused two-launch step
        opt = FusedRAdam(model.parameters(), lr=1e-4, max_grad_norm=0.1)
        reducer = GradientAllReduce(model.parameters())

        class MD:
            pass

        MD.cell_idx = cell_idx

        TB = min(args.train_batch, B)
        x_train = x_bcs[:TB].contiguous()

        def train_step():
            opt.zero_grad(set_to_none=True)
            loss, _ = gd(x_train, C, MD, None)
            loss.backward()
            reducer()
            opt.step()  # clip_grad_norm_(0.1) + RAdam
            return loss

        for _ in range(3):  # the first call captures the two CUDA graphs, the second uploads them
            train_step()
        barrier()
        n_t0 = _lib.launch_count() + model.engine().replayed_launches
        e0.record()
        th0 = time.perf_counter()
        for _ in range(args.train_steps):
            loss = train_step()
        host_ms = (time.perf_counter() - th0) * 1e3 / args.train_steps  # CPU time to enqueue one step (no sync inside)
        n_train_launches = (_lib.launch_count() + model.engine().replayed_launches - n_t0) // args.train_steps
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1) / args.train_steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # one more step with per-kernel events (every rank takes part: the step contains a collective)
        train_prof = None
        if rank == 0:
            _lib.PROFILE = {}
        train_step()
        torch.cuda.synchronize()
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
                 "includes": "q_sample + U-Net forward + backward + bucketed NCCL gradient all-reduce (N>1) + fused clip_grad_norm(0.1) + RAdam step (turbdiff_b200.optim.FusedRAdam)",
                 "train_flops_per_step": 3 * conv_flops_per_sample(geo.padded) * TB}
        train["tflops"] = train["train_flops_per_step"] / (train["ms_per_step"] * 1e-3) / 1e12
        model.eval()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores = time_cpu_reference(T, 3, 1)
        cpu = {"value": 1.0 / (sec * T), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 timed B=1 denoise steps of the same workload on the host ({sec:.2f} s/step), extrapolated to T={T}"}

    if rank == 0:
        h2d = x_pinned.numel() * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"configs[2] DDPM ancestral sampling, shapes config 194x50x50 u+p (dim 32, 4 levels, 55.2M params), T={T}",
                       "batch_per_gpu": B, "timesteps": T, "step": "one denoise step of the whole batch (U-Net forward + 2 randn + fused update)",
                       "l2": "activations exceed L2 (>= 270 MB per level-0 tensor), no flush needed", "cuda_graph": graph is not None,
                       "samples_per_sec_T500": value * T / 500},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d,
                    "how": f"GaussianDiffusion.p_sample_loop(start_from={S}) on pinned host x_bcs -> pinned host sample, scaled by T/{S}"},
            "roofline": roof, "cpu_baseline": cpu, "train": train,
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
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (turbdiff_b200 has no CPU path); use --impl reference for the host baseline")
        run_ours(args)


if __name__ == "__main__":
    main()
