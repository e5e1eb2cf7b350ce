"""Two-GPU checks of the multi-GPU paths (skipped on a single-GPU box): sharded sampling equals the unsharded
sampling of each shard with the shard's seed; data-parallel gradients after the bucketed NCCL all-reduce equal the
single-GPU gradients of the concatenated batch."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    sys.path[:0] = [str(root), str(root / "generative-turbulence_b200"), str(root / "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle.cases import CASES, case_inputs
        from oracle.unet_ref import synth_state_dict
        from turbdiff_b200 import DenoisingModel, GaussianDiffusion
        from turbdiff_b200.models.conditioning import Conditioning
        from turbdiff_b200.parallel import GradientAllReduce, sample_sharded, shard_range, shard_seed

        case = CASES["tiny"]
        spec = case["spec"]
        m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=spec.timesteps,
                           dim=spec.dim, u_net_levels=spec.u_net_levels, norm_type="group", precision="fp32")
        m.load_state_dict(synth_state_dict(spec, case["seed"]))
        m = m.cuda()
        gd = GaussianDiffusion(m, timesteps=spec.timesteps, beta_schedule="log-snr-linear", noise_bcs=True).cuda()
        x, _, c_local, geo = case_inputs(case)
        x = torch.cat([x, x.flip(0) * 0.5 + 0.1])  # 4 samples
        x = x.cuda()
        C = {Conditioning.Type.CELL_TYPE: c_local.cuda()}
        idx = torch.from_numpy(geo.cell_idx).cuda()

        # ---- sharded sampling == unsharded sampling of the shard with the shard's seed
        m.eval()
        full = sample_sharded(gd, x, C, idx, seed=5, gather=True, start_from=3)
        rng = shard_range(4, rank, world)
        torch.manual_seed(shard_seed(5, rank))
        mine = gd.p_sample_loop(x[rng.start : rng.stop], C, idx, start_from=3)
        ok = True
        if rank == 0:
            ok = full is not None and full.shape == x.shape and torch.equal(full[rng.start : rng.stop], mine)
        else:
            ok = full is None

        # ---- DP gradients after all-reduce == single-process gradients of the whole batch
        class MD:
            cell_idx = idx

        m.train()
        t_all = torch.tensor([1, 4, 7, 9], device="cuda")
        noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(3)).cuda()
        real = torch.randn_like

        def run(xs, ts, ns):
            m.zero_grad(set_to_none=True)
            torch.randn_like = lambda v, **k: ns
            try:
                loss, _ = gd.p_losses(xs, ts, C, MD, None)
            finally:
                torch.randn_like = real
            loss.backward()
            return loss

        run(x, t_all, noise)
        want = [p.grad.clone() for p in m.parameters()]
        sl = slice(rng.start, rng.stop)
        run(x[sl], t_all[sl], noise[sl].contiguous())
        GradientAllReduce(m.parameters(), bucket_mb=0.5)()
        worst = max(float((p.grad - w).norm() / w.norm().clamp_min(1e-12)) for p, w in zip(m.parameters(), want) if float(w.abs().max()) > 1e-8)
        ok = ok and worst < 2e-4

        # ---- attached fast path, MIXED modes: rank 0 runs the eager launch programs, rank 1 the CUDA-graph replay; both
        # must issue the same collectives (flat buffer, same chunks) and arrive at the global-batch gradients
        red = GradientAllReduce(m.parameters(), flat_chunks=3).attach(m)
        m.engine().train_graph = rank != 0
        for _ in range(2):
            run(x[sl], t_all[sl], noise[sl].contiguous())
            red()
        worst2 = max(float((p.grad - w).norm() / w.norm().clamp_min(1e-12)) for p, w in zip(m.parameters(), want) if float(w.abs().max()) > 1e-8)
        ok = ok and worst2 < 2e-4
        ret[rank] = (bool(ok), max(worst, worst2))
    finally:
        dist.destroy_process_group()


def test_sharded_sampling_and_dp_gradients_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0][0] and ret[1][0], dict(ret)
