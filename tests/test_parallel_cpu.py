"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sample sharding and the bucketed
data-parallel gradient all-reduce.  The kernels themselves are covered by the -m gpu tests."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from turbdiff_b200.parallel import GradientAllReduce, gather_shards, shard_range, shard_seed


def test_shard_range_partitions_in_order():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert shard_seed(10, 3) == 13


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
        data = torch.randn(8, 7)
        target = torch.randn(8, 2)
        # global-batch gradient (what a single process would compute)
        net.zero_grad()
        torch.nn.functional.mse_loss(net(data), target).backward()
        want = [p.grad.clone() for p in net.parameters()]
        # data parallel: each rank takes its slice, then the bucketed average all-reduce
        net.zero_grad()
        rng = shard_range(8, rank, world)
        sl = slice(rng.start, rng.stop)
        torch.nn.functional.mse_loss(net(data[sl]), target[sl]).backward()
        GradientAllReduce(net.parameters(), bucket_mb=1e-4)()  # tiny buckets: several all-reduces in flight
        ok = all(torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(net.parameters(), want))
        # a parameter without gradient on one rank still takes part
        extra = torch.nn.Parameter(torch.zeros(3))
        if rank == 0:
            extra.grad = torch.ones(3)
        GradientAllReduce([extra])()
        ok = ok and torch.allclose(extra.grad, torch.full((3,), 1.0 / world))
        # fast path: the exchange on an attached denoiser's flat gradient buffer (engine.grad_sync hook), after which the
        # generic call only reduces the parameters that were not covered
        class _Eng:
            grad_sync = None

        class _Den(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a = torch.nn.Parameter(torch.zeros(5))
                self.b = torch.nn.Parameter(torch.zeros(2, 3))
                self._eng = _Eng()

            def engine(self):
                return self._eng

        den = _Den()
        other = torch.nn.Parameter(torch.zeros(4))
        red = GradientAllReduce(list(den.parameters()) + [other], flat_chunks=3).attach(den)
        flat = torch.arange(11.0) * (rank + 1)              # what the backward program would hand out on this rank
        den.engine().grad_sync(flat)
        ok = ok and torch.allclose(flat, torch.arange(11.0) * (sum(range(1, world + 1)) / world))
        den.a.grad, den.b.grad = flat[:5].clone(), flat[5:].view(2, 3).clone()
        other.grad = torch.full((4,), float(rank))
        red()                                               # must not reduce den's gradients a second time
        ok = ok and torch.allclose(den.a.grad, flat[:5]) and torch.allclose(other.grad, torch.full((4,), (world - 1) / 2))
        # split exchange: the phase-1 part is started while the backward pass still runs, the rest afterwards, one finish
        flat2 = torch.arange(11.0) * (rank + 1)
        hook = den.engine().grad_sync
        hook.start(flat2[:4])
        hook.start(flat2[4:])
        hook.finish()
        ok = ok and torch.allclose(flat2, torch.arange(11.0) * (sum(range(1, world + 1)) / world)) and not hook.pending
        # gather of sharded "samples" restores the original order on rank 0
        full = torch.arange(10.0).reshape(5, 2)
        mine = full[shard_range(5, rank, world).start : shard_range(5, rank, world).stop]
        got = gather_shards(mine.contiguous(), 5)
        if rank == 0:
            ok = ok and torch.equal(got, full)
        else:
            ok = ok and got is None
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_gather_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
