"""HDF5 -> device input pipeline (reference OpenFOAMDataRepository.read_data, ofles.py:396-418, followed by
grid_embedding :220-240 and normalize_grid normalization.py:20-24): bit-identical grids, prefetching order, duplicates."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class H5Like:
    """Stand-in for an h5py.Dataset: fancy indexing only with sorted unique indices (h5py raises otherwise)."""

    def __init__(self, arr):
        self.arr, self.reads = arr, 0

    def __getitem__(self, idx):
        idx = np.asarray(idx)
        assert idx.ndim == 1 and np.all(np.diff(idx) > 0), "h5py requires sorted unique indices"
        self.reads += 1
        return self.arr[idx]


def test_pipeline_matches_reference_op_sequence():
    from oracle import grid_ref
    from turbdiff_b200.models import utils as U
    from turbdiff_b200.pipeline import DeviceBatchPipeline

    geo = grid_ref.channel_geometry(cells=(14, 8, 6), hole=((3, 6), (2, 5), (0, 3)), seed=3)
    n_cells, T = len(geo.cell_idx), 23
    rng = np.random.default_rng(0)
    u = rng.standard_normal((T, n_cells, 3)).astype(np.float32)
    p = rng.standard_normal((T, n_cells)).astype(np.float32)  # scalar field: (T, n_cells) on disk
    mean = torch.tensor([0.3, -1.2, 0.05, 101.5])
    std = torch.tensor([1.7, 0.4, 2.5, 13.0])
    idx = torch.from_numpy(geo.cell_idx).cuda()
    inlet = torch.from_numpy(np.flatnonzero(grid_ref.cell_type_map(geo).ravel() == 3))
    fixed = [(inlet, 0, torch.tensor([20.0, 0.0, 0.0]))]
    ds_u, ds_p = H5Like(u), H5Like(p)
    pipe = DeviceBatchPipeline([(ds_u, 3), (ds_p, 1)], idx, geo.padded, mean, std, fixed_values=fixed, max_batch=4, depth=2)
    batches = [[5, 2, 9, 2], [0, 22, 1], [7, 7, 7, 3], [11], [4, 3, 2, 1]]
    n = 0
    for want_idx, b in zip(batches, pipe.run(batches)):
        assert list(b.idxs) == want_idx
        samples = torch.from_numpy(np.concatenate([u[want_idx], p[want_idx][..., None]], -1)).cuda()
        assert torch.equal(b.samples, samples)
        # reference op sequence on the device (ofles.py:220-238, normalization.py:20-24)
        B, F = samples.shape[0], 4
        x = torch.zeros((B, F, geo.n_vox), device="cuda")
        x.transpose(-1, -2)[..., idx, :] = samples
        x.transpose(-1, -2)[..., inlet.cuda(), 0:3] = fixed[0][2].cuda()
        m3, s3 = mean.cuda().view(F, 1, 1, 1), std.cuda().view(F, 1, 1, 1)
        want = torch.addcmul(-m3 / s3, torch.reciprocal(s3), x.view(B, F, *geo.padded))
        assert torch.equal(b.x, want)
        assert torch.equal(b.x, U.scatter_normalize(samples, idx, geo.padded, mean, std, fixed_values=fixed))
        n += 1
    assert n == len(batches)
    assert ds_u.reads == len(batches) and ds_p.reads == len(batches)  # one hyperslab read per variable and batch
    one = pipe.load([3, 1])
    assert torch.equal(one.samples[0, :, :3], torch.from_numpy(u[3]).cuda())


def test_pipeline_rejects_cpu_and_oversized_batches():
    from turbdiff_b200.pipeline import DeviceBatchPipeline

    arr = np.zeros((4, 10, 1), dtype=np.float32)
    with pytest.raises(RuntimeError, match="CUDA"):
        DeviceBatchPipeline([(arr, 1)], np.arange(10), (5, 2, 2), [0.0], [1.0], device="cpu")
    pipe = DeviceBatchPipeline([(H5Like(arr), 1)], np.arange(10), (5, 2, 2), [0.0], [1.0], max_batch=2)
    with pytest.raises(ValueError, match="max_batch"):
        pipe.load([0, 1, 2])
