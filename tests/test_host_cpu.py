"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol
declared in include/turbdiff_b200.h, the module tree honours the reference's checkpoint contract,
the schedule buffers equal the reference's bit for bit, and the launch program is well formed.
No kernel is executed here (there is no GPU in the build container and no CPU fallback)."""

import ctypes
import json
import re
from collections import Counter
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge

    ge.build()
    from turbdiff_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol(built_lib):
    header = (ROOT / "include" / "turbdiff_b200.h").read_text()
    declared = set(re.findall(r"\b(tdb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 19
    lib = ctypes.CDLL(str(built_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    bound = set(built_lib.SIGNATURES) | set(built_lib.OTHER)
    assert declared == bound, declared ^ bound
    assert built_lib.load().tdb_version() >= 100
    assert built_lib.load().tdb_last_error() is not None


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    lib = built_lib.load()
    rc = lib.tdb_conv3d_f32(None, 8, None, None, None, 8, 1, 4, 4, 4, 8, 8, 27, None)
    assert rc == -1
    assert b"null pointer" in lib.tdb_last_error()
    with pytest.raises(RuntimeError, match="null pointer"):
        built_lib.call("tdb_gn_stats", None, 8, None, 1, 4, 4, 4, 8, 8, 0, None)


def test_no_cpu_path(built_lib):
    from turbdiff_b200 import DenoisingModel
    from turbdiff_b200.models.conditioning import Conditioning

    m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=10, dim=8, u_net_levels=2,
                       norm_type="group")
    with pytest.raises(RuntimeError, match="CUDA"), torch.no_grad():
        m(torch.zeros(1, 4, 16, 8, 8), torch.zeros(1, dtype=torch.long), {Conditioning.Type.CELL_TYPE: torch.zeros(4, 16, 8, 8)})


def test_checkpoint_contract(built_lib):
    from turbdiff_b200 import DenoisingModel, GaussianDiffusion

    ref = json.loads((GOLDEN / "state_dict_layout_shapes.json").read_text())
    with torch.device("meta"):
        m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=500, dim=32, u_net_levels=4,
                           norm_type="group")
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == ref
    gd = GaussianDiffusion(m, timesteps=10)
    assert all(k.startswith("model.") for k in gd.state_dict())  # schedule buffers are non-persistent
    with pytest.raises(RuntimeError, match="Unknown norm type"):
        DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=10, dim=8, u_net_levels=1,
                       norm_type="batch")
    with pytest.raises(ValueError, match="unknown beta schedule"):
        GaussianDiffusion(m, beta_schedule="nope")


@pytest.mark.parametrize("name", ["linear", "log-linear", "log-snr-linear", "cosine", "sigmoid"])
@pytest.mark.parametrize("T", [10, 500, 1000])
def test_schedule_buffers_equal_reference(built_lib, name, T):
    from turbdiff_b200 import GaussianDiffusion

    g = np.load(GOLDEN / "schedules.npz")
    gd = GaussianDiffusion(torch.nn.Identity(), timesteps=T, beta_schedule=name)
    bufs = dict(gd.named_buffers())
    assert len(bufs) == 10
    for b, v in bufs.items():
        np.testing.assert_array_equal(v.numpy(), g[f"{name}/{T}/{b}"], err_msg=b)
        assert v.dtype == torch.float32


def test_time_embedding_buffers(built_lib):
    from turbdiff_b200.models.ddpm import NyquistFrequencyEmbedding

    g = np.load(GOLDEN / "time_embedding.npz")
    np.testing.assert_array_equal(NyquistFrequencyEmbedding(32, 500).scale.numpy(), g["32/500/scale"])


def test_launch_program_is_well_formed(built_lib, monkeypatch):
    """Dry run of the launch program with the kernel calls recorded instead of executed."""
    from turbdiff_b200 import DenoisingModel, _lib, engine
    from turbdiff_b200.models.conditioning import Conditioning

    calls = []
    monkeypatch.setattr(engine, "call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, w: None)
    m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=500, dim=32, u_net_levels=4,
                       norm_type="group", precision="bf16")
    x = torch.zeros(1, 4, 26, 10, 10)
    with torch.no_grad():
        y = m(x, torch.zeros(1, dtype=torch.long), {Conditioning.Type.CELL_TYPE: torch.zeros(4, 26, 10, 10)})
    assert y.shape == x.shape
    n = Counter(c[0] for c in calls)
    # 22 block pointwise + 2 attention pointwise; 8 resamplings
    # 22 3x3x3 + 7 residual 1x1 + 2 attention 1x1 = 31 convolutions; the 1x1 projections of every block that has one
    # (down1-3, up0-3) ride on their conv1 launch (CTA-pair kernels), leaving 24 launches
    assert sum(n[k] for k in n if k.startswith("tdb_conv3d_bf16")) == 24
    assert n["tdb_pointwise"] == 24 and n["tdb_trilinear"] == 8
    # Cout <= 64: down0, up2, up3, decode (two 3x3x3 convs each); the three 32->32 layers stay single-CTA
    # ... and the wide layers of levels 1-3 run as 128-channel N tiles on CTA pairs (down1-3, up0, up1: two convs each)
    # kernels: 64->64 x3 and 128->32 on the kz-folded row-window pair kernel, the ten wide layers of levels 1-3 AND the four
    # 512->512 bottleneck convolutions on the row-window pair kernel with streamed weights, 256->64 on the kz-folded pair
    # kernel, 32->32 x3 single-CTA folded (the paired-row kernel when the input pitch is exactly 32, 128-byte aligned and
    # Z + 2 is even); only the two 1x1 attention projections stay on the per-tap kernel
    assert n["tdb_conv3d_bf16_fold"] + n["tdb_conv3d_bf16_winp"] == 3 and n["tdb_conv3d_bf16_fold2"] == 1 and n["tdb_conv3d_bf16_win"] == 14
    assert n["tdb_conv3d_bf16"] == 2 and n["tdb_gn_stats"] == 1
    assert n["tdb_conv3d_bf16_winz"] == 4
    assert n["tdb_attention"] == 1 and n["tdb_time_film"] == 1 and n["tdb_encode_input"] == 1 and n["tdb_decode_output"] == 1
    # level sizes follow max(int(s/2), 3)
    assert engine.level_sizes((194, 50, 50), 4) == [(194, 50, 50), (97, 25, 25), (48, 12, 12), (24, 6, 6), (12, 3, 3)]


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import it (bench.py may, in its CPU arms only)."""
    import pathlib
    import re

    root = pathlib.Path(__file__).resolve().parents[1]
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for f in (root / "generative-turbulence_b200").rglob("*.py"):
        assert not pat.search(f.read_text()), f
    bench_src = (root / "bench.py").read_text()
    gpu_arm = bench_src[bench_src.index("def run_ours("):]
    assert not pat.search(gpu_arm), "the GPU arm of bench.py must not touch oracle/"


def test_synthetic_workload_matches_the_survey_numbers():
    from turbdiff_b200 import synthetic

    geo = synthetic.channel_geometry()
    assert geo.padded == (194, 50, 50) and geo.cell_type.shape == geo.padded
    # 192*48*48 cells minus the 12x16x32 pillar
    assert len(geo.cell_idx) == 192 * 48 * 48 - 12 * 16 * 32 == len(set(geo.cell_idx.tolist()))
    assert (geo.cell_type.reshape(-1)[geo.cell_idx] == synthetic.INSIDE).all()
    assert abs(synthetic.conv_flops_per_sample(geo.padded) / 1e9 - 666.2) < 0.1


def test_fused_radam_refuses_cpu_tensors():
    from turbdiff_b200.optim import FusedRAdam

    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    opt = FusedRAdam([p], lr=1e-3)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
    with pytest.raises(ValueError):
        FusedRAdam([p], lr=-1.0)


def _golden_grid_case():
    """Inputs of tests/golden/make_golden.py gen_grid (the reference's OpenFOAMData.grid_embedding fixture)."""
    from oracle import grid_ref

    geo = grid_ref.channel_geometry(cells=(12, 6, 5), hole=((3, 6), (1, 4), (0, 3)), seed=3)
    rng = np.random.Generator(np.random.PCG64(11))
    B, n = 2, len(geo.cell_idx)
    u = rng.standard_normal((B, n, 3)).astype(np.float32)
    p = rng.standard_normal((B, n, 1)).astype(np.float32)
    fixed = [(torch.from_numpy(geo.boundaries["inlets"]), 0, torch.tensor([20.0, 0.0, 0.0])),
             (torch.from_numpy(geo.boundaries["walls"]), 0, torch.tensor([0.0, 0.0, 0.0])),
             (torch.from_numpy(geo.boundaries["outlets"]), 3, torch.tensor([0.0]))]
    return geo, torch.from_numpy(np.concatenate([u, p], axis=-1)), fixed


def _emulate_scatter(samples, cell_idx, nvox, cls, bc_has, bc_val):
    """What tdb_scatter_normalize computes (identity normalisation), in torch on the CPU."""
    B, n, F = samples.shape
    grid = torch.zeros((B, F, nvox))
    grid[:, :, cell_idx] = samples.permute(0, 2, 1)
    fixed = bc_has[cls].bool().t()           # (F, nvox)
    vals = bc_val[cls].t()                    # (F, nvox)
    return torch.where(fixed[None], vals[None].expand(B, -1, -1), grid)


def test_boundary_class_tables_reproduce_reference_grid_embedding(golden):
    """FIXED_VALUE boundary writes (data/ofles.py:233-238) as per-voxel classes: the host tables of
    turbdiff_b200.models.utils.boundary_classes against the reference's grid_embedding golden (bit-exact)."""
    from turbdiff_b200.models.utils import boundary_classes

    geo, samples, fixed = _golden_grid_case()
    cls, bc_has, bc_val = boundary_classes(geo.n_vox, 4, fixed, torch.device("cpu"))
    assert int(cls.max()) <= 127 and bc_has.shape == bc_val.shape == (int(cls.max()) + 1, 4)
    got = _emulate_scatter(samples, torch.from_numpy(geo.cell_idx), geo.n_vox, cls, bc_has, bc_val)
    want = golden["grid"]["grid_embedding"]
    np.testing.assert_array_equal(got.reshape(want.shape).numpy(), want)
    inlet = torch.from_numpy(geo.boundaries["inlets"])
    assert torch.equal(got[0, :3, inlet[0]], torch.tensor([20.0, 0.0, 0.0]))


def test_boundary_class_tables_overlaps_follow_write_order():
    """Two boundaries sharing voxels, and a boundary voxel that is also a cell: the later write wins per channel."""
    from turbdiff_b200.models.utils import boundary_classes

    nvox, F = 40, 4
    a = torch.tensor([3, 4, 5, 6])
    b = torch.tensor([5, 6, 7])
    c = torch.tensor([6, 30])
    fixed = [(a, 0, torch.tensor([1.0, 2.0, 3.0])), (b, 0, torch.tensor([-1.0, -2.0, -3.0])), (c, 3, torch.tensor([9.0])),
             (c, 1, torch.tensor([7.0]))]
    cls, bc_has, bc_val = boundary_classes(nvox, F, fixed, torch.device("cpu"))
    cell_idx = torch.tensor([30, 31, 2])  # voxel 30 is both a cell and a boundary voxel of channels 1 and 3
    samples = torch.arange(1 * 3 * F, dtype=torch.float32).reshape(1, 3, F) + 100
    got = _emulate_scatter(samples, cell_idx, nvox, cls, bc_has, bc_val)
    want = torch.zeros((1, F, nvox))
    want[:, :, cell_idx] = samples.permute(0, 2, 1)
    for idx, f0, v in fixed:  # the reference's sequential writes
        want[:, f0 : f0 + v.numel(), idx] = v[None, :, None]
    assert torch.equal(got, want)


def test_pipeline_host_read_matches_reference_read_data():
    """turbdiff_b200.pipeline.read_channels_last = OpenFOAMDataRepository.read_data (ofles.py:396-418): sorted unique hyperslab
    read, scalar fields get a feature axis, the requested order (with duplicates) is restored."""
    import numpy as np

    from turbdiff_b200.pipeline import read_channels_last, sorted_unique_read_plan

    class H5Like:
        def __init__(self, arr):
            self.arr = arr

        def __getitem__(self, idx):
            idx = np.asarray(idx)
            assert np.all(np.diff(idx) > 0), "h5py requires sorted unique indices"
            return self.arr[idx]

    rng = np.random.default_rng(1)
    u, p = rng.standard_normal((12, 7, 3)).astype(np.float32), rng.standard_normal((12, 7)).astype(np.float32)
    idxs = [9, 3, 3, 0, 11, 9]
    uniq, inv = sorted_unique_read_plan(idxs)
    assert list(uniq) == [0, 3, 9, 11] and list(uniq[inv]) == idxs
    got = read_channels_last([(H5Like(u), 3), (H5Like(p), 1)], idxs)
    want = np.concatenate([u[idxs], p[idxs][..., None]], axis=-1)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    out = np.full((8, 7, 4), 5.0, dtype=np.float32)
    view = read_channels_last([(H5Like(u), 3), (H5Like(p), 1)], idxs, out=out)
    assert np.array_equal(view, want) and np.all(out[6:] == 5.0)


def test_gradient_phases_partition_the_state_dict():
    """backward.grad_phase splits the 139 parameters into the part that is complete after the decoder / up path / centre
    (exchanged while the rest of the backward pass runs) and the late part: down path, encoders and everything fed by the
    timestep conditioning (its gradient sums over ALL blocks, so it is only complete at the very end)."""
    from turbdiff_b200 import DenoisingModel
    from turbdiff_b200.backward import grad_phase

    m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=10, dim=16, u_net_levels=2,
                       norm_type="group")
    names = [n for n, _ in m.named_parameters()]
    late = [n for n in names if grad_phase(n) == 2]
    early = [n for n in names if grad_phase(n) == 1]
    assert sorted(late + early) == sorted(names) and late and early
    assert all("downsampling_blocks" in n or "project_onto_scale_shift" in n or n.startswith(("process_c", "encode_")) for n in late)
    assert not any("downsampling_blocks" in n or "project_onto_scale_shift" in n or n.startswith(("process_c", "encode_")) for n in early)
    # every FiLM projection is late, including those of phase-1 blocks
    assert "decode.0.project_onto_scale_shift.weight" in late and "decode.0.block1.conv.weight" in early
    assert "u_net.center_block.1.fn.fn.to_qkv.weight" in early and "decode.1.weight" in early


def test_weight_layout_jobs_cover_every_convolution_and_specs_match_the_torch_layouts(built_lib):
    """Host logic of the batched weight re-layout (engine.weights / pack_many): the job list names every convolution weight of
    the state dict exactly once, and the shape _pack_spec reserves in the flat buffer for a layout (forward and input-gradient
    form, at the level whose kernel choice it follows) is the shape of the torch permute / flip / cast sequence it replaces."""
    from turbdiff_b200 import DenoisingModel

    m = DenoisingModel(in_features=4, out_features=4, c_local_features=4, c_global_features=0, timesteps=500, dim=32, u_net_levels=4,
                       norm_type="group", precision="bf16")
    eng = m.engine()
    eng._set_geometry((194, 50, 50))
    w = eng.weights()  # CPU parameters: the torch layouts
    convs = {k for k, v in m.state_dict().items() if v.dim() == 5 and k.split(".")[0] not in ("encode_x", "encode_c_local", "decode")
             or (k.startswith("decode.0") and v.dim() == 5)}
    assert len([k for k in w if k not in ("film_w", "film_b", "dgrad")]) == len(convs) == 31
    n = 0
    for name, bp in eng.blocks.items():
        lvl = eng._block_level(name)
        for key, conv in ((f"{name}.conv1", bp.blk.block1.conv), (f"{name}.conv2", bp.blk.block2.conv)) + (((f"{name}.proj", bp.blk.conv),) if bp.has_proj else ()):
            for dgrad in (False, True):
                spec = eng._pack_spec(conv.weight, lvl, dgrad)
                got = eng.pack_conv(conv.weight.detach(), lvl, dgrad)
                assert tuple(got.shape) == tuple(spec[5]) and got.dtype == torch.bfloat16, (key, dgrad)
                assert spec[5][0] * spec[5][1] == conv.weight.numel()
                if not dgrad:
                    assert torch.equal(got, w[key]), key
                n += 1
    assert n == 2 * 29


def test_gc_is_paused_only_for_the_duration_of_a_capture():
    """engine._gc_paused (wrapped around every CUDA-graph capture): cyclic GC off inside, restored afterwards - also when the
    capture raises, and left off if the caller had it off."""
    import gc

    from turbdiff_b200.engine import _gc_paused

    assert gc.isenabled()
    with _gc_paused():
        assert not gc.isenabled()
    assert gc.isenabled()
    with pytest.raises(RuntimeError):
        with _gc_paused():
            raise RuntimeError("capture failed")
    assert gc.isenabled()
    gc.disable()
    try:
        with _gc_paused():
            assert not gc.isenabled()
        assert not gc.isenabled()
    finally:
        gc.enable()
