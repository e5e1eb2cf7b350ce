"""Drop-in test through the reference's OWN task (turbdiff/models/diffusion.py:41-242): the unmodified
``DiffusionTraining`` is constructed twice - once as is (CPU fp32: the oracle) and once with the one-line import swap
of INTEGRATION.md section 2 applied (``DenoisingModel`` / ``GaussianDiffusion`` from turbdiff_b200, on the GPU) - and
``training_step``, ``loss.backward()``, the configured optimiser step and ``sample`` are compared on a synthetic batch
built from the reference's own dataclasses (grid_embedding with FIXED_VALUE boundaries, Normalization, Conditioning
with the learned cell-type embedding all run unmodified on both sides).

The reference package comes from /root/reference (build container) or oracle/_ref (GPU box); pytorch_lightning & co.
are stubbed by oracle/ref_shim.py."""

import copy

import pytest
import torch

from util import cpu_seeded_randn, rel_l2

pytestmark = pytest.mark.gpu

KW = dict(dim=16, cell_type_embedding_type="learned", cell_type_embedding_dim=4, normalization_mode="u:norm-max;p:abs-max",
          beta_schedule="log-snr-linear", timesteps=10, learning_rate=3e-3, min_learning_rate=1e-6, lr_decay="exp", loss="l2",
          noise_bcs=True, optimizer="radam", norm_type="group", with_geometry_embedding=False)


@pytest.fixture(scope="module")
def ns():
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference package not installed (oracle/install_ref.py)")
    return ref_shim.load(with_task=True)


def _to(batch, ns, device):
    """Copy of an OpenFOAMBatch on `device` (what Lightning's move_data_to_device does)."""
    d, st = batch.data, batch.stats
    md = copy.copy(d.metadata)
    md.cell_idx = md.cell_idx.to(device)
    md.boundaries = {k: {**v, "idx": v["idx"].to(device)} for k, v in md.boundaries.items()}
    BC = ns.ofles.BoundaryCondition
    md.boundary_conditions = {v: {n: BC(bc.type, None if bc.value is None else bc.value.to(device)) for n, bc in bcs.items()}
                              for v, bcs in md.boundary_conditions.items()}
    md._inside_mask = md._unpadded_cell_idx = None
    data = ns.ofles.OpenFOAMData(md, d.t.to(device), {v: s.to(device) for v, s in d.samples.items()})
    stats = ns.ofles.OpenFOAMStats({k: {n: t.to(device) for n, t in v.items()} for k, v in st.stats.items()})
    return ns.ofles.OpenFOAMBatch(data, stats)


@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 2e-5, 5e-4), ("bf16", 2e-2, 6e-2)])
def test_reference_task_with_the_import_swap(ns, tmp_path, monkeypatch, precision, tol, gtol):
    import turbdiff_b200.models.ddpm as fast
    from oracle.ref_batch import make_batch

    V = ns.ofles.Variable
    batch, geo = make_batch(ns, cells=(32, 16, 16), batch=2, seed=5)

    torch.manual_seed(0)
    ref_task = ns.diffusion.DiffusionTraining(data_dir=tmp_path, samples_root=tmp_path / "ref", variables=(V.U, V.P), **KW)

    # ---- the import swap: nothing else of the reference changes
    monkeypatch.setenv("TURBDIFF_B200_PRECISION", precision)
    monkeypatch.setattr(ns.diffusion, "DenoisingModel", fast.DenoisingModel)
    monkeypatch.setattr(ns.diffusion, "GaussianDiffusion", fast.GaussianDiffusion)
    torch.manual_seed(0)
    task = ns.diffusion.DiffusionTraining(data_dir=tmp_path, samples_root=tmp_path / "fast", variables=(V.U, V.P), **KW)
    assert isinstance(task.model, fast.GaussianDiffusion) and isinstance(task.model.model, fast.DenoisingModel)
    # same names, shapes AND initial values (same registration / initialisation order under the same seed)
    sd_ref, sd = ref_task.state_dict(), task.state_dict()
    assert list(sd) == list(sd_ref)
    for k in sd:
        assert torch.equal(sd[k], sd_ref[k]), k
    task.load_state_dict(sd_ref, strict=True)  # the checkpoint path of scripts/eval_ckpt.py:43-60
    task = task.cuda()
    gbatch = _to(batch, ns, "cuda")

    # ---- training_step + backward (diffusion.py:160-165)
    torch.manual_seed(11)
    ref_loss = ref_task.training_step(batch, 0)["loss"]
    ref_loss.backward()
    with cpu_seeded_randn(11):
        loss = task.training_step(gbatch, 0)["loss"]
    loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss.detach())) < tol * abs(float(ref_loss.detach()))
    ref_grads = dict(ref_task.named_parameters())
    errs = {}
    gmax = max(float(q.grad.abs().max()) for q in ref_grads.values())
    for k, p in task.named_parameters():
        assert p.grad is not None, k
        g = ref_grads[k].grad
        if float(g.abs().max()) < 1e-6 * gmax:
            # a conv bias in front of a GroupNorm has an exactly zero gradient: the reference holds rounding noise there
            assert float(p.grad.abs().max()) < 1e-4 * gmax, k
            continue
        errs[k] = rel_l2(p.grad, g)
    assert "cell_type_embedding.embedding.weight" in errs  # the gradient reaches the conditioning's 24 weights through C
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print("drop-in gradients", precision, [(k, f"{v:.2e}") for k, v in worst])
    assert worst[0][1] < gtol, worst

    # ---- the task's own optimiser (configure_optimizers: RAdam + exp decay, diffusion.py:210-235), one step each
    for t_ in (ref_task, task):
        cfg = t_.configure_optimizers()
        cfg["optimizer"].step()
        cfg["lr_scheduler"]["scheduler"].step()
    for (k, p), q in zip(task.named_parameters(), ref_task.parameters()):
        assert rel_l2(p, q) < (1e-5 if precision == "fp32" else 2e-2), k

    # ---- sample (diffusion.py:152-158): grid_embedding -> normalise -> p_sample_loop -> de-normalise, updated weights
    ref_task.eval()
    task.eval()
    torch.manual_seed(12)
    with torch.no_grad():
        want = ref_task.sample(batch, start_from=4)
    with cpu_seeded_randn(12), torch.no_grad():
        got = task.sample(gbatch, start_from=4)
    err = rel_l2(got, want)
    print("drop-in sample", precision, err)
    assert got.shape == want.shape == (2, 4, *geo.padded)
    assert err < tol
    assert task.model.model.engine().graph_fallbacks == 0
