"""FusedRAdam (tdb_grad_sqnorm + tdb_radam_step) against torch.optim.RAdam + torch.nn.utils.clip_grad_norm_, the
optimiser the reference trains with (turbdiff/models/diffusion.py:216, config/shapes_experiment.yaml:50-51)."""

import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 32, 3, 3, 3), (32,), (1,), (7, 5), (128, 64, 3, 3, 3), (16385,), (3, 16384)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.1).cuda()) for s in SHAPES]


@pytest.mark.parametrize("max_norm", [None, 0.1, 300.0, 1e6])  # ||g|| ~ 570*(1+step): 300 clips every step without crushing the update
@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_radam_matches_torch(max_norm, weight_decay):
    from turbdiff_b200.optim import FusedRAdam

    ref_p, our_p = _params(0), _params(0)
    ref = torch.optim.RAdam(ref_p, lr=1e-3, weight_decay=weight_decay)
    ours = FusedRAdam(our_p, lr=1e-3, weight_decay=weight_decay, max_grad_norm=max_norm)
    g = torch.Generator().manual_seed(1)
    for step in range(9):  # the rectified branch (rho_t > 5) starts at step 6 with beta2 = 0.999
        grads = [(torch.randn(*s, generator=g) * (1.0 + step)).cuda() for s in SHAPES]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        if max_norm is not None:
            total = torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
        ref.step()
        ours.step()
        if max_norm is not None:
            assert torch.allclose(ours.last_grad_sqnorm.sqrt().float(), total, rtol=1e-5)
        for i, (p, q) in enumerate(zip(ref_p, our_p)):
            torch.testing.assert_close(q, p, rtol=2e-6, atol=1e-7, msg=lambda m: f"step {step} tensor {i}: {m}")
    for p, q in zip(ref_p, our_p):
        sr, so = ref.state[p], ours.state[q]
        assert float(sr["step"]) == float(so["step"]) == 9.0
        # moments: fused multiply-adds here, separate mul/add kernels in torch - a few ulp per step, accumulated over 9 steps
        torch.testing.assert_close(so["exp_avg"], sr["exp_avg"], rtol=2e-5, atol=5e-6)
        torch.testing.assert_close(so["exp_avg_sq"], sr["exp_avg_sq"], rtol=1e-4, atol=1e-6)


def test_fused_radam_state_dict_is_interchangeable_with_torch():
    from turbdiff_b200.optim import FusedRAdam

    a, b = _params(2), _params(2)
    ours = FusedRAdam(a, lr=2e-3)
    for p in a:
        p.grad = torch.ones_like(p)
    ours.step()
    ref = torch.optim.RAdam(b, lr=2e-3)
    ref.load_state_dict(copy.deepcopy(ours.state_dict()))  # state_dict() hands out references to the live state tensors
    for p, q in zip(a, b):
        q.data.copy_(p.data)
        p.grad = torch.full_like(p, 0.5)
        q.grad = torch.full_like(q, 0.5)
    ours.step()
    ref.step()
    for p, q in zip(a, b):
        torch.testing.assert_close(p, q, rtol=2e-6, atol=1e-7)
