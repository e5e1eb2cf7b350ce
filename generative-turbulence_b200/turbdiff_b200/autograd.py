"""torch.autograd glue: pairs the forward launch programs with their backward programs."""

from __future__ import annotations

import torch

from . import _lib
from ._lib import call


class _MaskedLoss(torch.autograd.Function):
    """mean_b mean_{f, inside cells} |eps - noise|^p and d/d eps in one pass (ddpm.py:845-852)."""

    @staticmethod
    def forward(ctx, eps, noise, mask, n_inside, l1):
        eps = eps.contiguous()
        noise = noise.contiguous()
        B, F = eps.shape[:2]
        nvox = eps[0, 0].numel()
        acc = torch.zeros(1, dtype=torch.float64, device=eps.device)
        grad = torch.empty_like(eps) if eps.requires_grad else None
        call("tdb_masked_loss", eps.data_ptr(), noise.data_ptr(), mask.data_ptr(), acc.data_ptr(), _lib.ptr(grad), B, F, nvox,
             n_inside, 1 if l1 else 0, _lib.stream_ptr())
        ctx.grad = grad
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None, None, None, None


def masked_loss(eps, noise, mask, n_inside: int, l1: bool):
    return _MaskedLoss.apply(eps, noise, mask, n_inside, l1)


class _Denoise(torch.autograd.Function):
    """eps = U-Net(x, t, c_local): forward launch program (training mode keeps the per-block
    intermediates), backward launch program (turbdiff_b200.backward).  x itself gets no gradient:
    the reference never differentiates with respect to the noisy input."""

    @staticmethod
    def forward(ctx, model, x, t, c_local, *params):
        ctx.model = model
        ctx.n_params = len(params)
        return model.engine().train_forward(x, t, c_local).clone()

    @staticmethod
    def backward(ctx, g_eps):
        model = ctx.model
        eng = model.engine()
        from .backward import grad_phase

        grads, g_c_local = eng.train_backward(g_eps)
        g_c = None if g_c_local is None else g_c_local.clone()
        named = list(model.named_parameters())
        if torch.is_tensor(grads):
            # graph replay: one flat buffer (phase-1 parameters first); a single copy detaches it from the graph's static
            # memory.  The data-parallel exchange (parallel.GradientAllReduce.attach) was started by train_backward - the
            # phase-1 part already while the second backward graph ran - and is completed here.
            tg = eng._train_replay
            if eng.grad_sync is not None:
                eng.grad_sync.finish()
            piece = {n: v.view(shape) for n, v, shape in zip(tg["order"], grads.clone().split(tg["sizes"]), tg["shapes"])}
            return (None, None, None, g_c, *(piece[n] for n, _ in named))
        for name, _ in named:
            if grads.get(name) is None:
                raise RuntimeError(f"turbdiff_b200: no gradient produced for parameter {name}")
        if eng.grad_sync is not None:
            # eager launch programs under data parallelism: the SAME exchange as the graph path (flat fp32 buffers of the
            # phase-1 and phase-2 parameters, same chunking), so ranks may mix the two modes (a rank whose graph capture
            # fell back to the eager programs must still issue the collectives its peers issue)
            out = {}
            for phase in (1, 2):
                part = [(n, prm) for n, prm in named if grad_phase(n) == phase]
                flat = torch.cat([grads[n].reshape(-1).to(torch.float32) for n, _ in part])
                eng.grad_sync.start(flat)
                for v, (n, prm) in zip(flat.split([prm.numel() for _, prm in part]), part):
                    out[n] = v.view(prm.shape).to(prm.dtype)
            eng.grad_sync.finish()
            return (None, None, None, g_c, *(out[n] for n, _ in named))
        out = [grads[name].to(prm.dtype).reshape(prm.shape).clone() for name, prm in named]  # the program's buffers are reused by the next step
        return (None, None, None, g_c, *out)


def denoise_with_grad(model, x, t, c_local):
    _lib.require_cuda(x, "x")
    params = [p for _, p in model.named_parameters()]
    return _Denoise.apply(model, x, t, c_local, *params)
