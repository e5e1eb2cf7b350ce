"""Fused optimiser step of the training path.

The reference trains with ``torch.optim.RAdam(self.parameters(), lr)`` (turbdiff/models/diffusion.py:216) under
Lightning's ``gradient_clip_val: 0.1`` / ``gradient_clip_algorithm: norm`` (config/shapes_experiment.yaml:50-51).
``FusedRAdam`` performs both - ``clip_grad_norm_`` and the RAdam update of all 139 parameter tensors - in two kernel
launches (``tdb_grad_sqnorm``, ``tdb_radam_step``) instead of ~40 launches and ~140 Python-level tensor operations.
Hyper-parameters, state names (``step``, ``exp_avg``, ``exp_avg_sq``) and results match ``torch.optim.RAdam``
(``decoupled_weight_decay=False``), so optimizer state dicts are interchangeable."""

from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import call

CHUNK = 16384  # elements per block


class FusedRAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: float | None = None):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("FusedRAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.last_grad_sqnorm: torch.Tensor | None = None  # device scalar (double): squared total gradient norm before clipping
        self._tables = {}

    def _table(self, gi, params):
        """Device pointer / chunk tables of one parameter group, rebuilt only when an address changed."""
        st = [self.state[p] for p in params]
        # moments are part of the signature: load_state_dict() replaces them without touching the parameters
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr()) for p, s in zip(params, st))
        tb = self._tables.get(gi)
        if tb is not None and tb["sig"] == sig:
            return tb
        dev = params[0].device
        ptrs = [[p.data_ptr() for p in params], [p.grad.data_ptr() for p in params], [s["exp_avg"].data_ptr() for s in st],
                [s["exp_avg_sq"].data_ptr() for s in st], [p.numel() for p in params]]
        chunk_tensor, chunk_off = [], []
        for i, p in enumerate(params):
            for off in range(0, p.numel(), CHUNK):
                chunk_tensor.append(i)
                chunk_off.append(off)
        tb = {"sig": sig, "ptrs": torch.tensor(ptrs, dtype=torch.int64).to(dev, non_blocking=True),
              "chunk_tensor": torch.tensor(chunk_tensor, dtype=torch.int32).to(dev, non_blocking=True),
              "chunk_off": torch.tensor(chunk_off, dtype=torch.int64).to(dev, non_blocking=True), "n_chunks": len(chunk_tensor)}
        self._tables[gi] = tb
        return tb

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}  # the moment tensors were replaced

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = []
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                _lib.require_cuda(p, "FusedRAdam parameter")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedRAdam: parameters and gradients must be contiguous float32 tensors")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            groups.append((gi, group, params, self._table(gi, params)))
        if not groups:
            return loss
        s = _lib.stream_ptr()
        sq = None
        if self.max_grad_norm is not None:
            # clip_grad_norm_ over ALL parameters of all groups (what Lightning does before optimizer.step)
            sq = torch.zeros(1, dtype=torch.float64, device=groups[0][2][0].device)
            for _, _, _, tb in groups:
                P = tb["ptrs"]
                call("tdb_grad_sqnorm", P[1].data_ptr(), P[4].data_ptr(), tb["chunk_tensor"].data_ptr(), tb["chunk_off"].data_ptr(),
                     tb["n_chunks"], CHUNK, sq.data_ptr(), s)
            self.last_grad_sqnorm = sq
        for _, group, params, tb in groups:
            beta1, beta2 = group["betas"]
            st0 = self.state[params[0]]
            t = int(st0["step"].item()) + 1  # CPU scalar: all tensors of a group step together
            for p in params:
                self.state[p]["step"] += 1
            bc1 = 1.0 - beta1**t
            bc2 = 1.0 - beta2**t
            rho_inf = 2.0 / (1.0 - beta2) - 1.0
            rho_t = rho_inf - 2.0 * t * beta2**t / bc2
            if rho_t > 5.0:
                rect = math.sqrt((rho_t - 4.0) * (rho_t - 2.0) * rho_inf / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t))
                step_size, rectified = group["lr"] * rect * math.sqrt(bc2) / bc1, 1
            else:
                step_size, rectified = group["lr"] / bc1, 0
            P = tb["ptrs"]
            call("tdb_radam_step", P[0].data_ptr(), P[1].data_ptr(), P[2].data_ptr(), P[3].data_ptr(), P[4].data_ptr(),
                 tb["chunk_tensor"].data_ptr(), tb["chunk_off"].data_ptr(), tb["n_chunks"], CHUNK, None if sq is None else sq.data_ptr(),
                 float(self.max_grad_norm or 0.0), float(step_size), float(beta1), float(beta2), float(group["eps"]),
                 float(group["weight_decay"]), rectified, s)
            # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on
            # Tensor._version, e.g. the engine's kernel-layout weights) that they changed, like an in-place torch op would
            torch.autograd.graph.increment_version(params)
        return loss
