"""AdamW on the sm_100a multi-tensor kernel.

Drop-in for the ``torch.optim.AdamW(params, lr=self.lr)`` the reference builds in ``configure_optimizers``
(ref:src/model.py:144,164,359-361; torch defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0.01): same
constructor arguments, same update arithmetic (``torch.optim.adamw._single_tensor_adamw``), same state keys
(``step``, ``exp_avg``, ``exp_avg_sq``), so optimizer state dicts interchange with torch's. All parameters of
a group are updated by ``ub_adamw_step`` in ceil(n / 48) launches; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import AdamWTensor

__all__ = ["FusedAdamW"]


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = 1.0      # multiplies every gradient inside the kernel (1 / world_size after a SUM all-reduce)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            keep = []
            by_step = {}           # torch counts steps per parameter: one launch set per distinct count
            for p in live:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW updates contiguous fp32 CUDA parameters only (there is no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                by_step.setdefault(int(st["step"]), []).append(
                    AdamWTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
            b1, b2 = group["betas"]
            for step, entries in by_step.items():
                table = (AdamWTensor * len(entries))(*entries)
                _lib.check(lib.ub_adamw_step(table, len(entries), float(group["lr"]), float(b1), float(b2),
                                             float(group["eps"]), float(group["weight_decay"]), step,
                                             float(self.grad_scale), stream), "ub_adamw_step")
        return loss
