"""GAN optimisation step with the reference's ``training_step`` semantics (ref:src/model.py:259-281,
170-193) on the sm_100a modules, plus the data-parallel gradient all-reduce Lightning DDP performs
inside ``manual_backward`` (ref:src/train.py:30-32).

The perceptual term of ``PerceptualL1Loss`` is out of scope (its MedicalNet weights need the network,
SURVEY.md section 2 row 8) and is fixed to 0 while keeping the reference's averaging over its two
loss terms: ``recon = (L1 + 0) / 2 * recon_factor`` (ref:src/model.py:209).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

import os

from . import modules as _modules
from .modules import BCEWithLogitsLoss, L1Loss
from .optim import FusedAdamW

RECON_FACTOR = 1e2   # ref:src/model.py:147
N_RECON_TERMS = 2    # ref:src/model.py:138 (L1, Perceptual)


def _set_requires_grad(module, flag):
    for p in module.parameters():
        p.requires_grad_(flag)


class GradAllReducer:
    """Averages the gradients of one network over the data-parallel ranks with ONE NCCL all-reduce on
    a persistent flat fp32 buffer (what DDP's bucketed reducer does for the live network's slots)."""

    def __init__(self, module, group=None):
        self.params = [p for p in module.parameters()]
        self.group = group
        self.flat = None

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        n = sum(p.grad.numel() for p in live)
        if self.flat is None or self.flat.numel() != n or self.flat.device != live[0].grad.device:
            self.flat = torch.empty(n, dtype=torch.float32, device=live[0].grad.device)
        views, off = [], 0
        for p in live:
            views.append(self.flat[off:off + p.grad.numel()].view_as(p.grad))
            off += p.grad.numel()
        torch._foreach_copy_(views, [p.grad for p in live])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)   # SUM + scale: works on nccl and gloo
        self.flat.mul_(1.0 / dist.get_world_size(self.group))
        torch._foreach_copy_([p.grad for p in live], views)


class GanTrainer:
    """Holds the two networks, their AdamW optimisers (ref:src/model.py:359-361) and the loss modules."""

    def __init__(self, gen, discr, lr=1e-3, optimizer="ub", fused_optimizer=None):
        """``optimizer``: "ub" = ``FusedAdamW`` on the sm_100a multi-tensor kernel (default), "torch_fused" /
        "torch" = ``torch.optim.AdamW`` with / without ``fused=True`` (what Lightning would build from the
        reference's ``configure_optimizers``; kept for A/B and for the weight-cache regression tests)."""
        self.gen, self.discr = gen, discr
        if fused_optimizer is not None:            # older spelling
            optimizer = "torch_fused" if fused_optimizer else "torch"
        if optimizer == "ub":
            self.opt_g = FusedAdamW(gen.parameters(), lr=lr)
            self.opt_d = FusedAdamW(discr.parameters(), lr=lr)
        else:
            kw = {"fused": True} if optimizer == "torch_fused" else {}
            self.opt_g = torch.optim.AdamW(gen.parameters(), lr=lr, **kw)
            self.opt_d = torch.optim.AdamW(discr.parameters(), lr=lr, **kw)
        # D phase: the real-sample pass of the discriminator does not depend on the generator, so it CAN run (forward
        # and, through autograd's stream affinity, backward) on a second stream beside the generator forward and the
        # fake-sample pass. Measured on B200 at 8 x 128^3: 68.6 vs 68.5 ms/step -- the tcgen05 kernels of both
        # passes want whole SMs (shared memory, TMEM), so they interleave instead of co-running. Off by default;
        # UB_OVERLAP_REAL_BRANCH=1 enables it (bit-identical results, tests/test_model_gpu.py).
        self.overlap_real_branch = os.environ.get("UB_OVERLAP_REAL_BRANCH", "0") == "1"
        self._branch_stream = None
        self.l1 = L1Loss()
        self.bce = BCEWithLogitsLoss()
        self.reduce_g = GradAllReducer(gen)
        self.reduce_d = GradAllReducer(discr)

    def recon_loss(self, y_hat, y):
        return self.l1(y_hat, y) / N_RECON_TERMS * RECON_FACTOR

    def gen_loss(self, x, y):
        """ref:src/model.py:170-181."""
        y_hat = self.gen(x)
        logits = self.discr(x, y_hat)
        adv = self.bce(logits, torch.ones_like(logits))
        return adv + self.recon_loss(y_hat, y), y_hat

    def _can_overlap(self, x):
        return (self.overlap_real_branch and x.is_cuda and isinstance(self.discr, _modules.Discriminator)
                and _modules._precision_of(self.discr) == "bf16")

    def discr_loss(self, x, y):
        """ref:src/model.py:183-193."""
        if not self._can_overlap(x):
            with torch.no_grad():
                y_hat = self.gen(x)
            logits_hat = self.discr(x, y_hat)
            logits = self.discr(x, y)
            loss_hat = self.bce(logits_hat, torch.zeros_like(logits_hat))
            loss = self.bce(logits, torch.ones_like(logits))
            return (loss + loss_hat) / 2
        # Same arithmetic, two streams. BatchNorm running statistics keep the reference's order (fake pass, then
        # real pass): the real pass defers its updates until the fake pass has applied its own.
        main = torch.cuda.current_stream()
        if self._branch_stream is None or self._branch_stream.device != x.device:
            self._branch_stream = torch.cuda.Stream(device=x.device)
        side = self._branch_stream
        _modules.prepack_weights(self.discr)          # on the main stream: both passes read the same operand copies
        side.wait_stream(main)
        deferred = []
        with torch.cuda.stream(side), _modules.defer_bn_running_stats(deferred):
            logits = self.discr(x, y)
            loss = self.bce(logits, torch.ones_like(logits))
        x.record_stream(side)
        y.record_stream(side)
        with torch.no_grad():
            y_hat = self.gen(x)
        logits_hat = self.discr(x, y_hat)
        loss_hat = self.bce(logits_hat, torch.zeros_like(logits_hat))
        main.wait_stream(side)
        _modules.apply_deferred_bn(deferred)
        loss.record_stream(main)
        return (loss + loss_hat) / 2

    def step(self, x, y):
        """One ``training_step``: G phase (D frozen), AdamW, D phase on a fresh G forward, AdamW."""
        _set_requires_grad(self.discr, False)
        g_loss, _ = self.gen_loss(x, y)
        g_loss.backward()
        self.reduce_g()
        self.opt_g.step()
        self.opt_g.zero_grad(set_to_none=True)
        _set_requires_grad(self.discr, True)

        _set_requires_grad(self.gen, False)
        d_loss = self.discr_loss(x, y)
        d_loss.backward()
        if self._branch_stream is not None:
            torch.cuda.current_stream().wait_stream(self._branch_stream)   # the real pass's backward ran there
        self.reduce_d()
        self.opt_d.step()
        self.opt_d.zero_grad(set_to_none=True)
        _set_requires_grad(self.gen, True)
        return g_loss.detach(), d_loss.detach()
