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

import contextlib
import os

from . import modules as _modules
from .modules import BCEWithLogitsLoss, L1Loss
from .optim import FusedAdamW

RECON_FACTOR = 1e2   # ref:src/model.py:147
N_RECON_TERMS = 2    # ref:src/model.py:138 (L1, Perceptual)


def _set_requires_grad(module, flag):
    for p in module.parameters():
        p.requires_grad_(flag)


def _dist_on(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def broadcast_module_state(module, group=None, src=0, buffers_only=False):
    """What DDP does at construction (parameters + buffers) and before every forward
    (``broadcast_buffers=True``: the BatchNorm running statistics, C4): rank ``src``'s tensors replace everyone's,
    in ONE broadcast per dtype on a flat staging buffer."""
    if not _dist_on(group):
        return
    tensors = list(module.buffers()) if buffers_only else list(module.parameters()) + list(module.buffers())
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    with torch.no_grad():
        for dtype, ts in by_dtype.items():
            flat = torch.cat([t.detach().reshape(-1) for t in ts])
            dist.broadcast(flat, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
            off = 0
            for t in ts:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))     # in place: same storage, version bumped
                off += n


class GradAllReducer:
    """Averages the gradients of one network over the data-parallel ranks with ONE NCCL all-reduce on a persistent
    flat fp32 buffer with a FIXED layout over all of the module's parameters (what DDP's reducer does for the live
    network's slots, ref:src/train.py:30-32 ``ddp_find_unused_parameters_true``): a parameter without a gradient on
    this rank contributes zeros, so ranks never disagree about the payload. After the call every ``p.grad`` that any
    rank produced is a VIEW of the reduced buffer holding the SUM; ``scale`` (1 / world size) is handed to the
    optimizer (``FusedAdamW.grad_scale``) or applied here for foreign optimizers."""

    def __init__(self, module, group=None, check_unused=False):
        self.params = [p for p in module.parameters()]
        self.group = group
        # check_unused: agree across ranks on WHICH parameters received a gradient (one more small all-reduce and a
        # host sync per call). Off by default: in this model the set is a function of the phase, not of the data, so
        # every rank has the same one and the local answer is the global one.
        self.check_unused = check_unused
        self.flat = None
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += p.numel()
        self.numel = off

    def __call__(self, apply_scale=True):
        """-> the factor still to be applied to the gradients (1.0 when ``apply_scale`` or single-process)."""
        if not _dist_on(self.group):
            return 1.0
        world = dist.get_world_size(self.group)
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.empty(self.numel, dtype=torch.float32, device=dev)
        views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        have = [0.0 if p.grad is None else 1.0 for p in self.params]
        live = [(v, p.grad) for v, p in zip(views, self.params) if p.grad is not None]
        dead = [v for v, p in zip(views, self.params) if p.grad is None]
        if live:
            torch._foreach_copy_([v for v, _ in live], [g for _, g in live])
        if dead:
            torch._foreach_zero_(dead)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)   # SUM + scale: works on nccl and gloo
        if self.check_unused:
            have_t = torch.tensor(have, device=dev)
            dist.all_reduce(have_t, op=dist.ReduceOp.MAX, group=self.group)  # which slots ANY rank produced
            have = have_t.tolist()
        if apply_scale:
            self.flat.mul_(1.0 / world)
        for v, p, h in zip(views, self.params, have):
            if h > 0:
                p.grad = v                  # alias the reduced buffer: no copy back
        return 1.0 if apply_scale else 1.0 / world


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
        # DDP construction semantics: every rank starts from rank 0's parameters and buffers (C6)
        broadcast_module_state(gen)
        broadcast_module_state(discr)
        self.broadcast_buffers = True       # DDP default: rank 0's BatchNorm running statistics before every step (C4)

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

    def _reduce_and_step(self, reducer, opt):
        fused = isinstance(opt, FusedAdamW)
        scale = reducer(apply_scale=not fused)
        if fused:
            opt.grad_scale = scale          # 1 / world folded into the AdamW kernel's gradient read
        opt.step()
        opt.zero_grad(set_to_none=True)

    def step(self, x, y):
        """One ``training_step``: G phase (D frozen), AdamW, D phase on a fresh G forward, AdamW."""
        if self.broadcast_buffers and _dist_on():
            broadcast_module_state(self.gen, buffers_only=True)
            broadcast_module_state(self.discr, buffers_only=True)
        # the discriminator's parameters are constant from here to opt_d.step(): its six passes share two operand packs
        hold = _modules.weights_unchanged(self.discr) if isinstance(self.discr, _modules.Discriminator) else contextlib.nullcontext()
        with hold:
            _set_requires_grad(self.discr, False)
            g_loss, _ = self.gen_loss(x, y)
            g_loss.backward()
            self._reduce_and_step(self.reduce_g, self.opt_g)
            _set_requires_grad(self.discr, True)

            _set_requires_grad(self.gen, False)
            d_loss = self.discr_loss(x, y)
            d_loss.backward()
            if self._branch_stream is not None:
                torch.cuda.current_stream().wait_stream(self._branch_stream)   # the real pass's backward ran there
        self._reduce_and_step(self.reduce_d, self.opt_d)
        _set_requires_grad(self.gen, True)
        return g_loss.detach(), d_loss.detach()
