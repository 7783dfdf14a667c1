"""Drop-in ``torch.nn.Module`` mirror of the reference's hot-path classes, executing on sm_100a kernels.

Same constructor signatures, attribute names and state-dict keys as ref:src/model.py:15-92 (and the
monai==1.3.0 ``BasicUNet`` tree it instantiates, ref:src/model.py:22-28), so checkpoints interchange
and ``bSSFPToDWITensorModel`` (ref:src/model.py:141-361) can use these classes unchanged:

    Generator(input_modality)(x)            -> (B, 6, D, H, W)
    Discriminator(modality)(x, y)           -> (B, 1, D/32, H/32, W/32)
    DownSampleConv(in, out, kernel, strides, padding, activation, batchnorm)(x)
    L1Loss()(a, b), BCEWithLogitsLoss()(logits, target)

The ``torch.nn`` layer objects inside (Conv3d, BatchNorm3d, InstanceNorm3d, ...) are parameter and
buffer *containers* only (names, shapes, default init, ``.to()``, ``state_dict``): their ``forward``
is never called. Each network runs as ONE ``torch.autograd.Function`` whose forward/backward enqueue
the kernels of ``libubssfp.so`` (no cuDNN, no CPU fallback); gradients reach ``Parameter.grad``
through autograd, so DDP hooks and ``requires_grad`` toggling (ref:src/model.py:264,274) behave as
with the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import (UB_CONV_K1, UB_CONV_K3S1P1, UB_CONV_K4S2P1, UB_CONV_K4S2P1_S2D, UB_DECONV_K2S2,
                   UB_NORM_BATCH_EVAL, UB_NORM_BATCH_TRAIN, UB_NORM_INSTANCE, UB_NORM_NONE)

UNET_FEATURES = (32, 64, 128, 256, 512, 32)

# ---------------------------------------------------------------------------------------------------
# precision: "bf16" = the tcgen05 tensor-core path (default, the throughput path); "fp32" = the CUDA-core
# verification path of fp32_mode.py (1e-5 against the reference, small volumes). Per module
# (``module.precision`` / ``set_precision``) with the process default from UB_PRECISION.
# ---------------------------------------------------------------------------------------------------
import os as _os_env

_DEFAULT_PRECISION = _os_env.environ.get("UB_PRECISION", "bf16")
_PRECISIONS = ("bf16", "fp32")


def _precision_of(module) -> str:
    p = getattr(module, "precision", None) or _DEFAULT_PRECISION
    if p not in _PRECISIONS:
        raise ValueError(f"precision must be one of {_PRECISIONS}, got {p!r}")
    return p


def set_precision(module: nn.Module, precision: str):
    """Select the arithmetic of every hot-path module inside ``module``: "bf16" or "fp32"."""
    if precision not in _PRECISIONS:
        raise ValueError(f"precision must be one of {_PRECISIONS}, got {precision!r}")
    for m in module.modules():
        if isinstance(m, (Generator, Discriminator, DownSampleConv, BasicUNet)):
            m.precision = precision
    return module


# ---------------------------------------------------------------------------------------------------
# packed weights: bf16 operand copies of the fp32 Parameters. There is NO cross-call cache: every forward and every
# backward pass of a network re-packs the weights it is about to use from the LIVE parameter memory, all of them in
# one multi-tensor launch (``ub_pack_conv_weights_multi``: ~25 us for the generator's 22.6 M weights). The
# reference reads its fp32 weights afresh in every call; so does this -- optimizer steps (fused or not),
# ``load_state_dict``, ``p.data.copy_()``, EMA swaps or a foreign kernel writing parameter memory can never leave a
# stale operand behind. (Round 1 kept a version-keyed cache, which out-of-band ``.data`` writes defeated.)
# ---------------------------------------------------------------------------------------------------
def invalidate_packed_weights(module: nn.Module | None = None):
    """Kept for API compatibility: operand copies are rebuilt on every pass, there is nothing to invalidate."""
    return None


class _PackedWeights:
    """Packed operands of the CURRENT pass of one network: ``refresh`` fills it, ``get`` reads it."""

    def __init__(self):
        self._cur = {}
        self.hold = False      # inside ``weights_unchanged``: operand copies packed in the scope stay valid until it ends

    def refresh(self, blocks, direction, f16_src0=frozenset()):
        """Re-pack the conv weights of ``blocks`` for ``direction`` (0 forward, 1 dgrad) from the live parameters.
        ``f16_src0``: names of blocks whose source 0 arrives as an fp16-operand deferred activation in this pass
        (their source-0 weight columns are packed as fp16, ``UB_PACK_F16_SRC0``)."""
        todo, seen = [], set()
        for b in blocks:
            key = (id(b.conv.weight), b.spec, direction)
            if key not in seen and not (self.hold and key in self._cur):
                seen.add(key)
                todo.append((b.spec, b.conv.weight, direction | (ops.UB_PACK_F16_SRC0 if b.name in f16_src0 else 0)))
        if not self.hold:
            for k in [k for k in self._cur if k[2] == direction]:
                del self._cur[k]
        if not todo:
            return
        outs = ops.pack_conv_weights_multi(todo)
        for (spec, w, _), out in zip(todo, outs):
            self._cur[(id(w), spec, direction)] = out       # valid for THIS pass (fp16 columns included, if any)

    def get(self, spec, weight, direction):
        hit = self._cur.get((id(weight), spec, direction))
        if hit is None:      # a block run outside a refreshed pass: pack it alone, do not keep it
            return ops.pack_conv_weights(spec, weight, direction)
        return hit

    def release(self, direction):
        for k in [k for k in self._cur if k[2] == direction]:
            del self._cur[k]


# ---------------------------------------------------------------------------------------------------
# one convolution (+ optional norm, dropout, LeakyReLU, max-pool) of the network graph
# ---------------------------------------------------------------------------------------------------
class _Block:
    """conv -> [InstanceNorm | BatchNorm] -> [Dropout] -> [LeakyReLU] (-> MaxPool3d(2))."""

    def __init__(self, name, spec, conv, norm=None, norm_kind=None, slope=1.0, drop_mod=None, fused_act=False,
                 s2d_convert=False):
        self.name = name
        self.spec = spec
        self.s2d_convert = s2d_convert   # spec.kind is UB_CONV_K4S2P1_S2D and the source arrives as a plain NDHWC tensor:
                                         # the forward makes the space-to-depth copy (ops.to_s2d) and saves THAT
        self.conv = conv            # nn.Conv3d / nn.ConvTranspose3d container
        self.norm = norm            # nn.InstanceNorm3d / nn.BatchNorm3d container or None
        self.norm_kind = norm_kind  # "instance" | "batch" | None
        self.slope = slope
        self.drop_mod = drop_mod    # the block's nn.Dropout (monai ADN "D") or None: p and train/eval are read per call
        self.fused_act = fused_act  # activation applied in the conv epilogue (no norm in between)

    def dropout_p(self):
        """Dropout probability in effect NOW (0 in eval mode), as ``nn.Dropout.forward`` would use it."""
        d = self.drop_mod
        return float(d.p) if (d is not None and d.training) else 0.0

    def params(self):
        ps = [self.conv.weight, self.conv.bias]
        if self.norm is not None:
            ps += [self.norm.weight, self.norm.bias]
        return ps


# ---------------------------------------------------------------------------------------------------
# deferred BatchNorm running-statistics updates: inside ``defer_bn_running_stats(log)`` a training-mode BatchNorm
# block computes its batch statistics without touching running_mean / running_var and appends what the update
# needs to ``log``; ``apply_deferred_bn(log)`` performs the updates later, in log order. GanTrainer uses it to run
# the discriminator's real-sample pass on a second stream while keeping the reference's update order.
# ---------------------------------------------------------------------------------------------------
import contextlib as _contextlib
import threading as _threading
import weakref as _weakref

_bn_defer = _threading.local()


@_contextlib.contextmanager
def defer_bn_running_stats(log: list):
    prev = getattr(_bn_defer, "log", None)
    _bn_defer.log = log
    try:
        yield log
    finally:
        _bn_defer.log = prev


def apply_deferred_bn(log: list):
    for nm, mean, rstd, c, count, momentum in log:
        ops.bn_running_update(mean, rstd, c, count, nm.eps, momentum, nm.running_mean, nm.running_var)
        if nm.num_batches_tracked is not None:
            nm.num_batches_tracked.add_(1)
    log.clear()


@_contextlib.contextmanager
def weights_unchanged(module: nn.Module):
    """Caller's promise that the parameters of ``module`` (a Discriminator) are not written inside the scope: the
    passes inside it share ONE bf16 operand copy per direction instead of re-packing per pass. ``GanTrainer.step``
    wraps the part of a training step in which the discriminator is constant by construction -- the G phase's pass,
    the D phase's two passes and their backwards all precede ``opt_d.step()`` (ref:src/model.py:259-281) -- which
    takes 4 of the 6 discriminator packs out of a step. Leaving the scope drops the copies; outside it every pass
    packs from the live parameters as before."""
    caches = [m._cache for m in module.modules() if isinstance(getattr(m, "_cache", None), _PackedWeights)]
    for c in caches:
        c.hold = True
    try:
        yield module
    finally:
        for c in caches:
            c.hold = False
            c._cur.clear()


def prepack_weights(module: nn.Module):
    """Pack the forward bf16 weight operands of a Discriminator / Generator now, on the current stream. The next
    passes re-pack anyway; this exists for callers that want the first pack ordered on a particular stream."""
    if isinstance(module, Discriminator):
        chain = module._net()
        chain.cache.refresh(chain.blocks, 0)
    elif isinstance(module, Generator):
        net = module._net()
        net.cache.refresh(net.fwd_blocks(), 0)


class _Saved:
    __slots__ = ("src0", "src1", "y", "a", "mean", "rstd", "scale", "shift", "seed", "mode", "in_dhw", "drop_p")


def _bn_momentum(nm) -> float:
    """Exponential-average factor of a BatchNorm call in training mode (``momentum=None``: cumulative average)."""
    if nm.momentum is not None:
        return float(nm.momentum)
    seen = int(nm.num_batches_tracked.item()) if nm.num_batches_tracked is not None else 0
    return 1.0 / float(seen + 1)


class _Act:
    """A block's materialised activations: the bf16 tensor ``a`` (what a weight gradient, a max-pool backward or a
    generic-kernel conv reads; None when no consumer of this pass needs it) and, for blocks that feed a marching
    forward conv, ``a16``: the same fp32 values rounded to fp16, multiplied against fp16-packed weight columns."""
    __slots__ = ("a", "a16")

    def __init__(self, a, a16):
        self.a, self.a16 = a, a16


def _block_forward(blk: _Block, cache: _PackedWeights, src0, src1, seed, pool=False, save=True, defer=False,
                   defer_f16=False, f16_copy=False, keep_bf16=True, src0_saved=None):
    """-> (a, pooled, saved). ``src0`` may be an ``ops.DeferredAct``. ``defer``: do not materialise this block's
    activations -- ``a`` is returned as an ``ops.DeferredAct`` for consumers that apply it on their operand path
    (with ``pool`` only the pooled tensor is written). ``f16_copy``: return an ``_Act`` that also carries the fp16
    operand copy (``keep_bf16=False``: without the bf16 tensor). ``src0_saved``: what the weight gradient of this
    block should read as source 0 when the forward consumed the fp16 copy. Train / eval behaviour follows the
    block's own modules: the norm layer's ``training`` flag (batch vs running statistics) and the dropout layer's
    ``p`` / ``training``."""
    spec = blk.spec
    w = cache.get(spec, blk.conv.weight, 0)
    if blk.s2d_convert and src0.dim() == 5:
        src0 = ops.to_s2d(src0)        # the saved source (weight gradient) is the space-to-depth copy
    n, d, h, wd = spec.in_dims(src0)
    od, oh, ow = spec.out_dims(d, h, wd)
    sv = _Saved() if save else None
    if blk.norm is None:
        act = 1 if blk.fused_act else 0
        a, _ = ops.conv_fwd(spec, src0, src1, w, blk.conv.bias, act=act, slope=blk.slope)
        if save:
            sv.src0, sv.src1, sv.y, sv.a, sv.mode, sv.in_dhw = src0_saved if src0_saved is not None else src0, src1, None, a, \
                UB_NORM_NONE, (d, h, wd)
            sv.mean = sv.rstd = sv.scale = sv.shift = None
            sv.seed, sv.drop_p = 0, 0.0
        return a, None, sv
    y, stats = ops.conv_fwd(spec, src0, src1, w, blk.conv.bias, want_stats=True)
    nm = blk.norm
    momentum = 0.1
    if blk.norm_kind == "instance":
        mode = UB_NORM_INSTANCE
        rm = rv = None
    else:
        mode = UB_NORM_BATCH_TRAIN if (nm.training or not nm.track_running_stats) else UB_NORM_BATCH_EVAL
        rm, rv = nm.running_mean, nm.running_var
        if mode == UB_NORM_BATCH_TRAIN and rm is not None:
            momentum = _bn_momentum(nm)
    defer_log = getattr(_bn_defer, "log", None) if (mode == UB_NORM_BATCH_TRAIN and rm is not None) else None
    if defer_log is not None:
        rm = rv = None
    elif mode == UB_NORM_BATCH_TRAIN and nm.num_batches_tracked is not None:
        nm.num_batches_tracked.add_(1)
    scale, shift, mean, rstd = ops.norm_finalize(stats, n, od * oh * ow, spec.cop, spec.co, nm.weight, nm.bias,
                                                 nm.eps, mode, momentum, rm, rv)
    if defer_log is not None:
        defer_log.append((nm, mean, rstd, spec.co, float(n) * od * oh * ow, momentum))
    drop_p = blk.dropout_p()
    if defer:
        pooled = None
        if pool:
            _, pooled = ops.norm_act_fwd(y, scale, shift, blk.slope, drop_p, seed, pool=True, materialize=False)
        a = ops.DeferredAct(y, scale, shift, blk.slope, drop_p, seed, f16_operand=defer_f16)
    elif f16_copy:
        a_bf, pooled, a16 = ops.norm_act_fwd(y, scale, shift, blk.slope, drop_p, seed, pool=pool, materialize=keep_bf16,
                                             f16_copy=True)
        a = _Act(a_bf, a16)
    else:
        a, pooled = ops.norm_act_fwd(y, scale, shift, blk.slope, drop_p, seed, pool=pool)
    if save:
        sv.src0, sv.src1, sv.y, sv.mode, sv.in_dhw = src0_saved if src0_saved is not None else src0, src1, y, mode, (d, h, wd)
        sv.a = a.a if isinstance(a, _Act) else a
        sv.mean, sv.rstd, sv.scale, sv.shift, sv.seed, sv.drop_p = mean, rstd, scale, shift, seed, drop_p
    return a, pooled, sv


# ---------------------------------------------------------------------------------------------------
# weight gradients on a side stream. In the backward pass the chain  norm/act backward -> dgrad -> norm/act
# backward -> ...  is strictly sequential, while the weight gradient of a block only needs that block's dy.
# The wgrad kernels are tensor / shared-memory bound and leave HBM idle; the norm/activation backward
# passes are HBM bound and use neither shared memory nor TMEM, so their thread blocks co-reside with the
# wgrad CTAs on the same SMs. (wgrad and dgrad cannot co-reside -- both want the whole TMEM -- and simply
# interleave.) UB_WGRAD_STREAM=0 keeps everything on one stream.
# ---------------------------------------------------------------------------------------------------
import os as _os

_WGRAD_SIDE = _os.environ.get("UB_WGRAD_STREAM", "1") != "0"
_side_streams = {}


def _side_stream(device):
    st = _side_streams.get(device)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _side_streams[device] = st
    return st


_side_keep = {}     # device -> tensors the side stream may still be reading: held until the join


def _wgrad_async(spec, src0, src1, dy, weight_shape, ready=None):
    """conv_wgrad on the side stream, ordered after ``ready`` (an event recorded once dy was complete; default:
    everything enqueued so far on the current stream). The caller joins with ``_join_side_stream`` before
    the gradients leave the backward function.

    Memory: the operands stay referenced (``_side_keep``) until the join, after which the main stream is ordered
    behind every side-stream kernel and freeing them is safe -- no ``Tensor.record_stream``. (With record_stream the
    caching allocator could not reuse a freed operand until it had polled the side stream's events, answered with
    fresh cudaMallocs instead, and kept growing for ~20 steps: 80 GiB reserved for 26 GiB of live tensors, with
    synchronising device allocations inside timed steps -- tools/step_jitter.py, profiles/r02f_step_jitter.txt.)
    The output and the workspace are allocated by ``ops.conv_wgrad`` while the side stream is current, i.e. from the
    side stream's own pool, and ``dw`` is only read after the join."""
    if not _WGRAD_SIDE:
        return ops.conv_wgrad(spec, src0, src1, dy, weight_shape)
    main = torch.cuda.current_stream()
    side = _side_stream(dy.device)
    if ready is not None:
        side.wait_event(ready)
    else:
        side.wait_stream(main)
    with torch.cuda.stream(side):
        dw = ops.conv_wgrad(spec, src0, src1, dy, weight_shape)
    _side_keep.setdefault(dy.device, []).append((src0, src1, dy))
    return dw


def _join_side_stream(device):
    if _WGRAD_SIDE and device in _side_streams:
        torch.cuda.current_stream().wait_stream(_side_streams[device])
        _side_keep.pop(device, None)      # ordered behind the side stream now: the operands may be recycled


def _fusion_of(blk: _Block, sv: _Saved, any_width=False):
    """Norm-backward fusion descriptor of a block whose dA a downstream dgrad (32-channel blocks: the marching
    kernel's epilogue) or, with ``any_width``, the max-pool backward is about to produce."""
    if blk.norm is None or sv is None or sv.y is None or (sv.y.shape[-1] != 32 and not any_width):
        return None
    return ops.NormBwdFusion(sv.y, sv.scale, sv.shift, sv.mean, sv.rstd, blk.slope, sv.drop_p, sv.seed)


# the max-pool backward of an encoder level completes dA of the level's second conv block and CAN accumulate that
# block's norm-backward reductions on the way (ub_maxpool_bwd_fused). Measured on B200 at 8 x 128^3 x 32 it does not
# pay: 1.10 ms fused against 0.54 + 0.49 ms for the two passes (the fused kernel is issue- and latency-bound at 128
# registers, profiles/r02d_pool_fused_ncu_summary.txt): opt-in with UB_POOL_FUSE=1
_POOL_FUSE = _os.environ.get("UB_POOL_FUSE", "0") == "1"
# the PatchGAN body d2 .. d5 reads space-to-depth copies of its inputs (UB_D_S2D=0: strided TMA on the plain tensors)
_D_S2D = _os.environ.get("UB_D_S2D", "1") != "0"
# the output head's backward runs fused with the norm backward of the (deferred) block in front of it
_HEAD_FUSE = _os.environ.get("UB_HEAD_FUSE", "1") != "0"


def _block_backward(blk: _Block, cache: _PackedWeights, sv: _Saved, dA, need_w, need_in, grads, partial=None,
                    producer=None, dy_pre=None):
    """Accumulates parameter grads into ``grads`` (dict id(param) -> tensor); returns
    (d_src0, d_src1, partial_of_producer). ``partial``: this block's norm-backward reductions if the
    dgrad that produced ``dA`` already accumulated them. ``producer`` = (block, saved) of the block
    whose activations are this conv's src0: when the dgrad runs on the marching kernel its epilogue
    accumulates the producer's reductions. ``dy_pre`` = (dy, dgamma, dbeta, dbias): the norm / activation backward of
    this block already ran (fused into the output head's backward); ``dA`` is not used."""
    spec = blk.spec
    co = spec.co
    if dy_pre is not None:
        dy, dgamma, dbeta, dbias = dy_pre
        if need_w:
            grads[id(blk.norm.weight)] = dgamma
            grads[id(blk.norm.bias)] = dbeta
            grads[id(blk.conv.bias)] = dbias
    elif blk.norm is not None:
        dy, dgamma, dbeta, dbias = ops.norm_act_bwd(dA, None, sv.y, sv.mode, sv.mean, sv.rstd, sv.scale, blk.slope,
                                                    sv.drop_p, sv.seed, co, want_param_grads=need_w,
                                                    want_bias_grad=need_w, shift=sv.shift, partial=partial)
        if need_w:
            grads[id(blk.norm.weight)] = dgamma
            grads[id(blk.norm.bias)] = dbeta
            grads[id(blk.conv.bias)] = dbias
    else:
        if blk.fused_act:
            dy, _, _, _ = ops.norm_act_bwd(dA, sv.a, None, UB_NORM_NONE, None, None, None, blk.slope, 0.0, 0, co)
        else:
            dy = dA
        if need_w:
            grads[id(blk.conv.bias)] = ops.colsum(dy, co)
    # The dgrad (the critical chain) is enqueued FIRST; the wgrad goes to the side stream afterwards, ordered
    # only after dy: it then fills the SMs while the next block's memory-bound norm/activation backward runs.
    dy_ready = None
    if need_w and need_in and _WGRAD_SIDE:
        dy_ready = torch.cuda.Event()
        dy_ready.record()
    result = (None, None, None)
    if need_in:
        wd = cache.get(spec, blk.conv.weight, 1)
        fuse = _fusion_of(*producer) if producer is not None else None
        if fuse is not None and ops.dgrad_fuse_records(spec, dy.shape[0], *sv.in_dhw) > 0:
            result = ops.conv_dgrad(spec, dy, wd, sv.in_dhw, fuse=fuse)
        else:
            d0, d1 = ops.conv_dgrad(spec, dy, wd, sv.in_dhw)
            result = (d0, d1, None)
    if need_w:
        grads[id(blk.conv.weight)] = _wgrad_async(spec, sv.src0, sv.src1, dy, tuple(blk.conv.weight.shape), ready=dy_ready)
    return result


class _InputPackCache:
    """bf16 NDHWC pack of the most recent module input. ``training_step`` feeds the same ``x`` to the
    generator twice per step (ref:src/model.py:171,184); the second call reuses the pack. The entry is keyed on
    the identity of the tensor object (held WEAKLY: the cache never extends the life of a batch), its in-place
    version counter and its data pointer; it is dropped as soon as the tensor dies or another input arrives."""

    def __init__(self):
        self._ref = self._key = self._packed = None

    def get(self, x):
        key = (x.data_ptr(), x._version, tuple(x.shape), x.dtype, x.device)
        if self._ref is not None and self._ref() is x and self._key == key:
            return self._packed
        packed = ops.pack_ncdhw(x)
        self._key, self._packed = key, packed
        self._ref = _weakref.ref(x, self._drop)
        return packed

    def _drop(self, _):
        self._ref = self._key = self._packed = None

    def clear(self):
        self._drop(None)


def _out_dtype(x):
    """Module outputs follow the input dtype, except for a bf16 input: that is a TRANSPORT format of the fp32 pipeline
    (it halves the host-to-device bytes of the conditioning input and packs to the same bits), so the result stays
    fp32 and the whole step is bit-identical to feeding the fp32 tensor."""
    return torch.float32 if x.dtype == torch.bfloat16 else x.dtype


def _fresh_seed() -> int:
    # host RNG (follows torch.manual_seed), no device sync
    return int(torch.empty((), dtype=torch.int64).random_().item()) & 0x7FFFFFFF


# ---------------------------------------------------------------------------------------------------
# parameter containers with the reference's names
# ---------------------------------------------------------------------------------------------------
class DownSampleConv(nn.Module):
    """ref:src/model.py:42-65. Supported kernels on the sm_100a path: (kernel=1, strides=1, padding=0)
    and (kernel=4, strides=2, padding=1); anything else raises (there is no fallback)."""

    def __init__(self, in_channels, out_channels, kernel=4, strides=2, padding=1, activation=True, batchnorm=True):
        super().__init__()
        self.activation = activation
        self.batchnorm = batchnorm
        self.conv = nn.Conv3d(in_channels, out_channels, kernel, strides, padding)
        if batchnorm:
            self.bn = nn.BatchNorm3d(out_channels)
        if activation:
            self.act = nn.LeakyReLU(0.2)
        if (kernel, strides, padding) == (1, 1, 0):
            kind = UB_CONV_K1
        elif (kernel, strides, padding) == (4, 2, 1):
            kind = UB_CONV_K4S2P1
        else:
            raise NotImplementedError(
                f"DownSampleConv(kernel={kernel}, strides={strides}, padding={padding}) has no sm_100a kernel")
        self._cache = _PackedWeights()
        self._kind = kind

    def _block(self, name="dsc", s2d_input=False, s2d_convert=False):
        """``s2d_input``: the block is the first of a chain and reads the space-to-depth pack of the
        module input (stride-2 stem of the PatchGAN). ``s2d_convert``: a stride-2 block inside a chain reads a
        space-to-depth COPY of its input (made in the forward, ``ops.to_s2d``) with unstrided TMA boxes instead of
        eight parity tiles with element stride 2."""
        # copies pay where the output tile is one N tile (depth shifts folded into N = 2 * co <= 256): measured at the
        # bench shapes d2 forward 0.225 -> 0.086 ms, d3 0.075 -> 0.057, but d4 0.072 -> 0.115 and d5 0.129 -> 0.213
        # (profiles/r02g_patchgan_s2d.txt), so the wide layers keep the strided boxes
        s2d_convert = s2d_convert and self.conv.out_channels <= 128
        s2d = (s2d_input or s2d_convert) and self._kind == UB_CONV_K4S2P1
        kind = UB_CONV_K4S2P1_S2D if s2d else self._kind
        spec = ops.ConvSpec(kind, self.conv.in_channels, self.conv.out_channels)
        slope = 0.2 if self.activation else 1.0
        conv = s2d_convert and s2d
        if self.batchnorm:
            return _Block(name, spec, self.conv, self.bn, "batch", slope=slope, s2d_convert=conv)
        return _Block(name, spec, self.conv, None, None, slope=slope, fused_act=self.activation, s2d_convert=conv)

    def forward(self, x):
        blk = self._block()
        if _precision_of(self) == "fp32":
            from . import fp32_mode
            return fp32_mode.chain_forward(_Chain([blk], self._cache, self, blk.spec.co), x)
        params = blk.params()
        return _ChainFunction.apply(x, None, _Chain([blk], self._cache, self, blk.spec.co), torch.is_grad_enabled(),
                                    *params)


class _ADN(nn.Sequential):
    def __init__(self, channels, dropout):
        super().__init__()
        self.add_module("N", nn.InstanceNorm3d(channels, affine=True))
        if dropout is not None and dropout > 0:
            self.add_module("D", nn.Dropout(dropout))
        self.add_module("A", nn.LeakyReLU(negative_slope=0.1, inplace=True))


class _Convolution(nn.Sequential):
    def __init__(self, in_chns, out_chns, dropout):
        super().__init__()
        self.add_module("conv", nn.Conv3d(in_chns, out_chns, kernel_size=3, stride=1, padding=1, bias=True))
        self.add_module("adn", _ADN(out_chns, dropout))


class _TwoConv(nn.Sequential):
    def __init__(self, in_chns, out_chns, dropout):
        super().__init__()
        self.add_module("conv_0", _Convolution(in_chns, out_chns, dropout))
        self.add_module("conv_1", _Convolution(out_chns, out_chns, dropout))


class _Down(nn.Sequential):
    def __init__(self, in_chns, out_chns, dropout):
        super().__init__()
        self.add_module("max_pooling", nn.MaxPool3d(kernel_size=2))
        self.add_module("convs", _TwoConv(in_chns, out_chns, dropout))


class _UpSample(nn.Sequential):
    def __init__(self, in_chns, out_chns):
        super().__init__()
        self.add_module("deconv", nn.ConvTranspose3d(in_chns, out_chns, kernel_size=2, stride=2, bias=True))


class _UpCat(nn.Module):
    def __init__(self, in_chns, cat_chns, out_chns, dropout, halves=True):
        super().__init__()
        up_chns = in_chns // 2 if halves else in_chns
        self.upsample = _UpSample(in_chns, up_chns)
        self.convs = _TwoConv(cat_chns + up_chns, out_chns, dropout)


class BasicUNet(nn.Module):
    """Parameter tree of monai==1.3.0 ``BasicUNet(spatial_dims=3, ...)`` (keys of SURVEY.md Appendix A).
    ``forward`` runs the sm_100a engine (same path ``Generator`` uses, without an input head)."""

    def __init__(self, spatial_dims=3, in_channels=24, out_channels=6, features=UNET_FEATURES, dropout=0.05):
        super().__init__()
        if spatial_dims != 3:
            raise NotImplementedError("only spatial_dims=3 is on the hot path")
        f = tuple(features)
        self.in_channels, self.out_channels, self.features, self.dropout = in_channels, out_channels, f, dropout
        self.conv_0 = _TwoConv(in_channels, f[0], dropout)
        self.down_1 = _Down(f[0], f[1], dropout)
        self.down_2 = _Down(f[1], f[2], dropout)
        self.down_3 = _Down(f[2], f[3], dropout)
        self.down_4 = _Down(f[3], f[4], dropout)
        self.upcat_4 = _UpCat(f[4], f[3], f[3], dropout)
        self.upcat_3 = _UpCat(f[3], f[2], f[2], dropout)
        self.upcat_2 = _UpCat(f[2], f[1], f[1], dropout)
        self.upcat_1 = _UpCat(f[1], f[0], f[5], dropout, halves=False)
        self.final_conv = nn.Conv3d(f[5], out_channels, kernel_size=1)
        self._cache = _PackedWeights()

    def forward(self, x):
        net = _UNetGraph(None, self, self._cache)
        if _precision_of(self) == "fp32":
            from . import fp32_mode
            _check_unet_dims(*x.shape[2:])
            return fp32_mode.generator_forward(net, x, _fresh_seed)
        return _GeneratorFunction.apply(x, net, torch.is_grad_enabled(), *net.params)


# ---------------------------------------------------------------------------------------------------
# generic chain of blocks (DownSampleConv standalone, Discriminator)
# ---------------------------------------------------------------------------------------------------
class _Chain:
    def __init__(self, blocks, cache, owner, out_channels):
        self.blocks = blocks
        self.cache = cache
        self.owner = owner
        self.out_channels = out_channels
        self.params = []
        seen = set()
        for b in blocks:
            for p in b.params():
                if id(p) not in seen:
                    seen.add(id(p))
                    self.params.append(p)


class _ChainFunction(torch.autograd.Function):
    """x (and optionally y, concatenated on channels) -> blocks... -> NCDHW fp32."""

    @staticmethod
    def forward(ctx, x, y, chain: _Chain, grad_enabled, *params):
        need_bwd = grad_enabled and (x.requires_grad or (y is not None and y.requires_grad) or
                                     any(p.requires_grad for p in params))
        chain.cache.refresh(chain.blocks, 0)
        a = ops.pack_ncdhw(x, y, s2d=chain.blocks[0].spec.kind == UB_CONV_K4S2P1_S2D)
        saved = []
        for blk in chain.blocks:
            a, _, sv = _block_forward(blk, chain.cache, a, None, 0, save=need_bwd)
            saved.append(sv)
        out = ops.unpack_ncdhw(a, chain.out_channels)
        ctx.chain, ctx.saved = chain, saved
        ctx.cx = x.shape[1]
        ctx.cy = y.shape[1] if y is not None else 0
        ctx.in_dtype = x.dtype
        ctx.y_dtype = y.dtype if y is not None else None
        return out.to(_out_dtype(x))

    @staticmethod
    def backward(ctx, dout):
        chain, saved = ctx.chain, ctx.saved
        if saved is None or any(sv is None for sv in saved):
            raise RuntimeError("backward called a second time (or without saved activations): the intermediate "
                               "activations are freed as the first backward consumes them; run the forward again")
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        pneed = {id(p): ctx.needs_input_grad[4 + i] for i, p in enumerate(chain.params)}
        grads = {}
        dA = ops.pack_ncdhw(dout.contiguous().float())
        nblk = len(chain.blocks)
        chain.cache.refresh(chain.blocks[1:] if not (need_x or need_y) else chain.blocks, 1)
        d0 = None
        for i in range(nblk - 1, -1, -1):
            blk = chain.blocks[i]
            need_w = any(pneed[id(p)] for p in blk.params())
            need_in = i > 0 or need_x or need_y
            d0, _, _ = _block_backward(blk, chain.cache, saved[i], dA, need_w, need_in, grads)
            saved[i] = None
            dA = d0
        dx = ops.unpack_ncdhw(d0, ctx.cx, 0).to(ctx.in_dtype) if need_x else None
        dy = ops.unpack_ncdhw(d0, ctx.cy, ctx.cx).to(ctx.y_dtype) if need_y else None
        _join_side_stream(dout.device)
        ctx.saved = None
        pg = [grads.get(id(p)) if pneed[id(p)] else None for p in chain.params]
        return (dx, dy, None, None, *pg)


# ---------------------------------------------------------------------------------------------------
# Generator graph: head -> U-Net with skip concat folded into the consumer conv
# ---------------------------------------------------------------------------------------------------
class _UNetGraph:
    def __init__(self, head: DownSampleConv | None, unet: BasicUNet, cache: _PackedWeights):
        self.head_mod = head
        self.unet = unet
        self.cache = cache
        self.input_pack = _InputPackCache()
        f = unet.features
        K3 = UB_CONV_K3S1P1

        def cb(name, mod, cin, cout, c1=0):
            # the block's own dropout layer (monai ADN "D", absent when dropout == 0) and activation slope are
            # read at run time, so ``nn.Dropout.p`` / ``train()`` / ``eval()`` changes on the modules take effect
            return _Block(name, ops.ConvSpec(K3, cin, cout, c1), mod.conv, mod.adn.N, "instance",
                          slope=float(mod.adn.A.negative_slope), drop_mod=getattr(mod.adn, "D", None))

        self.head = head._block("head") if head is not None else None
        self.enc = [(cb("conv_0.conv_0", unet.conv_0.conv_0, unet.in_channels, f[0]),
                     cb("conv_0.conv_1", unet.conv_0.conv_1, f[0], f[0]))]
        for k, dn in enumerate((unet.down_1, unet.down_2, unet.down_3, unet.down_4), start=1):
            self.enc.append((cb(f"down_{k}.conv_0", dn.convs.conv_0, f[k - 1], f[k]),
                             cb(f"down_{k}.conv_1", dn.convs.conv_1, f[k], f[k])))
        self.dec = []  # upcat_4 .. upcat_1
        for k, up in zip((4, 3, 2, 1), (unet.upcat_4, unet.upcat_3, unet.upcat_2, unet.upcat_1)):
            dc = up.upsample.deconv
            cin, cup = dc.in_channels, dc.out_channels
            cat = f[k - 1]
            cout = f[5] if k == 1 else f[k - 1]
            self.dec.append((
                _Block(f"upcat_{k}.deconv", ops.ConvSpec(UB_DECONV_K2S2, cin, cup), dc),
                cb(f"upcat_{k}.conv_0", up.convs.conv_0, cat, cout, c1=cup),
                cb(f"upcat_{k}.conv_1", up.convs.conv_1, cout, cout),
            ))
        self.final = _Block("final_conv", ops.ConvSpec(UB_CONV_K1, f[5], unet.out_channels), unet.final_conv)
        self._defer_plans = {}
        self.blocks = ([self.head] if self.head else []) + [b for pair in self.enc for b in pair] + \
                      [b for tri in self.dec for b in tri] + [self.final]
        self.params = []
        seen = set()
        for b in self.blocks:
            for q in b.params():
                if id(q) not in seen:
                    seen.add(id(q))
                    self.params.append(q)

    @property
    def training(self):
        return self.unet.training

    def fwd_blocks(self):
        """Blocks whose forward runs on a conv kernel (the output head runs fused with the layout change)."""
        return [b for b in self.blocks if not (b is self.final and _final_is_fusable(self))]

    def defer_plan(self, n, d, h, w, ncdhw_out, need_bwd):
        """-> (deferred, f16_consumers, f16_producers): names of
          * the blocks whose activations stay DEFERRED (never materialised): every consumer applies the norm + dropout +
            LeakyReLU on its own operand path. By default that is the block in front of the fused output head (a
            memory-bound consumer: in registers). With UB_DEFER_CONV=1, in passes without a backward, also the blocks
            in front of 3x3x3 convs on the marching kernel (in shared memory, between the TMA arrival and the MMAs;
            measured on B200 NOT to pay on these shared-memory-bound kernels, profiles/r02_deferred_microbench.txt);
          * the conv blocks that read source 0 as an fp16 operand against fp16-packed weight columns, and
          * the blocks that therefore write an fp16 copy of their activations beside (or, without a backward to
            come, instead of) the bf16 tensor. These are the full-resolution 32-channel layers: they carry the skip
            path straight to the output, and their bf16 operand rounding was the largest single term of the
            end-to-end error (generator output rel-L2 1.25e-2 -> 0.86e-2 at 8 x 128^3).
        UB_DEFER=0 / UB_FWD_F16=0 disable the two mechanisms."""
        key = (n, d, h, w, ncdhw_out, need_bwd)
        hit = self._defer_plans.get(key)
        if hit is not None:
            return hit
        plan, f16c, f16p = set(), set(), set()
        nlev = len(self.enc)

        def ok(consumer, lvl):
            return ops.deferred_src0_ok(consumer.spec, n, d >> lvl, h >> lvl, w >> lvl)

        edges = []          # (producer block, consumer conv block, level) along which source 0 travels
        if self.head is not None and self.head.spec.cop == 32:
            edges.append((self.head, self.enc[0][0], 0))
        for lvl, (c0, c1) in enumerate(self.enc):
            edges.append((c0, c1, lvl))
            if lvl < nlev - 1:
                edges.append((c1, self.dec[nlev - 2 - lvl][1], lvl))      # the skip conv (+ the pooling pass)
        for j, (dc, c0, c1) in enumerate(self.dec):
            edges.append((c0, c1, nlev - 2 - j))
        edges = [(p, c, lvl) for p, c, lvl in edges if ok(c, lvl)]
        if _DEFER:
            last_c1 = self.dec[-1][2]
            if ncdhw_out and _final_is_fusable(self) and last_c1.spec.cop == 32:
                plan.add(last_c1.name)                         # consumer: the fused output head
            if not need_bwd and _DEFER_CONV:
                for p, c, _ in edges:
                    plan.add(p.name); f16c.add(c.name)
        if _FWD_F16:
            for p, c, _ in edges:
                if p.name not in plan:
                    f16p.add(p.name); f16c.add(c.name)
        self._defer_plans[key] = (frozenset(plan), frozenset(f16c), frozenset(f16p))
        return self._defer_plans[key]


_DEFER = _os_env.environ.get("UB_DEFER", "1") != "0"
# the conv operand path of deferred activations (deferred_tile.cuh): correct and tested, but on these shared-memory-
# bound kernels it does not beat one materialising pass at HBM speed (profiles/r02_deferred_microbench.txt): opt-in
_DEFER_CONV = _os_env.environ.get("UB_DEFER_CONV", "0") == "1"
# fp16 operand copies of the activations in front of the marching forward convs (see _UNetGraph.defer_plan)
_FWD_F16 = _os_env.environ.get("UB_FWD_F16", "1") != "0"


def _check_unet_dims(d, h, w):
    if d % 16 or h % 16 or w % 16:
        raise RuntimeError(f"U-Net input size ({d},{h},{w}) must be divisible by 16 (the replicate-pad branch of "
                           "monai UpCat for odd sizes is not on the sm_100a path)")


def _final_is_fusable(net: "_UNetGraph") -> bool:
    sp = net.final.spec
    return sp.c0p == 32 and sp.c0 <= 32 and sp.co <= 8


def _generator_run(net: _UNetGraph, a, need_bwd: bool, ncdhw_out: bool = False):
    """Packed input (N,D,H,W,32) bf16 -> output of ``final_conv`` and the saved-for-backward dict (or
    None). ``ncdhw_out``: the output head runs fused with the layout change and returns (N,6,D,H,W) fp32
    (module boundary); otherwise the packed (N,D,H,W,32) bf16 tensor (inference keeps it packed). The one
    place that sequences the generator's kernels."""
    any_dropout = any(b.dropout_p() > 0.0 for b in net.blocks)
    base_seed = _fresh_seed() if any_dropout else 0
    cache = net.cache
    n, d, h, w, _ = a.shape
    deferred, f16_consumers, f16_producers = net.defer_plan(n, d, h, w, ncdhw_out, need_bwd)
    cache.refresh(net.fwd_blocks(), 0, f16_src0=f16_consumers)
    S = {}          # block name -> saved
    lid = [0]

    def run(blk, s0, s1=None, pool=False, bf16_consumer=False):
        """``s0`` may be an ``_Act``: a conv listed in ``f16_consumers`` reads its fp16 copy, everything else (and
        the weight gradient later) its bf16 tensor. ``bf16_consumer``: somebody downstream in THIS pass reads the
        block's bf16 tensor (a generic-kernel conv), so it is written even without a backward."""
        lid[0] += 1
        fwd_src, saved_src = s0, None
        if isinstance(s0, _Act):
            fwd_src = s0.a16 if blk.name in f16_consumers else s0.a
            saved_src = s0.a
            if fwd_src is None:
                raise RuntimeError(f"{blk.name}: the producer's bf16 activations were not materialised in this pass")
        a_, pooled, sv = _block_forward(blk, cache, fwd_src, s1, (base_seed + 7919 * lid[0]) & 0x7FFFFFFF,
                                        pool=pool, save=need_bwd, defer=blk.name in deferred,
                                        defer_f16=bool(f16_consumers), f16_copy=blk.name in f16_producers,
                                        keep_bf16=need_bwd or bf16_consumer, src0_saved=saved_src)
        S[blk.name] = sv
        return a_, pooled

    if net.head is not None:
        a, _ = run(net.head, a)
    skips = []
    cur = a
    for lvl, (c0, c1) in enumerate(net.enc):
        t, _ = run(c0, cur)
        last = lvl == len(net.enc) - 1
        xk, pooled = run(c1, t, pool=not last)
        skips.append(xk)
        cur = xk if last else pooled
    u = skips[-1]
    for j, (dc, c0, c1) in enumerate(net.dec):
        x_e = skips[-2 - j]
        up, _ = run(dc, u)
        t, _ = run(c0, x_e, up)
        u, _ = run(c1, t)
    if ncdhw_out and _final_is_fusable(net):
        out = ops.conv1x1_to_ncdhw(u, net.final.conv.weight, net.final.conv.bias)
        if need_bwd:
            sv = _Saved()
            sv.src0 = u
            S[net.final.name] = sv
        return out, (S if need_bwd else None)
    yf, _ = run(net.final, u)
    if ncdhw_out:
        yf = ops.unpack_ncdhw(yf, net.unet.out_channels)
    return yf, (S if need_bwd else None)


class _GeneratorFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net: _UNetGraph, grad_enabled, *params):
        need_bwd = grad_enabled and (x.requires_grad or any(p.requires_grad for p in params))
        _check_unet_dims(*x.shape[2:])
        a = net.input_pack.get(x)
        out, S = _generator_run(net, a, need_bwd, ncdhw_out=True)
        ctx.net, ctx.S = net, S
        ctx.cx = x.shape[1]
        ctx.in_dtype = x.dtype
        return out.to(_out_dtype(x))

    @staticmethod
    def backward(ctx, dout):
        net, S = ctx.net, ctx.S
        if S is None:
            raise RuntimeError("Generator backward called a second time (or without saved activations): the "
                               "intermediate activations are freed as the first backward consumes them; run the "
                               "forward again")
        cache = net.cache
        pneed = {id(p): ctx.needs_input_grad[3 + i] for i, p in enumerate(net.params)}
        need_x = ctx.needs_input_grad[0]
        grads = {}
        first_blk = net.head if net.head is not None else net.enc[0][0]
        cache.refresh([b for b in net.fwd_blocks() if need_x or b is not first_blk], 1)

        def bwd(blk, dA, need_in=True, partial=None, producer=None, dy_pre=None):
            """-> (d_src0, d_src1, partial of ``producer``)."""
            need_w = any(pneed[id(p)] for p in blk.params())
            prod = (producer, S[producer.name]) if producer is not None else None
            r = _block_backward(blk, cache, S[blk.name], dA, need_w, need_in, grads, partial=partial, producer=prod,
                                dy_pre=dy_pre)
            S[blk.name] = None
            return r

        head_dy = None      # (dy, dgamma, dbeta, dbias) of upcat_1.conv_1 when the head's backward produced them
        if _final_is_fusable(net):
            # output head: dgrad, wgrad and bias gradient in one pass over dout (NCDHW) and the saved input
            fw, fb = net.final.conv.weight, net.final.conv.bias
            need_w = pneed[id(fw)] or pneed[id(fb)]
            last_c1 = net.dec[-1][2]
            src = S[net.final.name].src0
            fuse = _fusion_of(last_c1, S[last_c1.name]) if (_HEAD_FUSE and isinstance(src, ops.DeferredAct)) else None
            if fuse is not None and ops.head_bwd_fused_ok(dout, fuse):
                # the block in front of the head is deferred: its norm backward runs inside the head's backward
                # (ub_head_bwd_fused) and du is never written
                need_blk = any(pneed[id(p)] for p in last_c1.params())
                dy_, dg_, dbt_, dbs_, dw_, db_ = ops.head_bwd_fused(dout, fuse, fw, S[last_c1.name].mode, last_c1.spec.co,
                                                                    need_params=need_w, want_block_grads=need_blk)
                head_dy = (dy_, dg_, dbt_, dbs_)
                du = None
            else:
                du, dw_, db_ = ops.conv1x1_from_ncdhw_bwd(dout, src, fw, need_input=True, need_params=need_w)
            if need_w:
                grads[id(fw)], grads[id(fb)] = dw_, db_
            S[net.final.name] = None
        else:
            dA = ops.pack_ncdhw(dout.contiguous().float())
            du, _, _ = bwd(net.final, dA)
        nlev = len(net.enc)
        dskip = [None] * nlev
        for j in range(len(net.dec) - 1, -1, -1):      # upcat_1 first
            dc, c0, c1 = net.dec[j]
            dt, _, part = bwd(c1, du, producer=c0, dy_pre=head_dy)      # dt = dA of c0 (same resolution, chained)
            head_dy = None
            d_xe, d_up, _ = bwd(c0, dt, partial=part)
            dskip[nlev - 2 - j] = d_xe
            du, _, _ = bwd(dc, d_up)
        dcur = du                                        # grad of x4 (bottom)
        part_head = None
        for lvl in range(nlev - 1, -1, -1):
            c0, c1 = net.enc[lvl]
            part_c1 = None
            if lvl != nlev - 1:
                # dcur is the grad of the pooled tensor; route it onto the skip grad of x_lvl
                fuse = _fusion_of(c1, S[c1.name], any_width=True) if _POOL_FUSE else None
                if fuse is not None and ops.maxpool_bwd_fuse_records(*fuse.y.shape) > 0:
                    dcur, part_c1 = ops.maxpool_bwd_fused(dcur, dskip[lvl], fuse)
                else:
                    dcur = ops.maxpool_bwd(S[c1.name].a, dcur, dskip[lvl])
            dt, _, part = bwd(c1, dcur, partial=part_c1, producer=c0)
            first = lvl == 0
            dcur, _, part_head = bwd(c0, dt, need_in=(not first) or net.head is not None or need_x, partial=part,
                                     producer=net.head if (first and net.head is not None) else None)
        dx = None
        if net.head is not None:
            dcur, _, _ = bwd(net.head, dcur, need_in=need_x, partial=part_head)
        if need_x:
            dx = ops.unpack_ncdhw(dcur, ctx.cx).to(ctx.in_dtype)
        _join_side_stream(dout.device)
        ctx.S = None
        pg = [grads.get(id(p)) if pneed[id(p)] else None for p in net.params]
        return (dx, None, None, *pg)


class Generator(nn.Module):
    """ref:src/model.py:15-39 -- 1x1x1 input head + BasicUNet(24 -> 6, features 32..512, dropout 0.05)."""

    def __init__(self, input_modality):
        super().__init__()
        self.input_modality = input_modality
        dwi_tensor_input = DownSampleConv(6, 24, kernel=1, strides=1, padding=0)
        bssfp_input = DownSampleConv(24, 24, kernel=1, strides=1, padding=0)
        unet = BasicUNet(spatial_dims=3, in_channels=24, out_channels=6, features=UNET_FEATURES, dropout=0.05)
        self.blocks = nn.ModuleDict({
            "dwi-tensor": dwi_tensor_input,
            "pc-bssfp": bssfp_input,
            "bssfp": bssfp_input,
            "t1w": dwi_tensor_input,
            "unet": unet,
        })
        self._cache = _PackedWeights()
        self._graph = None
        self._graph_key = None

    def _net(self):
        # the graph only wires modules together (hyper-parameters are read from them per call); it is rebuilt when
        # the input modality -- i.e. which head module is wired in -- changes
        head = self.blocks[self.input_modality]
        key = (self.input_modality, id(head), id(self.blocks["unet"]))
        if self._graph is None or self._graph_key != key:
            self._graph = _UNetGraph(head, self.blocks["unet"], self._cache)
            self._graph_key = key
        return self._graph

    def forward(self, x):
        net = self._net()
        if _precision_of(self) == "fp32":
            from . import fp32_mode
            _check_unet_dims(*x.shape[2:])
            return fp32_mode.generator_forward(net, x, _fresh_seed)
        return _GeneratorFunction.apply(x, net, torch.is_grad_enabled(), *net.params)

    @torch.no_grad()
    def forward_packed(self, a):
        """Inference on an already packed batch: (N,D,H,W,32) bf16 NDHWC -> (N,6,D,H,W) fp32, the same
        arithmetic as ``forward`` without packing the input (used by ``inference.predict_volume``)."""
        _check_unet_dims(*a.shape[1:4])
        out, _ = _generator_run(self._net(), a, need_bwd=False, ncdhw_out=True)
        return out


class Discriminator(nn.Module):
    """ref:src/model.py:68-92 -- PatchGAN on cat[x, y] (the concat is folded into the layout pack)."""

    def __init__(self, modality):
        super().__init__()
        self.modality = modality
        d1_bssfp = DownSampleConv(30, 32, batchnorm=False)
        d1_dwi = DownSampleConv(12, 32, batchnorm=False)
        self.d1 = self.blocks = nn.ModuleDict({
            "dwi-tensor": d1_dwi,
            "pc-bssfp": d1_bssfp,
            "bssfp": d1_bssfp,
            "t1w": d1_dwi,
        })
        self.d2 = DownSampleConv(32, 64)
        self.d3 = DownSampleConv(64, 128)
        self.d4 = DownSampleConv(128, 256)
        self.d5 = DownSampleConv(256, 512)
        self.final = nn.Conv3d(512, 1, kernel_size=1)
        self._cache = _PackedWeights()
        self._chain = None
        self._chain_key = None

    def _net(self):
        key = (self.modality, id(self.d1[self.modality]))
        if self._chain is None or self._chain_key != key:
            cv = _D_S2D
            blocks = [self.d1[self.modality]._block("d1", s2d_input=True), self.d2._block("d2", s2d_convert=cv),
                      self.d3._block("d3", s2d_convert=cv), self.d4._block("d4", s2d_convert=cv),
                      self.d5._block("d5", s2d_convert=cv),
                      _Block("final", ops.ConvSpec(UB_CONV_K1, 512, 1), self.final)]
            self._chain = _Chain(blocks, self._cache, self, 1)
            self._chain_key = key
        return self._chain

    def forward(self, x, y):
        d, h, w = x.shape[2:]
        if d % 32 or h % 32 or w % 32:
            raise RuntimeError(f"PatchGAN input size ({d},{h},{w}) must be divisible by 32")
        chain = self._net()
        if _precision_of(self) == "fp32":
            from . import fp32_mode
            return fp32_mode.chain_forward(chain, x, y)
        return _ChainFunction.apply(x, y, chain, torch.is_grad_enabled(), *chain.params)


# ---------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------
class _L1Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a32, b32 = a.contiguous().float(), b.contiguous().float()
        ctx.save_for_backward(a32, b32)
        ctx.dt = (a.dtype, b.dtype)
        return ops.l1_fwd(a32, b32).to(a.dtype)

    @staticmethod
    def backward(ctx, g):
        a32, b32 = ctx.saved_tensors
        g32 = g.contiguous().float()
        da = db = None
        if ctx.needs_input_grad[0]:
            da = ops.l1_bwd(a32, b32, g32).to(ctx.dt[0])
        if ctx.needs_input_grad[1]:
            db = ops.l1_bwd(b32, a32, g32).to(ctx.dt[1])
        return da, db


class L1Loss(nn.Module):
    """Drop-in for ``torch.nn.L1Loss()`` (mean reduction), ref:src/model.py:126,136."""

    def forward(self, input, target):
        if input.shape != target.shape:
            raise RuntimeError("L1Loss: shape mismatch")
        return _L1Function.apply(input, target)


class _BCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        x32 = x.contiguous().float()
        t32 = target.contiguous().float()
        loss, dx = ops.bce_logits(x32, t32, want_grad=True)
        ctx.save_for_backward(dx)
        ctx.dt = x.dtype
        return loss.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return ops.scale_by(dx, g.contiguous().float()).to(ctx.dt), None


class BCEWithLogitsLoss(nn.Module):
    """Drop-in for ``torch.nn.BCEWithLogitsLoss()`` (mean reduction), ref:src/model.py:155."""

    def forward(self, input, target):
        if input.shape != target.shape:
            raise RuntimeError("BCEWithLogitsLoss: shape mismatch")
        return _BCEFunction.apply(input, target)
