"""Whole-volume sliding-window inference and its evaluation, on device.

Mirrors what ``predict_step`` / ``test_step`` of the reference do with torchio (ref:src/model.py:291-333,
ref:src/data_module.py:168-183; torchio==0.19.6 ``GridSampler(subject, patch_size)`` with the default
``patch_overlap = 0`` and ``GridAggregator`` in its default ``'crop'`` mode): the volume is covered by a
grid of patches whose last patch per axis is shifted back to end at the border, every patch goes
through the generator, and ``add_batch`` writes each prediction into the output volume in sampler
order -- where patches overlap the later one wins.

Here the sampler and aggregator are index arithmetic inside the layout kernels (``ub_pack_patches`` /
``ub_paste_patch``): no host round trip and no intermediate NCDHW patch tensors.
``relative_error`` then runs the fused error-map + ROI reduction of ref:src/eval.py:154-166,217-258.
"""
from __future__ import annotations

import itertools
import os

import torch

from . import ops

__all__ = ["grid_locations", "shard_indices", "predict_volume", "relative_error"]

_USE_GRAPH = os.environ.get("UB_INFER_GRAPH", "0") == "1"


class _GraphedForward:
    """``gen.forward_packed`` on one batch shape, captured once as a CUDA graph and replayed per batch (one launch
    instead of ~110). Measured on B200 at config 5 (160 x 192 x 160, 27 patches of 64^3): 5.9 ms per volume either
    way -- the eager path is already GPU-bound there -- so it is opt-in, for small patches / slow hosts.
    Valid while nothing the graph baked in has moved: the storage of the parameters and buffers (the captured
    forward re-packs its bf16 weight operands from the live parameter memory on every replay, so weight VALUES may
    change freely) and eval mode."""

    def __init__(self, gen, shape, device):
        self.gen = gen
        self.inp = torch.zeros(shape, dtype=torch.bfloat16, device=device)
        self.fingerprint = self._fingerprint(gen)
        main = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(main)
        with torch.cuda.stream(side):               # eager warm-up: packs weights, sets kernel attributes
            gen.forward_packed(self.inp)
        main.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = gen.forward_packed(self.inp)

    @staticmethod
    def _fingerprint(gen):
        return tuple(t.data_ptr() for t in itertools.chain(gen.parameters(), gen.buffers()))

    def valid(self):
        return (not self.gen.training) and self.fingerprint == self._fingerprint(self.gen)

    def __call__(self):
        self.graph.replay()
        return self.out


def _graphed_forward(gen, shape, device):
    cache = gen.__dict__.setdefault("_infer_graphs", {})
    key = (tuple(shape), str(device))
    g = cache.get(key)
    if g is None or not g.valid():
        g = _GraphedForward(gen, shape, device)
        cache[key] = g
    return g


def grid_locations(shape, patch):
    """Patch origins of torchio's GridSampler for ``patch_overlap = 0``, in sampler order
    (lexicographic in (axis0, axis1, axis2)): per axis ``range(0, size + 1 - patch, patch)`` plus a
    last origin at ``size - patch`` when the grid does not end at the border."""
    per_axis = []
    for size, p in zip(shape, patch):
        if p > size:
            raise ValueError(f"patch size {p} larger than the volume ({size})")
        idx = list(range(0, size + 1 - p, p))
        if idx[-1] != size - p:
            idx.append(size - p)
        per_axis.append(idx)
    return list(itertools.product(*per_axis))


def shard_indices(n: int, rank: int, world: int):
    """Round-robin share of ``n`` independent work items (patches of one volume, or volumes of a test set)."""
    return list(range(rank, n, world))


@torch.no_grad()
def predict_volume(gen, volume: torch.Tensor, patch=64, batch: int = 8, use_graph: bool | None = None,
                   group=None, shard: bool | None = None) -> torch.Tensor:
    """``volume``: (C,D,H,W) or (1,C,D,H,W) fp32 CUDA tensor. Returns the aggregated prediction
    (6,D,H,W) fp32 on the same device. ``gen`` is a ``unet_bssfp_b200.Generator`` (its train/eval mode is
    respected, as in the reference where ``predict_step`` runs under ``model.eval()``).
    ``use_graph`` (default off; ``UB_INFER_GRAPH=1`` turns it on, eval mode only): replay the generator forward as
    a CUDA graph on a static batch buffer; a short last batch is padded by repeating its last patch.

    Multi-GPU (``shard``; default: on when ``torch.distributed`` is initialised with more than one rank, every rank
    holding the same volume and weights): the patches are dealt round-robin over the ranks of ``group`` -- the
    generator work, which is all of the cost, needs no exchange -- and the per-rank partial volumes are combined so
    that the result is bit-identical to the single-process one, including the sampler-order rule "the later patch
    wins": every rank records which patch wrote each voxel last, an all-reduce (MAX) of that owner map decides the
    winner, and an all-reduce (SUM) of the partial volumes masked to the winners delivers it to every rank. (For
    a test SET, give each rank its own volumes -- ``shard_indices`` -- and no collective is needed at all.)"""
    vol = volume if volume.dim() == 4 else volume[0]
    if not vol.is_cuda:
        raise RuntimeError("predict_volume runs on CUDA tensors only (there is no CPU fallback)")
    patch = (patch,) * 3 if isinstance(patch, int) else tuple(patch)
    origins = grid_locations(tuple(vol.shape[1:]), patch)
    out_c = gen.blocks["unet"].out_channels
    import torch.distributed as dist
    dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if shard is None:
        shard = dist_on
    if shard and dist_on:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        mine = shard_indices(len(origins), rank, world)
        part = torch.zeros((out_c,) + tuple(vol.shape[1:]), dtype=torch.float32, device=vol.device)
        owner = torch.full(tuple(vol.shape[1:]), -1, dtype=torch.int32, device=vol.device)
        for i in range(0, len(mine), batch):
            idx = mine[i:i + batch]
            a = ops.pack_patches(vol, [origins[k] for k in idx], patch)
            y = gen.forward_packed(a)
            for j, k in enumerate(idx):          # increasing sampler index: locally the later patch wins
                z, yy, x = origins[k]
                ops.paste_patch(y[j], part, origins[k])
                owner[z:z + patch[0], yy:yy + patch[1], x:x + patch[2]] = k
        winner = owner.clone()
        dist.all_reduce(winner, op=dist.ReduceOp.MAX, group=group)
        part.mul_((owner == winner).unsqueeze(0))
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
        return part
    out = torch.zeros((out_c,) + tuple(vol.shape[1:]), dtype=torch.float32, device=vol.device)
    from .modules import _precision_of
    if use_graph is None:
        use_graph = _USE_GRAPH
    if use_graph and not gen.training and _precision_of(gen) == "bf16":
        n = min(batch, len(origins))
        fwd = _graphed_forward(gen, (n,) + patch + (ops.pad32(vol.shape[0]),), vol.device)
        for i in range(0, len(origins), n):
            group = origins[i:i + n]
            padded = group + [group[-1]] * (n - len(group))
            ops.pack_patches(vol, padded, patch, out=fwd.inp)
            y = fwd()                            # (n, 6, pd, ph, pw) fp32, the graph's static output
            for k, org in enumerate(group):      # sampler order: the later patch wins
                ops.paste_patch(y[k], out, org)
        return out
    for i in range(0, len(origins), batch):
        group = origins[i:i + batch]
        a = ops.pack_patches(vol, group, patch)
        y = gen.forward_packed(a)                # (n, 6, pd, ph, pw) fp32
        for k, org in enumerate(group):          # sampler order: the later patch wins
            ops.paste_patch(y[k], out, org)
    return out


def relative_error(pred: torch.Tensor, target: torch.Tensor, mask=None, probseg=None, angular: bool = False):
    """``pred`` / ``target``: (C,D,H,W) fp32 volumes (module layout). Returns ``(diff (D,H,W,C), errs [R][C] |
    None)``: the map of ``do_calc_diff_maps`` (channel-last, as the reference stores NIfTI volumes) and the
    per-ROI probseg-weighted means of ``do_calc_error_avg``."""
    # module layout (C,D,H,W) -> the channel-last layout of the reference's NIfTI volumes (one strided copy each)
    p = pred.permute(1, 2, 3, 0).contiguous().float()
    t = target.permute(1, 2, 3, 0).contiguous().float()
    diff, sums, norms = ops.relerr_map_reduce(p, t, mask, probseg, angular=angular)
    errs = None if sums is None else sums / norms[:, None]
    return diff, errs
