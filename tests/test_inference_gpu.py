"""Sliding-window inference (BASELINE config 5): patch gather / generator / later-wins aggregation on
device against the oracle restatement of torchio's GridSampler + GridAggregator, then the fused
relative-error evaluation against the NumPy oracle of ref:eval.py."""
import numpy as np
import pytest
import torch

from oracle import eval_oracle as E
from oracle import infer_oracle as I
from tests.util import bf16_round, rel_l2, strict_fp32

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_patch_gather_scatter_roundtrip():
    """ub_pack_patches / ub_unpack_patch == slicing + channel-last bf16 pack (bit-exact), later patch wins."""
    from unet_bssfp_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(0)
    vol = torch.randn((24, 20, 36, 28), device=DEV, generator=g)
    patch = (16, 16, 16)
    locs = I.grid_locations(vol.shape[1:], patch)
    assert locs == sorted(set(locs)) and locs[0] == (0, 0, 0) and locs[-1] == (4, 20, 12)
    a = ops.pack_patches(vol, locs[:5], patch)
    for k, (z, y, x) in enumerate(locs[:5]):
        want = vol[:, z:z + 16, y:y + 16, x:x + 16].permute(1, 2, 3, 0).to(torch.bfloat16)
        assert torch.equal(a[k, ..., :24], want) and a[k, ..., 24:].abs().max().item() == 0
    # scatter channels [2, 8) of every patch, in order, into a 6-channel volume
    allp = ops.pack_patches(vol, locs, patch)
    out = torch.full((6, 20, 36, 28), -7.0, device=DEV)
    for k, org in enumerate(locs):
        ops.unpack_patch(allp, k, 6, out, org, c_begin=2)
    ref = I.aggregate(torch.full((6, 20, 36, 28), -7.0, device=DEV),
                      [bf16_round(vol[2:8, z:z + 16, y:y + 16, x:x + 16]) for (z, y, x) in locs], locs)
    assert torch.equal(out, ref)
    with pytest.raises(RuntimeError):
        ops.pack_patches(vol, [(8, 0, 0)], patch)


def test_predict_volume_matches_oracle_aggregation():
    """Whole volume through predict_volume == the oracle sampler/aggregator driving the same generator
    patch by patch (bit-exact), and close to the fp32 oracle generator (bf16 bar)."""
    strict_fp32()
    import unet_bssfp_b200 as ub
    from oracle import model_oracle as O
    torch.manual_seed(0)
    og = O.Generator("bssfp").to(DEV).eval()
    g = ub.Generator("bssfp").to(DEV).eval()
    g.load_state_dict(og.state_dict())
    torch.manual_seed(3)
    vol = torch.rand((24, 48, 80, 40), device=DEV)          # 2 x 3 x 2 = 12 patches of 32^3, overlapping in every axis
    got = ub.inference.predict_volume(g, vol, patch=32, batch=5)
    same = I.predict_volume(lambda x: g(x), vol, patch=(32, 32, 32), batch=5)
    assert torch.equal(got, same)
    ref = I.predict_volume(og, vol, patch=(32, 32, 32), batch=4)
    assert rel_l2(got, ref) < 2e-2


def test_config5_geometry_and_eval():
    """160 x 192 x 160 with 64^3 patches: the 27 locations of SURVEY section 8d; a random-init generator
    over the whole volume; the relative-error map and ROI means against the NumPy oracle (1e-4)."""
    import unet_bssfp_b200 as ub
    shape = (160, 192, 160)
    locs = ub.inference.grid_locations(shape, (64, 64, 64))
    assert locs == I.grid_locations(shape, (64, 64, 64)) and len(locs) == 27
    assert sorted({z for z, _, _ in locs}) == [0, 64, 96] and sorted({y for _, y, _ in locs}) == [0, 64, 128]
    torch.manual_seed(0)
    g = ub.Generator("bssfp").to(DEV).eval()
    torch.manual_seed(5)
    vol = torch.rand((24,) + shape, device=DEV)
    pred = ub.inference.predict_volume(g, vol, patch=64, batch=9)
    assert pred.shape == (6,) + shape and torch.isfinite(pred).all()
    rng = np.random.default_rng(0)
    tgt = torch.from_numpy(rng.uniform(0.05, 1.0, size=(6,) + shape).astype(np.float32)).to(DEV)
    zz, yy, xx = np.meshgrid(*[np.arange(s) - s / 2 for s in shape], indexing="ij")
    mask = ((zz ** 2 + yy ** 2 + xx ** 2) < (0.4 * min(shape)) ** 2).astype(np.uint8)
    probseg = rng.dirichlet((1, 1, 1), size=shape).astype(np.float32)
    diff, errs = ub.inference.relative_error(pred, tgt, torch.from_numpy(mask).to(DEV), torch.from_numpy(probseg).to(DEV))
    p_np, t_np = pred.permute(1, 2, 3, 0).cpu().numpy(), tgt.permute(1, 2, 3, 0).cpu().numpy()
    ref = E.rel_error_map(p_np, t_np)
    ref_errs, _ = E.roi_error_avg(ref, mask, probseg)
    np.testing.assert_allclose(diff.cpu().numpy(), ref, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(errs.cpu().numpy(), ref_errs, rtol=1e-4)


def test_graph_replay_equals_eager_and_follows_weight_updates():
    """predict_volume with the CUDA-graph replay of the generator forward is bit-identical to eager launches,
    pads a short last batch, and FOLLOWS weight updates without a re-capture: the captured forward re-packs its
    bf16 weight operands from the live parameter memory on every replay."""
    import unet_bssfp_b200 as ub
    torch.manual_seed(0)
    g = ub.Generator("t1w").to(DEV).eval()
    torch.manual_seed(3)
    vol = torch.rand((6, 48, 80, 40), device=DEV)                     # 12 patches of 32^3: batches of 5, 5, 2
    eager = ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=False)
    graph = ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=True)
    assert torch.equal(eager, graph)
    again = ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=True)      # replay of the cached graph
    assert torch.equal(eager, again) and len(g._infer_graphs) == 1
    first = next(iter(g._infer_graphs.values()))
    with torch.no_grad():
        g.blocks["unet"].final_conv.bias.add_(0.5)                     # in-place update of a parameter the graph reads
    moved = ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=True)
    assert next(iter(g._infer_graphs.values())) is first                # same captured graph ...
    assert torch.equal(moved, ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=False))
    assert (moved - eager - 0.5).abs().max().item() < 1e-5              # ... and it saw the new bias
    with torch.no_grad():
        g.blocks["unet"].conv_0.conv_1.conv.weight.data.mul_(0.5)       # a conv weight, written out of band
    scaled = ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=True)
    assert next(iter(g._infer_graphs.values())) is first
    assert torch.equal(scaled, ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=False))
    assert not torch.equal(scaled, moved)
    g.train()                                                          # train mode (dropout, batch statistics): eager path
    torch.manual_seed(1); a = ub.inference.predict_volume(g, vol, patch=32, batch=5)
    torch.manual_seed(1); b = ub.inference.predict_volume(g, vol, patch=32, batch=5, use_graph=False)
    assert torch.equal(a, b)
