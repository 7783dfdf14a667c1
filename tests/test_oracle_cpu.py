"""CPU tests: the oracle against its structural known answers and
the committed golden vectors; hand-checkable cases of the evaluation arithmetic."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as E
from oracle import model_oracle as O
from tests.golden.make_golden import state_checksum, synth_eval_volumes

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_parameter_counts_and_keys():
    g, d = O.Generator("bssfp"), O.Discriminator("bssfp")
    assert sum(p.numel() for p in g.parameters()) == 22646182
    assert sum(p.numel() for p in d.parameters()) == 11230593
    assert sum(p.numel() for p in g.blocks["unet"].parameters()) == 22645318
    gs, ds = g.state_dict(), d.state_dict()
    assert len(gs) == 110 and len(ds) == 46
    assert tuple(gs["blocks.unet.upcat_1.convs.conv_0.conv.weight"].shape) == (32, 96, 3, 3, 3)
    assert tuple(gs["blocks.unet.upcat_4.upsample.deconv.weight"].shape) == (512, 256, 2, 2, 2)
    assert tuple(gs["blocks.unet.upcat_1.upsample.deconv.weight"].shape) == (64, 64, 2, 2, 2)
    assert tuple(gs["blocks.t1w.conv.weight"].shape) == (24, 6, 1, 1, 1)
    assert tuple(gs["blocks.unet.final_conv.weight"].shape) == (6, 32, 1, 1, 1)
    assert tuple(ds["d1.bssfp.conv.weight"].shape) == (32, 30, 4, 4, 4)
    assert tuple(ds["blocks.t1w.conv.weight"].shape) == (32, 12, 4, 4, 4)
    assert tuple(ds["final.weight"].shape) == (1, 512, 1, 1, 1)
    # shared head objects: registered twice, counted once
    assert g.blocks["bssfp"] is g.blocks["pc-bssfp"] and g.blocks["t1w"] is g.blocks["dwi-tensor"]
    assert d.d1 is d.blocks


def test_flop_census_matches_survey():
    """conv/deconv MACs x 2 per voxel: F_G = 554 176, F_D = 23 040 (bssfp)."""
    def conv_flops(m, vox_out):
        k = m.kernel_size[0] ** 3
        if isinstance(m, torch.nn.ConvTranspose3d):
            return 2 * m.in_channels * m.out_channels * k * vox_out / 8
        return 2 * m.in_channels * m.out_channels * k * vox_out
    g = O.Generator("bssfp")
    u = g.blocks["unet"]
    tot = conv_flops(g.blocks["bssfp"].conv, 1.0)
    res = {"conv_0": 1.0, "down_1": 1 / 8, "down_2": 1 / 64, "down_3": 1 / 512, "down_4": 1 / 4096,
           "upcat_4": 1 / 512, "upcat_3": 1 / 64, "upcat_2": 1 / 8, "upcat_1": 1.0}
    for name, frac in res.items():
        for m in getattr(u, name).modules():
            if isinstance(m, (torch.nn.Conv3d, torch.nn.ConvTranspose3d)):
                tot += conv_flops(m, frac)
    tot += conv_flops(u.final_conv, 1.0)
    assert round(tot) == 554176
    d = O.Discriminator("bssfp")
    fd = sum(conv_flops(m.conv, 1 / 8 ** (i + 1)) for i, m in enumerate([d.d1["bssfp"], d.d2, d.d3, d.d4, d.d5]))
    fd += conv_flops(d.final, 1 / 8 ** 5)
    assert abs(fd - 23040) < 1.0


@pytest.mark.parametrize("mod", ["bssfp", "t1w"])
def test_oracle_matches_golden(mod):
    torch.manual_seed(0)
    g, d = O.Generator(mod), O.Discriminator(mod)
    if abs(state_checksum(g) - float(GOLD[f"{mod}_g_checksum"])) > 1e-6 * float(GOLD[f"{mod}_g_checksum"]):
        pytest.skip("default torch init differs from the torch version that generated the goldens")
    cin = O.in_channels_of(mod)
    torch.manual_seed(1234)
    x, y = torch.rand(1, cin, 32, 32, 32), torch.rand(1, 6, 32, 32, 32)
    g.eval()
    with torch.no_grad():
        yh = g(x)
    assert yh.shape == (1, 6, 32, 32, 32)
    np.testing.assert_allclose(yh.numpy(), GOLD[f"{mod}_g_eval_32"], rtol=1e-4, atol=2e-5)
    d.train()
    with torch.no_grad():
        logits = d(torch.cat([x, x.flip(2)], 0), torch.cat([y, y.flip(3)], 0))
    assert logits.shape == (2, 1, 1, 1, 1)
    np.testing.assert_allclose(logits.numpy(), GOLD[f"{mod}_d_train_32_b2"], rtol=1e-3, atol=1e-4)


def test_gan_step_semantics():
    """ordering and freezing of ref:model.py:259-281: G phase leaves D untouched and vice versa."""
    torch.manual_seed(0)
    g, d = O.Generator("t1w"), O.Discriminator("t1w")
    opt_g, opt_d = O.make_optimizers(g, d)
    x, y = torch.rand(1, 6, 32, 32, 32), torch.rand(1, 6, 32, 32, 32)
    # batch of 2 so the PatchGAN's last BatchNorm sees more than one value per channel
    x, y = torch.cat([x, x.flip(2)]), torch.cat([y, y.flip(2)])
    g0 = {k: v.clone() for k, v in g.state_dict().items()}
    d0 = {k: v.clone() for k, v in d.state_dict().items()}
    gl, dl = O.gan_step(g, d, opt_g, opt_d, x, y)
    assert torch.isfinite(gl) and torch.isfinite(dl)
    assert not torch.equal(g.state_dict()["blocks.unet.final_conv.weight"], g0["blocks.unet.final_conv.weight"])
    assert not torch.equal(d.state_dict()["final.weight"], d0["final.weight"])
    # heads of the unused modality never receive a gradient
    assert torch.equal(g.state_dict()["blocks.bssfp.conv.weight"], g0["blocks.bssfp.conv.weight"])
    assert all(p.requires_grad for p in list(g.parameters()) + list(d.parameters()))
    assert all(p.grad is None for p in list(g.parameters()) + list(d.parameters()))


def test_eval_hand_cases():
    p = np.array([[[[2.0, 1.0]]]])
    t = np.array([[[[1.0, 0.0]]]])
    d = E.rel_error_map(p, t)
    assert d[0, 0, 0, 0] == 1.0 and np.isinf(d[0, 0, 0, 1])
    assert np.isnan(E.rel_error_map(np.zeros((1, 1, 1, 1)), np.zeros((1, 1, 1, 1)))[0, 0, 0, 0])
    a = E.rel_error_map(np.array([350.0, -170.0, 10.0]), np.array([10.0, 170.0, 350.0]), kind="azimuth")
    np.testing.assert_allclose(a, [20.0, 20.0, 20.0])
    errs, cleaned = E.roi_error_avg(d, np.ones((1, 1, 1)), np.ones((1, 1, 1, 2)))
    np.testing.assert_allclose(errs, [[1.0, 0.0], [1.0, 0.0]])       # inf -> 0
    errs, _ = E.roi_error_avg(d, np.zeros((1, 1, 1)), np.ones((1, 1, 1, 1)))
    np.testing.assert_allclose(errs, [[0.0, 0.0]])                   # outside the mask


def test_eval_matches_golden():
    pred, tgt, mask, probseg = synth_eval_volumes()
    diff = E.rel_error_map(pred, tgt)
    errs, _ = E.roi_error_avg(diff, mask, probseg)
    np.testing.assert_allclose(diff.astype(np.float32), GOLD["eval_diff_rel"], rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(errs, GOLD["eval_errs_rel"], rtol=1e-12, equal_nan=True)
    assert np.isnan(errs[:, 0]).all() and np.isfinite(errs[:, 1:]).all()   # the 0/0 voxel poisons channel 0 only


def test_grid_sampler_restatement_kat():
    """torchio GridSampler (overlap 0) locations: SURVEY section 8d KAT and the host logic of the product."""
    import importlib
    from oracle import infer_oracle as I
    locs = I.grid_locations((160, 192, 160), (64, 64, 64))
    assert len(locs) == 27 and locs[0] == (0, 0, 0) and locs[-1] == (96, 128, 96)
    assert locs == sorted(locs)
    assert I.grid_locations((64, 64, 64), (64, 64, 64)) == [(0, 0, 0)]
    assert I.grid_locations((96, 128, 128), (64, 64, 64)) == [(z, y, x) for z in (0, 32) for y in (0, 64) for x in (0, 64)]
    inf = importlib.import_module("unet_bssfp_b200.inference")
    for shape, patch in [((160, 192, 160), (64, 64, 64)), ((96, 128, 128), (64, 64, 64)), ((70, 64, 100), (32, 64, 48))]:
        assert inf.grid_locations(shape, patch) == I.grid_locations(shape, patch)
    # later patch wins
    import torch
    out = I.aggregate(torch.zeros((1, 4, 4, 6)), [torch.full((1, 4, 4, 4), 1.0), torch.full((1, 4, 4, 4), 2.0)],
                      [(0, 0, 0), (0, 0, 2)])
    assert out[0, 0, 0].tolist() == [1, 1, 2, 2, 2, 2]


def test_dti_oracle_hand_cases():
    """Known answers of the DTI scalar maps (ref:eval.py:73-116 formulas)."""
    from oracle import eval_oracle as E
    t = np.array([[3e-3, 0, 0, 2e-3, 0, 1e-3],        # diagonal, principal axis x
                  [2e-3, 0, 0, 2e-3, 0, 2e-3],        # isotropic
                  [1.0, 0, 0, 0.0, 0, 0.0]])          # stick along x: FA = 1
    m = E.dti_scalar_maps(t)
    assert np.allclose(m["ad"], [3e-3, 2e-3, 1.0]) and np.allclose(m["rd"], [1.5e-3, 2e-3, 0.0])
    assert np.allclose(m["md"], [2e-3, 2e-3, 1 / 3])
    lam = np.array([1e-3, 2e-3, 3e-3])
    fa0 = np.sqrt(1.5) * np.sqrt(((lam - 2e-3) ** 2).sum()) / np.sqrt((lam ** 2).sum())
    assert np.allclose(m["fa"], [fa0, 0.0, 1.0])
    assert np.allclose(np.abs(m["inclination"][[0, 2]]), 90.0) and np.allclose(m["rgb"][2], [1.0, 0.0, 0.0])
