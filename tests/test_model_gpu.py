"""GPU parity of the drop-in modules (sm_100a kernels through the C ABI) against the torch.nn oracle
on the same weights and inputs, and against the committed goldens.

Tolerances. north_star asks for 1e-2 relative in bf16. Per-op on identical inputs that bar is met
(tests/test_conv_gpu.py). End to end through 24 stacked layers the generator's eval forward is at
rel-L2 6.8e-3 ... 8.9e-3 against the fp32 oracle for every shape tested here (32^3 ... 8 x 128^3, both
modalities; tools/parity_values.py, profiles/r02f_parity_values.txt) -- torch's own bf16 autocast of the
oracle is at 1.5e-2 ... 2.2e-2 on the same inputs -- so the forward tests assert the literal 1e-2.
Weight gradients: LeakyReLU / max-pool sign flips make per-tensor errors of ANY bf16 evaluation ~0.3 at
the deep levels, so gradient tests use the yardstick "every error <= 1.25 x the
error of the oracle run under torch.autocast(bf16), measured in the same test", plus the well-conditioned
VJP check against the fp64 oracle further down."""
import copy
import os

import numpy as np
import pytest
import torch

from tests.golden.make_golden import state_checksum
from tests.util import no_dropout, rel_l2, strict_fp32

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
DEV = "cuda"


def _pair(mod, seed=0):
    import unet_bssfp_b200 as ub
    from oracle import model_oracle as O
    torch.manual_seed(seed)
    og, od = O.Generator(mod).to(DEV), O.Discriminator(mod).to(DEV)
    g, d = ub.Generator(mod).to(DEV), ub.Discriminator(mod).to(DEV)
    g.load_state_dict(og.state_dict())
    d.load_state_dict(od.state_dict())
    return O, og, od, g, d


def _no_dropout(og, g):
    no_dropout(og, g)


@pytest.mark.parametrize("mod,shape", [("bssfp", (1, 32, 32, 32)), ("t1w", (2, 32, 48, 32)), ("bssfp", (1, 64, 64, 64))])
def test_generator_eval_forward(mod, shape):
    strict_fp32()
    O, og, od, g, d = _pair(mod)
    n, dd, hh, ww = shape
    torch.manual_seed(1234)
    x = torch.rand(n, O.in_channels_of(mod), dd, hh, ww, device=DEV)
    og.eval(); g.eval()
    with torch.no_grad():
        ref, got = og(x), g(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yard = og(x).float()
    assert got.shape == ref.shape and got.dtype == torch.float32
    e, ey = rel_l2(got, ref), rel_l2(yard, ref)
    assert e < 1e-2, e                       # the north_star bar, literally (measured 6.8e-3 ... 7.3e-3)
    assert e < ey, (e, ey)                   # and closer to fp32 than stock bf16 autocast (1.5e-2 ... 1.7e-2)


@pytest.mark.parametrize("mod", ["bssfp", "t1w"])
def test_generator_matches_golden(mod):
    import unet_bssfp_b200 as ub
    torch.manual_seed(0)
    g = ub.Generator(mod)
    if abs(state_checksum(g) - float(GOLD[f"{mod}_g_checksum"])) > 1e-6 * float(GOLD[f"{mod}_g_checksum"]):
        pytest.skip("default torch init differs from the torch version that generated the goldens")
    g = g.to(DEV).eval()
    cin = 24 if mod == "bssfp" else 6
    torch.manual_seed(1234)
    x = torch.rand(1, cin, 32, 32, 32)
    with torch.no_grad():
        got = g(x.to(DEV)).cpu()
    ref = torch.from_numpy(GOLD[f"{mod}_g_eval_32"])
    assert rel_l2(got, ref) < 1e-2


def test_generator_train_backward_vs_autocast_yardstick():
    strict_fp32()
    O, og, od, g, d = _pair("bssfp")
    _no_dropout(og, g)
    og.train(); g.train()
    torch.manual_seed(1234)
    x = torch.rand(1, 24, 64, 64, 64, device=DEV)
    ref = og(x)
    dY = torch.randn_like(ref)
    ref.backward(dY)
    got = g(x)
    got.backward(dY)
    oa = copy.deepcopy(og)
    for p in oa.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = oa(x)
    ya.float().backward(dY)
    assert rel_l2(got, ref) < 2e-2
    ours, yard = [], []
    for (n1, p1), (n2, p2), (n3, p3) in zip(og.named_parameters(), g.named_parameters(), oa.named_parameters()):
        assert n1 == n2
        assert (p1.grad is None) == (p2.grad is None), n1         # same set of live parameters
        if p1.grad is None:
            continue
        if n1.endswith("conv.bias") and "final_conv" not in n1 and "deconv" not in n1:
            # bias feeding a batch-statistics norm: analytically zero gradient (compare absolutely)
            assert p2.grad.abs().max().item() <= 1e-5 + p1.grad.abs().max().item()
            continue
        ours.append(rel_l2(p2.grad, p1.grad)); yard.append(rel_l2(p3.grad, p1.grad))
        assert ours[-1] < 1.25 * yard[-1] + 2e-2, (n1, ours[-1], yard[-1])
    assert np.median(ours) < 1.1 * np.median(yard) + 1e-2
    fo = torch.cat([p.grad.flatten() for p in og.parameters() if p.grad is not None and p.ndim > 1])
    fp = torch.cat([p.grad.flatten() for p in g.parameters() if p.grad is not None and p.ndim > 1])
    fa = torch.cat([p.grad.flatten() for p in oa.parameters() if p.grad is not None and p.ndim > 1])
    assert rel_l2(fp, fo) < 1.1 * rel_l2(fa, fo) + 1e-2
    # BatchNorm running statistics of the head follow torch (momentum 0.1, unbiased variance)
    assert rel_l2(g.blocks["bssfp"].bn.running_var, og.blocks["bssfp"].bn.running_var) < 1e-3
    assert int(g.blocks["bssfp"].bn.num_batches_tracked) == int(og.blocks["bssfp"].bn.num_batches_tracked)


def test_generator_dropout_train_mode_is_statistical():
    O, og, od, g, d = _pair("t1w")
    g.train()
    torch.manual_seed(5)
    x = torch.rand(1, 6, 32, 32, 32, device=DEV)
    with torch.no_grad():
        torch.manual_seed(11); a = g(x)
        torch.manual_seed(11); b = g(x)
        torch.manual_seed(12); c = g(x)
    assert torch.equal(a, b)                 # mask stream follows torch.manual_seed
    assert not torch.equal(a, c)
    # the size of the perturbation matches what torch's own Dropout(0.05) does to the oracle
    og.train()
    with torch.no_grad():
        torch.manual_seed(11); ao = og(x)
        torch.manual_seed(12); co = og(x)
    ours, theirs = rel_l2(c, a), rel_l2(co, ao)
    assert 0.5 * theirs < ours < 2.0 * theirs, (ours, theirs)


@pytest.mark.parametrize("mod", ["bssfp", "t1w"])
def test_discriminator_forward_backward(mod):
    strict_fp32()
    O, og, od, g, d = _pair(mod)
    od.train(); d.train()
    torch.manual_seed(1234)
    n, s = 4, 64
    x = torch.rand(n, O.in_channels_of(mod), s, s, s, device=DEV)
    y = torch.rand(n, 6, s, s, s, device=DEV)
    yo, yp = y.clone().requires_grad_(True), y.clone().requires_grad_(True)
    lo, lp = od(x, yo), d(x, yp)
    assert lp.shape == lo.shape == (n, 1, 2, 2, 2)
    oa = copy.deepcopy(od)
    ya_in = y.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la = oa(x, ya_in)
    e, ey = rel_l2(lp, lo), rel_l2(la.float(), lo)
    assert e < 3e-2 and e < 1.25 * ey + 5e-3, (e, ey)
    dl = torch.randn_like(lo)
    lo.backward(dl); lp.backward(dl); la.float().backward(dl)
    assert rel_l2(yp.grad, yo.grad) < 1.25 * rel_l2(ya_in.grad, yo.grad) + 2e-2
    for (n1, p1), (n2, p2), (n3, p3) in zip(od.named_parameters(), d.named_parameters(), oa.named_parameters()):
        assert (p1.grad is None) == (p2.grad is None), n1
        if p1.grad is None:
            continue
        if n1.endswith("conv.bias") and not n1.startswith("d1"):
            assert p2.grad.abs().max().item() <= 1e-5 + p1.grad.abs().max().item()
            continue
        assert rel_l2(p2.grad, p1.grad) < 1.25 * rel_l2(p3.grad, p1.grad) + 2e-2, n1
    assert rel_l2(d.d3.bn.running_mean, od.d3.bn.running_mean) < 1e-2


def test_frozen_parameters_and_phase_semantics():
    """ref:model.py:264,274 -- toggle_optimizer: D frozen in the G phase (dgrad only), G frozen in the D phase."""
    O, og, od, g, d = _pair("bssfp")
    from unet_bssfp_b200.train_step import GanTrainer
    _no_dropout(og, g)
    tr = GanTrainer(g, d)
    torch.manual_seed(3)
    x = torch.rand(2, 24, 32, 32, 32, device=DEV)
    y = torch.rand(2, 6, 32, 32, 32, device=DEV)
    for p in d.parameters():
        p.requires_grad_(False)
    gl, _ = tr.gen_loss(x, y)
    gl.backward()
    assert all(p.grad is None for p in d.parameters())
    live = [n for n, p in g.named_parameters() if p.grad is not None]
    assert "blocks.pc-bssfp.conv.weight" in live and not any(n.startswith("blocks.dwi-tensor") for n in live)
    ogl, _ = O.gen_loss(og, od, x, y)
    assert abs(gl.item() - ogl.item()) < 0.03 * abs(ogl.item())
    for p in d.parameters():
        p.requires_grad_(True)
    g.zero_grad(set_to_none=True)
    for p in g.parameters():
        p.requires_grad_(False)
    dl = tr.discr_loss(x, y)
    dl.backward()
    assert all(p.grad is None for p in g.parameters())
    assert all(p.grad is not None for n, p in d.named_parameters() if "dwi-tensor" not in n and "t1w" not in n)
    odl = O.discr_loss(og, od, x, y)
    assert abs(dl.item() - odl.item()) < 0.03 * abs(odl.item()) + 0.01


def test_full_step_runs_and_updates_both_networks():
    O, og, od, g, d = _pair("t1w")
    from unet_bssfp_b200.train_step import GanTrainer
    from unet_bssfp_b200 import _lib
    tr = GanTrainer(g, d)
    torch.manual_seed(3)
    x = torch.rand(2, 6, 32, 32, 32, device=DEV)
    y = torch.rand(2, 6, 32, 32, 32, device=DEV)
    w_g = g.blocks["unet"].final_conv.weight.detach().clone()
    w_d = d.final.weight.detach().clone()
    n0 = _lib.load().ub_launch_count()
    for _ in range(2):
        gl, dl = tr.step(x, y)
    torch.cuda.synchronize()
    assert _lib.load().ub_launch_count() - n0 > 500          # the CUDA path ran (no fallback exists)
    assert torch.isfinite(gl) and torch.isfinite(dl)
    assert not torch.equal(w_g, g.blocks["unet"].final_conv.weight) and not torch.equal(w_d, d.final.weight)
    assert all(p.requires_grad for p in list(g.parameters()) + list(d.parameters()))


def test_step_shares_discriminator_operand_packs_bit_identically(monkeypatch):
    """``GanTrainer.step`` declares the discriminator constant up to ``opt_d.step()`` (``weights_unchanged``): its six
    passes share two operand packs. Same losses and same parameters, bit for bit, as with a pack per pass -- and
    after the scope a pass packs from the live parameters again (an out-of-band write is seen)."""
    import contextlib
    from unet_bssfp_b200 import modules as M, ops
    from unet_bssfp_b200.train_step import GanTrainer
    torch.manual_seed(5)
    x = torch.rand(2, 6, 32, 32, 32, device=DEV)
    y = torch.rand(2, 6, 32, 32, 32, device=DEV)
    results, packs = [], []
    real_pack = ops.pack_conv_weights_multi
    for shared in (True, False):
        O, og, od, g, d = _pair("t1w")
        for m in g.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        if not shared:
            monkeypatch.setattr(M, "weights_unchanged", lambda module: contextlib.nullcontext())
        count = [0]
        def counting(items, _c=count):
            _c[0] += 1
            return real_pack(items)
        monkeypatch.setattr(ops, "pack_conv_weights_multi", counting)
        tr = GanTrainer(g, d)
        out = [tr.step(x, y) for _ in range(2)]
        torch.cuda.synchronize()
        results.append(([t.item() for pair in out for t in pair], [p.detach().clone() for p in d.parameters()]))
        packs.append(count[0])
        monkeypatch.setattr(ops, "pack_conv_weights_multi", real_pack)
        if shared:
            assert not d._cache.hold and not d._cache._cur          # nothing outlives the scope
            d.eval()
            with torch.no_grad():
                before = d(x, y).clone()
                d.final.weight.data.mul_(0.0)
                after = d(x, y)
            assert not torch.equal(before, after)
    assert results[0][0] == results[1][0]
    assert all(torch.equal(a, b) for a, b in zip(results[0][1], results[1][1]))
    assert packs[1] - packs[0] == 2 * 4       # four discriminator packs less per step


@pytest.mark.parametrize("optimizer", ["ub", "torch_fused", "torch"])
def test_updated_weights_are_used_after_optimizer_steps(optimizer):
    """The bf16 operand copies of the weights must follow the fp32 Parameters through optimizer steps.
    torch's fused AdamW and this package's FusedAdamW kernel write parameters without bumping ``tensor._version`` (regression: the packed
    weights stayed at their initial values): after two full steps the modules must agree with an oracle
    that is loaded with the UPDATED state dict -- and disagree with the initial-weight output."""
    strict_fp32()
    O, og, od, g, d = _pair("bssfp")
    from unet_bssfp_b200.train_step import GanTrainer
    tr = GanTrainer(g, d, lr=2e-2, optimizer=optimizer)
    torch.manual_seed(3)
    x = torch.rand(1, 24, 32, 32, 32, device=DEV)
    y = torch.rand(1, 6, 32, 32, 32, device=DEV)
    g.eval()
    with torch.no_grad():
        before = g(x).clone()
        logit_before = d(x, y).clone()
    g.train()
    for _ in range(2):
        tr.step(x, y)
    og.load_state_dict(g.state_dict())
    od.load_state_dict(d.state_dict())
    g.eval(); og.eval(); d.eval(); od.eval()
    with torch.no_grad():
        after, ref = g(x), og(x)
        logit_after, logit_ref = d(x, y), od(x, y)
    assert rel_l2(after, ref) < 3e-2, rel_l2(after, ref)                     # follows the updated weights
    assert rel_l2(after, before) > 10 * rel_l2(after, ref)                    # ... and they did change the output
    assert rel_l2(logit_after, logit_ref) < 3e-2
    assert rel_l2(logit_after, logit_before) > 5 * rel_l2(logit_after, logit_ref)
    with torch.no_grad():
        w = g.blocks["unet"].final_conv.weight
        w.detach().view(-1)[:].mul_(0.0)
        assert g(x).abs().max().item() <= g.blocks["unet"].final_conv.bias.abs().max().item() + 1e-6


def test_out_of_band_weight_writes_are_seen():
    """writes through ``.data`` change neither ``_version`` nor ``data_ptr`` nor any optimizer
    stamp. The bf16 weight operands are re-packed from the live parameter memory on every pass, so such writes --
    re-initialisation after a first forward, EMA swaps, a foreign kernel -- take effect, in forward AND backward,
    exactly as with the reference, which reads its fp32 weights afresh in every call."""
    O, og, od, g, d = _pair("bssfp", seed=5)
    no_dropout(og, g)
    torch.manual_seed(11)
    x = torch.rand(1, 24, 32, 32, 32, device=DEV)
    y = torch.rand(1, 6, 32, 32, 32, device=DEV)
    g.eval(); og.eval(); d.eval(); od.eval()
    with torch.no_grad():
        g(x); d(x, y)                                   # first forward: operands packed once
    # out-of-band writes: a 3x3x3 conv (marching kernel), a deep conv (generic kernel), a transposed conv, the stem
    torch.manual_seed(12)
    touched_g = [g.blocks["unet"].conv_0.conv_1.conv.weight, g.blocks["unet"].down_3.convs.conv_0.conv.weight,
                 g.blocks["unet"].upcat_2.upsample.deconv.weight]
    touched_d = [d.d1["bssfp"].conv.weight, d.d3.conv.weight]
    for w in touched_g + touched_d:
        v0, p0 = w._version, w.data_ptr()
        w.data.normal_(0.0, 0.05)
        assert w._version == v0 and w.data_ptr() == p0     # nothing a version-keyed cache could have noticed
    og.load_state_dict(g.state_dict()); od.load_state_dict(d.state_dict())
    with torch.no_grad():
        e_g = rel_l2(g(x), og(x))
        e_d = rel_l2(d(x, y), od(x, y))
    assert e_g < 3e-2 and e_d < 3e-2, (e_g, e_d)
    # backward (dgrad operand copies) after another out-of-band write
    g.train(); og.train()
    w = g.blocks["unet"].upcat_1.convs.conv_1.conv.weight
    xb = x.clone().requires_grad_(True)
    g(xb).square().mean().backward()                     # gradient with the OLD conv_1 weights
    w.data.mul_(-1.5)
    og.load_state_dict(g.state_dict())
    xg = x.clone().requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    g(xg).square().mean().backward()
    og(xo).square().mean().backward()
    # an input gradient through 24 bf16 layers carries ~20 % rounding / sign-flip noise whoever computes it (see
    # test_generator_vjp_well_conditioned_vs_fp64_oracle); a stale operand would leave it at the old value instead
    e_new, e_old = rel_l2(xg.grad, xo.grad), rel_l2(xb.grad, xo.grad)
    assert e_new < 0.35 and e_old > 2 * e_new, (e_new, e_old)


def test_shape_errors():
    import unet_bssfp_b200 as ub
    g, d = ub.Generator("t1w").to(DEV), ub.Discriminator("t1w").to(DEV)
    with pytest.raises(RuntimeError, match="divisible by 16"):
        g(torch.rand(1, 6, 24, 32, 32, device=DEV))
    with pytest.raises(RuntimeError, match="divisible by 32"):
        d(torch.rand(1, 6, 48, 32, 32, device=DEV), torch.rand(1, 6, 48, 32, 32, device=DEV))


def test_discriminator_phase_on_two_streams_equals_one_stream():
    """GanTrainer.discr_loss runs the real-sample pass on a second stream (deferred BatchNorm running-statistics
    updates keep the reference's order): loss, every gradient and the running statistics must equal the
    single-stream evaluation."""
    import unet_bssfp_b200 as ub
    from unet_bssfp_b200.train_step import GanTrainer
    O, og, od, g, d = _pair("bssfp")
    _no_dropout(og, g)
    g2, d2 = ub.Generator("bssfp").to(DEV), ub.Discriminator("bssfp").to(DEV)
    g2.load_state_dict(g.state_dict()); d2.load_state_dict(d.state_dict())
    no_dropout(g2)
    ta, tb = GanTrainer(g, d), GanTrainer(g2, d2)
    ta.overlap_real_branch, tb.overlap_real_branch = True, False
    torch.manual_seed(3)
    x = torch.rand(2, 24, 64, 32, 32, device=DEV)
    y = torch.rand(2, 6, 64, 32, 32, device=DEV)
    for net in (g, g2):
        for p in net.parameters():
            p.requires_grad_(False)
    for _ in range(2):
        la, lb = ta.discr_loss(x, y), tb.discr_loss(x, y)
        la.backward(); lb.backward()
        torch.cuda.synchronize()
    assert ta._branch_stream is not None and tb._branch_stream is None
    assert torch.equal(la, lb)
    for (n1, p1), (n2, p2) in zip(d.named_parameters(), d2.named_parameters()):
        assert (p1.grad is None) == (p2.grad is None), n1
        if p1.grad is not None:
            assert torch.equal(p1.grad, p2.grad), n1
    for m1, m2 in zip(d.modules(), d2.modules()):
        if isinstance(m1, torch.nn.BatchNorm3d):
            assert int(m1.num_batches_tracked) == int(m2.num_batches_tracked) == 4
            torch.testing.assert_close(m1.running_mean, m2.running_mean, rtol=1e-6, atol=1e-7)
            torch.testing.assert_close(m1.running_var, m2.running_var, rtol=1e-5, atol=1e-7)
    # and both follow the oracle's running statistics (fake pass first, then the real pass)
    od.train(); og.train()
    for _ in range(2):
        O.discr_loss(og, od, x, y)
    assert rel_l2(d.d2.bn.running_mean, od.d2.bn.running_mean) < 2e-2
    assert rel_l2(d.d3.bn.running_var, od.d3.bn.running_var) < 2e-2


def test_config4_full_size_t1w_batch16():
    """BASELINE config 4 at full size (t1w: 6 input channels, batch 16, 128^3). Size-independent properties:
    samples are independent in eval mode (InstanceNorm per sample, BatchNorm on running statistics), so a slice of
    the batch output equals the output of the sliced batch; one full optimisation step runs and stays finite."""
    import unet_bssfp_b200 as ub
    from unet_bssfp_b200.train_step import GanTrainer
    torch.manual_seed(0)
    g, d = ub.Generator("t1w").to(DEV), ub.Discriminator("t1w").to(DEV)
    torch.manual_seed(7)
    x = torch.rand(16, 6, 128, 128, 128, device=DEV)
    y = torch.rand(16, 6, 128, 128, 128, device=DEV)
    g.eval()
    with torch.no_grad():
        full = g(x)
        part = g(x[5:7])
    assert full.shape == (16, 6, 128, 128, 128)
    # not bitwise: the launch geometry (depth segments per CTA, split-K) depends on the batch size, so fp32 sums
    # associate differently and a few bf16 roundings flip per layer (measured 4.5e-3 after 24 layers); leaking
    # statistics or data between samples would show as an O(1) error
    assert rel_l2(full[5:7], part) < 1e-2
    del full, part
    g.train()
    tr = GanTrainer(g, d)
    gl, dl = tr.step(x, y)
    torch.cuda.synchronize()
    assert torch.isfinite(gl) and torch.isfinite(dl)
    assert 0.3 < float(dl) < 1.5                       # (BCE(real, 1) + BCE(fake, 0)) / 2 near ln 2 at initialisation


def test_config2_full_size_properties():
    """BASELINE config 2 at full size (bssfp, batch 8, 128^3): size-independent properties. Eval-mode samples are
    independent (generator: InstanceNorm; discriminator: running statistics); the fused L1 / BCE losses equal their
    definitions on the full tensors; the discriminator's logits of a permuted batch are the permuted logits."""
    import unet_bssfp_b200 as ub
    torch.manual_seed(0)
    g, d = ub.Generator("bssfp").to(DEV).eval(), ub.Discriminator("bssfp").to(DEV).eval()
    torch.manual_seed(11)
    x = torch.rand(8, 24, 128, 128, 128, device=DEV)
    y = torch.rand(8, 6, 128, 128, 128, device=DEV)
    with torch.no_grad():
        y_hat = g(x)
        assert y_hat.shape == (8, 6, 128, 128, 128) and torch.isfinite(y_hat).all()
        assert rel_l2(y_hat[2:3], g(x[2:3])) < 1e-2                    # bf16 noise level, see the config-4 test
        logits = d(x, y_hat)
        assert logits.shape == (8, 1, 4, 4, 4)
        perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4], device=DEV)
        assert rel_l2(d(x[perm], y_hat[perm]), logits[perm]) < 1e-2
        l1 = ub.L1Loss()(y_hat, y)
        assert abs(l1.item() - (y_hat - y).abs().mean().item()) < 1e-6 * max(1.0, l1.item())
        bce = ub.BCEWithLogitsLoss()(logits, torch.ones_like(logits))
        ref = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.ones_like(logits))
        assert abs(bce.item() - ref.item()) < 1e-6


@pytest.mark.parametrize("mod,batch", [("bssfp", 1), ("bssfp", 8), ("t1w", 1), ("t1w", 8)])
def test_full_size_128_eval_forward_vs_oracle(mod, batch):
    """VERDICT r1 missing #5: the generator and the discriminator at the size the bench runs (128^3; batch 1 and the
    bench's batch 8) against the fp32 oracle on the same weights and inputs -- the 16-w-tile CTA-pair geometry, the
    128-plane marching segments, the deferred operand transforms and the two-stage BatchNorm statistics (32 768 tile
    records) compared VALUE BY VALUE, not through size-independent properties."""
    strict_fp32()
    O, og, od, g, d = _pair(mod, seed=3)
    cin = O.in_channels_of(mod)
    torch.manual_seed(4321)
    x = torch.rand(batch, cin, 128, 128, 128, device=DEV)
    y = torch.rand(batch, 6, 128, 128, 128, device=DEV)
    g.eval(); og.eval()
    with torch.no_grad():
        got = g(x)
        ref = og(x)
    e_g = rel_l2(got, ref)
    print(f"\n[{mod} batch {batch} @128^3] generator eval rel-L2 vs fp32 oracle: {e_g:.3e}")
    assert got.shape == ref.shape and e_g < 1e-2, e_g           # measured 8.6e-3 (stock autocast: 1.7e-2 ... 2.2e-2)
    # the discriminator in TRAIN mode: its BatchNorm layers take batch statistics (over 8 x 64^3 ... 8 x 4^3 values)
    d.train(); od.train()
    with torch.no_grad():
        lg, lr = d(x, y), od(x, y)
    e_d = rel_l2(lg, lr)
    print(f"[{mod} batch {batch} @128^3] discriminator (train-mode BN) logits rel-L2: {e_d:.3e}")
    assert lg.shape == lr.shape == (batch, 1, 4, 4, 4) and e_d < 2e-2, e_d
    for name in ("d2", "d5"):
        a, b = getattr(d, name).bn, getattr(od, name).bn
        assert rel_l2(a.running_mean, b.running_mean) < 2e-2 and rel_l2(a.running_var, b.running_var) < 2e-2
    # the input head's BatchNorm in train mode: batch statistics over batch x 128^3 values per channel (two-stage
    # reduction of the per-tile partials), checked through the running buffers it updates
    g.train(); og.train()
    no_dropout(g, og)
    with torch.no_grad():
        got_t, ref_t = g(x), og(x)
    assert rel_l2(got_t, ref_t) < 1.5e-2
    # (the head convolves the bf16-rounded input with bf16-rounded weights: its batch moments differ from the fp32
    # ones at the 1e-3 level)
    assert rel_l2(g.blocks[mod].bn.running_var, og.blocks[mod].bn.running_var) < 3e-3
    assert rel_l2(g.blocks[mod].bn.running_mean, og.blocks[mod].bn.running_mean) < 5e-3


def test_bf16_input_is_a_bit_identical_transport_format():
    """A loader may ship the conditioning input x in bf16 (half the host-to-device bytes): the pack kernel reads it
    as it is, the packed operand has the same bits as packing the fp32 tensor (which the kernel rounds to bf16
    anyway), outputs stay fp32 and the whole G / D forward and backward are bit-identical."""
    import unet_bssfp_b200 as ub
    torch.manual_seed(0)
    g, d = ub.Generator("bssfp").to(DEV), ub.Discriminator("bssfp").to(DEV)
    no_dropout(g)
    torch.manual_seed(2)
    x16 = torch.rand(2, 24, 32, 32, 32, device=DEV).to(torch.bfloat16)
    x32 = x16.float()
    y = torch.rand(2, 6, 32, 32, 32, device=DEV)
    assert torch.equal(ub.ops.pack_ncdhw(x16), ub.ops.pack_ncdhw(x32))
    assert torch.equal(ub.ops.pack_ncdhw(x16, y, s2d=True), ub.ops.pack_ncdhw(x32, y, s2d=True))
    ub.ops._WIDEN_BF16_INPUT = False            # the kernels' own bf16 read path (ub_pack_ncdhw with a_bf16 = 1)
    try:
        assert torch.equal(ub.ops.pack_ncdhw(x16), ub.ops.pack_ncdhw(x32))
        assert torch.equal(ub.ops.pack_ncdhw(x16, y, s2d=True), ub.ops.pack_ncdhw(x32, y, s2d=True))
        odd = x16[:, :, :5, :5, :5].contiguous()        # odd voxel count: the one-voxel-per-thread kernel
        assert torch.equal(ub.ops.pack_ncdhw(odd), ub.ops.pack_ncdhw(odd.float()))
    finally:
        ub.ops._WIDEN_BF16_INPUT = True
    outs = []
    for x in (x16, x32):
        for p in list(g.parameters()) + list(d.parameters()):
            p.grad = None
        y_hat = g(x)
        logits = d(x, y_hat)
        assert y_hat.dtype == torch.float32 and logits.dtype == torch.float32
        (logits.mean() + (y_hat - y).abs().mean()).backward()
        outs.append((y_hat.detach(), logits.detach(), [p.grad.clone() for p in g.parameters() if p.grad is not None]))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert all(torch.equal(a, b) for a, b in zip(outs[0][2], outs[1][2]))


def _smooth_field(shape, seed):
    """A smooth, fixed upstream gradient: a few low-frequency plane waves per channel."""
    n, c, d, h, w = shape
    g = torch.Generator(device="cpu").manual_seed(seed)
    zz, yy, xx = torch.meshgrid(torch.linspace(0, 1, d), torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    out = torch.zeros(shape)
    for i in range(n):
        for j in range(c):
            k = torch.rand(3, 3, generator=g) * 3.0
            ph = torch.rand(3, generator=g) * 6.28
            out[i, j] = sum(torch.sin(6.28 * (k[t, 0] * zz + k[t, 1] * yy + k[t, 2] * xx) + ph[t]) for t in range(3)) / 3.0
    return out.to(DEV)


def test_generator_vjp_well_conditioned_vs_fp64_oracle():
    """an end-to-end gradient check that is well conditioned and deterministic -- eval mode
    (running statistics in the head, no dropout), a smooth fixed upstream gradient, truth = the oracle in fp64.
    Per weight tensor: the cosine with the fp64 gradient and the rel-L2 error against the yardstick of stock
    PyTorch in the same precision class (bf16 autocast, cuDNN). No escape clauses."""
    strict_fp32()
    O, og, od, g, d = _pair("bssfp", seed=1)
    g.eval(); og.eval()
    torch.manual_seed(99)
    x = torch.rand(1, 24, 64, 64, 64, device=DEV)
    dY = _smooth_field((1, 6, 64, 64, 64), 5)
    o64 = copy.deepcopy(og).double()
    o64(x.double()).backward(dY.double())
    g(x).backward(dY)
    oa = copy.deepcopy(og)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = oa(x)
    ya.float().backward(dY)
    rows = []
    for (n1, p64), (_, pg), (_, pa) in zip(o64.named_parameters(), g.named_parameters(), oa.named_parameters()):
        if p64.grad is None or p64.ndim < 2:
            assert (pg.grad is None) == (p64.grad is None), n1
            continue
        t = p64.grad.float()
        cos = torch.nn.functional.cosine_similarity(pg.grad.flatten(), t.flatten(), dim=0).item()
        cos_y = torch.nn.functional.cosine_similarity(pa.grad.float().flatten(), t.flatten(), dim=0).item()
        rows.append((n1, cos, rel_l2(pg.grad, t), cos_y, rel_l2(pa.grad.float(), t)))
    assert len(rows) >= 24
    worst = min(rows, key=lambda r: r[1])
    print("\nper-tensor gradient cosine (ours / autocast) and rel-L2 (ours / autocast) vs the fp64 oracle:")
    for r in rows:
        print(f"  {r[0]:52s} {r[1]:.5f} {r[3]:.5f}   {r[2]:.3e} {r[4]:.3e}")
    # What bf16 operands allow (measured, B200): the cosine falls from 0.9999 at the output head to ~0.92 at the
    # 8^3 level for ANY bf16 evaluation of this network -- LeakyReLU / max-pool decisions of near-tie pre-activations
    # flip under a 2^-9 perturbation (SURVEY.md section 4 saw the same between fp32 and fp64 at the 3e-3 level).
    # The bar is therefore the yardstick, tensor by tensor, with no escape clause, plus absolute floors.
    for n1, cos, e, cos_y, e_y in rows:
        assert e <= 1.05 * e_y + 2e-3, (n1, e, e_y)
        assert cos >= cos_y - 2e-3 and cos >= 0.9, (n1, cos, cos_y)
    assert rows[-1][0].endswith("final_conv.weight") and rows[-1][1] > 0.9999 and rows[-1][2] < 5e-3
    flat = lambda m: torch.cat([p.grad.flatten().float() for p in m.parameters() if p.grad is not None and p.ndim > 1])
    f64, fg, fa = flat(o64), flat(g), flat(oa)
    cos_all = torch.nn.functional.cosine_similarity(fg, f64, dim=0).item()
    cos_all_y = torch.nn.functional.cosine_similarity(fa, f64, dim=0).item()
    print(f"  whole generator: cosine {cos_all:.5f} (autocast {cos_all_y:.5f}), rel-L2 {rel_l2(fg, f64):.3e} "
          f"(autocast {rel_l2(fa, f64):.3e}); worst tensor {worst[0]} cos {worst[1]:.4f}")
    assert cos_all >= cos_all_y - 1e-3 and cos_all >= 0.95, (cos_all, cos_all_y)
    assert rel_l2(fg, f64) <= 1.05 * rel_l2(fa, f64) + 2e-3, (rel_l2(fg, f64), rel_l2(fa, f64), worst)
