"""GPU parity of the kernels either side of the network path: the multi-tensor AdamW (N3)
against torch.optim.AdamW, and the de-normalised NIfTI-order volume writer (N4) against NumPy."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed):
    torch.manual_seed(seed)
    shapes = [(1,), (3,), (5, 7), (32, 24, 3, 3, 3), (4097,), (8192,), (12289,), (64, 64, 2, 2, 2)] + [(17, 3)] * 60
    ps = [torch.randn(s, device=DEV) for s in shapes]
    # one parameter that is a 4-byte-aligned (not 16-byte-aligned) view: exercises the scalar path
    base = torch.randn(1001, device=DEV)
    ps.append(base[1:])
    return ps


@pytest.mark.parametrize("wd,lr", [(1e-2, 1e-3), (0.0, 3e-2)])
def test_adamw_matches_torch(wd, lr):
    import unet_bssfp_b200 as ub
    ours = [torch.nn.Parameter(p.clone()) for p in _params(0)]
    ours[-1] = torch.nn.Parameter(_params(0)[-1])                 # keep the unaligned view un-cloned
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    assert ours[-1].data_ptr() % 16 != 0
    oa = ub.FusedAdamW(ours, lr=lr, weight_decay=wd)
    ob = torch.optim.AdamW(theirs, lr=lr, weight_decay=wd, foreach=False, fused=False)
    for step in range(6):
        torch.manual_seed(100 + step)
        for i, (a, b) in enumerate(zip(ours, theirs)):
            if step == 2 and i == 3:
                a.grad = b.grad = None             # skipped this step; its own step count lags from here on (as in torch)
                continue
            gr = torch.randn_like(a) * (10.0 ** ((i % 5) - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    for a, b in zip(ours, theirs):
        torch.testing.assert_close(a, b, rtol=2e-6, atol=1e-7)


def test_adamw_matches_torch_five_steps_and_state_dict():
    import unet_bssfp_b200 as ub
    src = _params(1)
    ours = [torch.nn.Parameter(p.clone()) for p in src]
    theirs = [torch.nn.Parameter(p.clone()) for p in src]
    oa = ub.FusedAdamW(ours)
    ob = torch.optim.AdamW(theirs, foreach=False, fused=False)
    for step in range(5):
        torch.manual_seed(200 + step)
        for a, b in zip(ours, theirs):
            gr = torch.randn_like(a)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    for a, b in zip(ours, theirs):
        torch.testing.assert_close(a, b, rtol=2e-6, atol=1e-7)
    sa, sb = oa.state_dict(), ob.state_dict()
    assert set(sa["state"][0]) == set(sb["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    torch.testing.assert_close(sa["state"][3]["exp_avg_sq"], sb["state"][3]["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    assert float(sa["state"][0]["step"]) == float(sb["state"][0]["step"]) == 5.0
    # torch's optimizer resumes from our state and vice versa
    ob2 = torch.optim.AdamW(theirs, foreach=False, fused=False)
    ob2.load_state_dict(sa)
    oa2 = ub.FusedAdamW(ours)
    oa2.load_state_dict(sb)
    for a, b in zip(ours, theirs):
        gr = torch.randn_like(a)
        a.grad, b.grad = gr.clone(), gr.clone()
    oa2.step(); ob2.step()
    for a, b in zip(ours, theirs):
        torch.testing.assert_close(a, b, rtol=3e-6, atol=1e-7)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cpu = torch.nn.Parameter(torch.zeros(4))
        cpu.grad = torch.ones(4)
        ub.FusedAdamW([cpu]).step()


def test_adamw_grad_scale_equals_prescaled_gradients():
    import unet_bssfp_b200 as ub
    a = torch.nn.Parameter(torch.randn(5000, device=DEV))
    b = torch.nn.Parameter(a.detach().clone())
    gr = torch.randn(5000, device=DEV)
    oa, ob = ub.FusedAdamW([a]), ub.FusedAdamW([b])
    oa.grad_scale = 0.125
    a.grad, b.grad = gr.clone(), gr * 0.125
    oa.step(); ob.step()
    torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("shape", [(6, 20, 24, 16), (1, 33, 5, 70), (6, 160, 192, 160)])
def test_denorm_to_nifti_order(shape):
    from unet_bssfp_b200 import nifti
    torch.manual_seed(0)
    vol = torch.rand(shape, device=DEV)
    lo, hi = -0.0031, 0.0042
    block = nifti.volume_to_nifti_order(vol, (lo, hi))
    assert tuple(block.shape) == (shape[0], shape[3], shape[2], shape[1])
    want = (vol.cpu().numpy().astype(np.float64) * abs(hi - lo) + lo).astype(np.float32)     # (C,X,Y,Z)
    np.testing.assert_array_equal(block.cpu().numpy(), want.transpose(0, 3, 2, 1))
    plain = nifti.volume_to_nifti_order(vol)
    np.testing.assert_array_equal(plain.cpu().numpy(), vol.cpu().numpy().transpose(0, 3, 2, 1))


def test_save_prediction_roundtrip(tmp_path):
    from unet_bssfp_b200 import nifti
    torch.manual_seed(1)
    vol = torch.rand(6, 32, 48, 16, device=DEV)
    for name in ("pred.nii", "pred.nii.gz"):
        path = os.path.join(tmp_path, name)
        nifti.save_prediction(path, vol[None], denorm=(0.5, 2.5))
        data, affine = nifti.read_nifti(path)
        assert data.shape == (32, 48, 16, 6)                        # channel-last, as np.moveaxis(volume, 0, -1)
        want = (np.moveaxis(vol.cpu().numpy(), 0, -1).astype(np.float64) * 2.0 + 0.5).astype(np.float32)
        np.testing.assert_array_equal(data, want)
        np.testing.assert_array_equal(affine, np.eye(4))
