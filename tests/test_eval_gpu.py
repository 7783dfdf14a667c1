"""GPU parity of the evaluation kernel (relative / angular error map + masked, probseg-weighted ROI
means) against the NumPy oracle of ref:eval.py:154-166,217-258 and the goldens. Tolerance 1e-4
(north_star: "voxel error maps agree to 1e-4"); the kernel computes in fp32, the reference in fp64."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as E
from tests.golden.make_golden import synth_eval_volumes

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def _run(pred, tgt, mask, probseg, angular):
    from unet_bssfp_b200 import ops
    dev = "cuda"
    diff, sums, norms = ops.relerr_map_reduce(torch.from_numpy(pred).to(dev), torch.from_numpy(tgt).to(dev),
                                              None if mask is None else torch.from_numpy(mask).to(dev),
                                              None if probseg is None else torch.from_numpy(probseg).to(dev),
                                              angular=angular)
    torch.cuda.synchronize()
    errs = None if sums is None else (sums / norms[:, None]).cpu().numpy()
    return diff.cpu().numpy(), errs


def test_relative_error_matches_oracle_and_golden():
    pred, tgt, mask, probseg = synth_eval_volumes()
    diff, errs = _run(pred, tgt, mask, probseg, False)
    ref = E.rel_error_map(pred, tgt)
    ref_errs, _ = E.roi_error_avg(ref, mask, probseg)
    np.testing.assert_allclose(diff, ref, rtol=1e-4, atol=1e-6, equal_nan=True)
    assert np.isinf(diff[0, 0, 0]).all() and np.isnan(diff[1, 1, 1, 0])           # x/0 -> inf, 0/0 -> nan as numpy
    np.testing.assert_allclose(errs, ref_errs, rtol=1e-4, equal_nan=True)         # inf zeroed, NaN kept
    np.testing.assert_allclose(diff, GOLD["eval_diff_rel"], rtol=1e-4, atol=1e-6, equal_nan=True)
    np.testing.assert_allclose(errs, GOLD["eval_errs_rel"], rtol=1e-4, equal_nan=True)


def test_angular_error_matches_oracle_and_golden():
    pred, tgt, mask, probseg = synth_eval_volumes()
    ang_p = (pred * 720.0 - 180.0).astype(np.float32)[..., :1].copy()
    ang_t = (tgt * 360.0).astype(np.float32)[..., :1].copy()
    diff, errs = _run(ang_p, ang_t, mask, probseg, True)
    ref = E.rel_error_map(ang_p, ang_t, kind="azimuth")
    ref_errs, _ = E.roi_error_avg(ref, mask, probseg)
    np.testing.assert_allclose(diff, ref, rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(errs, ref_errs, rtol=1e-4)
    np.testing.assert_allclose(errs, GOLD["eval_errs_ang"], rtol=1e-4)
    assert diff.min() >= 0 and diff.max() <= 180.0


def test_full_size_volume_properties():
    """config 5 size (160x192x160x6): linear in probseg, zero outside the mask, map-only mode."""
    rng = np.random.default_rng(0)
    shape = (160, 192, 160)
    tgt = rng.uniform(0.05, 1.0, size=shape + (6,)).astype(np.float32)
    pred = (tgt * rng.uniform(0.9, 1.1, size=tgt.shape)).astype(np.float32)
    mask = (rng.uniform(size=shape) > 0.3).astype(np.uint8)
    ps = rng.dirichlet([1, 1, 1], size=shape).astype(np.float32)
    diff, errs = _run(pred, tgt, mask, ps, False)
    diff2, errs2 = _run(pred, tgt, mask, (2.0 * ps).astype(np.float32), False)
    np.testing.assert_allclose(errs, errs2, rtol=1e-6)                     # ratio is scale-free in probseg
    np.testing.assert_array_equal(diff, diff2)
    _, errs0 = _run(pred, tgt, np.zeros(shape, np.uint8), ps, False)
    assert np.all(errs0 == 0)
    sub = (slice(0, 16), slice(0, 16), slice(0, 16))
    np.testing.assert_allclose(diff[sub], E.rel_error_map(pred[sub], tgt[sub]), rtol=1e-4)
    ref_errs, _ = E.roi_error_avg(E.rel_error_map(pred, tgt), mask, ps)
    np.testing.assert_allclose(errs, ref_errs, rtol=1e-4)
    d_only, none = _run(pred, tgt, None, None, False)
    assert none is None
    np.testing.assert_array_equal(d_only, diff)


def _synth_dti(shape, seed=0):
    """SPD-ish diffusion tensors with a clear principal axis plus isotropic / degenerate / zero voxels."""
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    q, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    lam = np.sort(rng.uniform(0.1e-3, 3e-3, size=(n, 3)), axis=1)
    lam[:, 2] *= rng.uniform(1.05, 3.0, size=n)
    d = np.einsum("nij,nj,nkj->nik", q, lam, q)
    t6 = np.stack([d[:, 0, 0], d[:, 0, 1], d[:, 0, 2], d[:, 1, 1], d[:, 1, 2], d[:, 2, 2]], -1).astype(np.float32)
    t6[0] = 0.0                                   # zero tensor: FA = 0/0 = NaN, axis e_z
    t6[1] = [2e-3, 0, 0, 2e-3, 0, 2e-3]           # isotropic: FA = 0
    t6[2] = [1e-3, 0, 0, 2e-3, 0, 3e-3]           # already diagonal
    t6[3] = [3e-3, 0, 0, 2e-3, 0, 1e-3]           # principal axis = x
    return t6.reshape(tuple(shape) + (6,))


def test_dti_scalar_maps_match_oracle():
    """ub_dti_scalar_maps vs the NumPy/LAPACK oracle of ref:eval.py:73-116 (1e-4). Sign-free maps are
    compared as is; azimuth / inclination against the oracle with the same principal-axis orientation
    (v_z >= 0), since LAPACK's eigenvector sign is implementation-defined."""
    from unet_bssfp_b200 import ops
    t6 = _synth_dti((24, 20, 16))
    got = {k: v.cpu().numpy() for k, v in ops.dti_scalar_maps(torch.from_numpy(t6).cuda()).items()}
    ref = E.dti_scalar_maps(t6, canonical_sign=True)
    for k in ("md", "ad", "rd"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-4, atol=1e-9)
    flat_fa, ref_fa = got["fa"].reshape(-1), ref["fa"].reshape(-1)
    assert np.isnan(flat_fa[0]) and np.isnan(ref_fa[0])                      # zero tensor: 0/0 kept
    np.testing.assert_allclose(flat_fa[1:], ref_fa[1:], rtol=1e-4, atol=1e-6)
    ok = np.ones(flat_fa.shape, bool)
    ok[:3] = False                                                           # zero / isotropic / tie: axis not unique
    np.testing.assert_allclose(got["rgb"].reshape(-1, 3)[ok], ref["rgb"].reshape(-1, 3)[ok], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got["inclination"].reshape(-1)[ok], ref["inclination"].reshape(-1)[ok], rtol=1e-4, atol=2e-3)
    da = np.abs(got["azimuth"].reshape(-1)[ok] - ref["azimuth"].reshape(-1)[ok])
    da = np.minimum(da, 360 - da)                                            # +-180 is the same direction at v_z = 0
    horizontal = ref["inclination"].reshape(-1)[ok] > 89.999
    assert (da[~horizontal] < 2e-3).all()
    assert (np.minimum(da[horizontal], np.abs(da[horizontal] - 180)) < 2e-3).all()
    # against the raw (reference-sign) oracle: every voxel is either identical or the antipode
    raw = E.dti_scalar_maps(t6)
    inc_g, inc_r = got["inclination"].reshape(-1)[ok], raw["inclination"].reshape(-1)[ok]
    assert (np.minimum(np.abs(inc_g - inc_r), np.abs(inc_g - (180 - inc_r))) < 2e-3).all()
    # hand cases
    assert abs(flat_fa[1]) < 1e-6 and abs(got["azimuth"].reshape(-1)[3]) < 1e-4 and abs(got["inclination"].reshape(-1)[3] - 90) < 1e-4
    assert abs(got["ad"].reshape(-1)[2] - 3e-3) < 1e-9 and abs(got["inclination"].reshape(-1)[2]) < 1e-4


def test_dti_scalar_maps_full_volume_invariants():
    """160x192x160: MD = trace / 3, AD >= MD >= RD, 0 <= FA <= 1, rotation-free identities."""
    from unet_bssfp_b200 import ops
    t6 = _synth_dti((160, 192, 160), seed=1)
    m = ops.dti_scalar_maps(torch.from_numpy(t6).cuda())
    torch.cuda.synchronize()
    tr = torch.from_numpy((t6[..., 0] + t6[..., 3] + t6[..., 5]) / 3).cuda()
    assert ((m["md"] - tr).abs() <= 1e-6 * tr.abs() + 1e-10).all()
    fa = m["fa"].reshape(-1)[1:]
    assert (fa >= -1e-6).all() and (fa <= 1 + 1e-6).all()
    assert (m["ad"] >= m["md"] - 1e-9).all() and (m["md"] >= m["rd"] - 1e-9).all()
    assert (m["inclination"] >= 0).all() and (m["inclination"] <= 90.0001).all()
    assert ((m["ad"] + 2 * m["rd"]) / 3 - m["md"]).abs().max().item() < 1e-8
