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
