"""ORACLE (test infrastructure only) -- NumPy restatement of the evaluation arithmetic on the hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

* ``rel_error_map``   ref:src/eval.py:154-166 (do_calc_diff_maps; file I/O stripped)
* ``dti_scalar_maps`` ref:src/eval.py:73-116 (do_calc_scalar_maps; the Python triple loop is batched through
                      ``np.linalg.eigh(..., 'U')``, which runs the same LAPACK routine per matrix)
* ``roi_error_avg``   ref:src/eval.py:217-258 (do_calc_error_avg; file-name parsing, pandas and NIfTI
                      I/O stripped -- the arithmetic of lines 238-249 is kept verbatim in meaning)
* ``invert_dwi_tensor_norm`` ref:src/eval.py:39-47 (do_invert_dwi_tensor_norm; min-max de-normalisation)

PARITY PIN: the reference has no tests or golden vectors of its own for these functions (SURVEY.md
section 4). Pinned by (1) ``tests/golden/golden_ref_v1.npz``: outputs of the reference's OWN
``do_calc_diff_maps / do_calc_error_avg / do_calc_scalar_maps / do_invert_dwi_tensor_norm``, executed
unmodified from /root/reference/src/eval.py by ``tests/golden/make_golden_ref.py`` (nibabel replaced by an
in-memory image store), checked in ``tests/test_reference_golden_cpu.py``; (2) hand-checkable cases in
tests/test_oracle_cpu.py (zeros -> inf -> 0, NaN kept, angular wrap-around).
"""
import numpy as np

ANGULAR_KINDS = ("azimuth", "inclination")


def rel_error_map(pred, target, kind="denorm"):
    """ref:src/eval.py:160-164. float64 like nibabel's get_fdata()."""
    pred = np.asarray(pred, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    if kind not in ANGULAR_KINDS:
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.abs(pred - target) / target
    diff = (pred - target) % 360
    return np.where(diff < 180, diff, 360 - diff)


def roi_error_avg(diff, mask, probseg):
    """ref:src/eval.py:238-249. diff (X,Y,Z[,C]), mask (X,Y,Z), probseg (X,Y,Z,R) -> errs [R][C]
    and the masked / inf-cleaned map that the reference writes back (line 251)."""
    diff_map = np.abs(np.asarray(diff, dtype=np.float64))
    diff_map = diff_map if diff_map.ndim == 4 else diff_map[..., np.newaxis]
    diff_map = diff_map.copy()
    probseg = np.asarray(probseg, dtype=np.float64)
    errs = np.zeros((probseg.shape[-1], diff_map.shape[-1]), dtype=np.float64)
    for i in range(diff_map.shape[-1]):
        diff_map[..., i] = np.where(mask > 0, diff_map[..., i], 0)
        diff_map[..., i] = np.where(diff_map[..., i] == np.inf, 0, diff_map[..., i])
        for roi_idx in range(probseg.shape[-1]):
            segmented = probseg[..., roi_idx] * diff_map[..., i]
            norm = probseg[..., roi_idx].sum()
            errs[roi_idx, i] = segmented.sum() / norm
    return errs, diff_map


def invert_dwi_tensor_norm(data, min_v, max_v):
    """ref:src/eval.py:39-47: data * |max - min| + min per channel, float64 like get_fdata()."""
    data = np.array(data, dtype=np.float64)
    for i in range(data.shape[-1]):
        data[..., i] = (data[..., i] * np.abs(max_v - min_v)) + min_v
    return data


def dti_scalar_maps(data, canonical_sign=False):
    """ref:src/eval.py:73-116. data (..., 6) -> dict(fa, md, ad, rd, azimuth, inclination, rgb) in float64.
    ``canonical_sign``: orient the principal eigenvector with v_z >= 0 (then v_y, then v_x) before taking
    the angles -- LAPACK leaves the sign implementation-defined; the reference keeps whatever it returns."""
    data = np.asarray(data, dtype=np.float64)
    d = np.zeros(data.shape[:-1] + (3, 3))
    d[..., 0, 0], d[..., 0, 1], d[..., 0, 2] = data[..., 0], data[..., 1], data[..., 2]
    d[..., 1, 0], d[..., 1, 1], d[..., 1, 2] = data[..., 1], data[..., 3], data[..., 4]
    d[..., 2, 0], d[..., 2, 1], d[..., 2, 2] = data[..., 2], data[..., 4], data[..., 5]
    eigvals, eigvecs = np.linalg.eigh(d, "U")
    ad = eigvals[..., 2]
    rd = (eigvals[..., 0] + eigvals[..., 1]) / 2
    md = eigvals.mean(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        var = np.sqrt(((eigvals - md[..., None]) ** 2).sum(-1))
        norm = np.sqrt((eigvals ** 2).sum(-1))
        fa = np.sqrt(1.5) * var / norm
    v = eigvecs[..., :, 2].copy()
    if canonical_sign:
        flip = (v[..., 2] < 0) | ((v[..., 2] == 0) & ((v[..., 1] < 0) | ((v[..., 1] == 0) & (v[..., 0] < 0))))
        v[flip] *= -1
    azimuth = 180 / np.pi * np.arctan2(v[..., 1], v[..., 0])
    azimuth = np.where(azimuth > 180, azimuth - 360, azimuth)
    r = np.sqrt((v ** 2).sum(-1))
    inclination = 180 / np.pi * np.arccos(np.clip(v[..., 2] / r, -1.0, 1.0))
    rgb = fa[..., None] * np.abs(v)
    return {"fa": fa, "md": md, "ad": ad, "rd": rd, "azimuth": azimuth, "inclination": inclination, "rgb": rgb}
