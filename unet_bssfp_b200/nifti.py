"""Prediction volumes on disk: what ``save_predicitions`` + ``do_invert_dwi_tensor_norm``
of the reference produce (ref:src/model.py:335-357, ref:src/eval.py:39-47), without nibabel.

The reference moves the channel axis last (``np.moveaxis(volume, 0, -1)``), wraps the array in
``nib.Nifti1Image(array, np.eye(4))`` and saves ``.nii.gz``; evaluation then reloads it as float64, applies
``v * |max - min| + min`` and saves again under the original (float32) header. Here ``ub_denorm_to_nifti``
writes the de-normalised volume straight in NIfTI storage order on the device (one pass), and this module adds
the 348-byte NIfTI-1 header (single-file ``n+1`` flavour, identity affine as sform like nibabel's default for
an affine-only image) around it. ``read_nifti`` is the inverse, used by the tests.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np
import torch

from . import ops

__all__ = ["volume_to_nifti_order", "write_nifti", "read_nifti", "save_prediction"]

_DT_FLOAT32 = 16


def volume_to_nifti_order(volume: torch.Tensor, denorm=None) -> torch.Tensor:
    """(C,X,Y,Z) fp32 CUDA volume (module layout) -> fp32 tensor of shape (C,Z,Y,X): the data block of the
    NIfTI file of the channel-last array (X,Y,Z,C). ``denorm = (min_v, max_v)`` applies ref:src/eval.py:44."""
    scale, offset = 1.0, 0.0
    if denorm is not None:
        min_v, max_v = float(denorm[0]), float(denorm[1])
        scale, offset = abs(max_v - min_v), min_v
    return ops.denorm_to_nifti(volume, scale, offset)


def _header(shape, affine):
    dim = [len(shape)] + list(shape) + [1] * (7 - len(shape))
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)                       # sizeof_hdr
    struct.pack_into("<8h", hdr, 40, *dim)                    # dim[8]
    struct.pack_into("<h", hdr, 70, _DT_FLOAT32)              # datatype
    struct.pack_into("<h", hdr, 72, 32)                       # bitpix
    struct.pack_into("<8f", hdr, 76, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)   # pixdim (qfac = 1)
    struct.pack_into("<f", hdr, 108, 352.0)                   # vox_offset
    struct.pack_into("<f", hdr, 112, float("nan"))            # scl_slope: NaN = "no scaling" (nibabel's default)
    struct.pack_into("<f", hdr, 116, float("nan"))            # scl_inter
    struct.pack_into("<h", hdr, 252, 0)                       # qform_code: unknown
    struct.pack_into("<h", hdr, 254, 2)                       # sform_code: aligned
    struct.pack_into("<3f", hdr, 256, 0.0, 0.0, 0.0)          # quatern_b, c, d (identity rotation)
    struct.pack_into("<3f", hdr, 268, *[float(affine[i][3]) for i in range(3)])   # qoffset
    for r in range(3):
        struct.pack_into("<4f", hdr, 280 + 16 * r, *[float(v) for v in affine[r]])  # srow_x / y / z
    hdr[344:348] = b"n+1\x00"
    return bytes(hdr) + b"\x00\x00\x00\x00"                   # no header extensions


def write_nifti(path, data_nifti_order: np.ndarray, shape, affine=None):
    """``data_nifti_order``: float32 values already in storage order (first axis of ``shape`` fastest)."""
    affine = np.eye(4) if affine is None else np.asarray(affine)
    data = np.ascontiguousarray(data_nifti_order, dtype="<f4")
    if data.size != int(np.prod(shape)):
        raise ValueError("data does not match the declared shape")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(_header(tuple(int(s) for s in shape), affine))
        f.write(data.tobytes())


def read_nifti(path):
    """-> (array in logical index order (X,Y,Z[,C]) float32, affine (4,4) from the sform)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if struct.unpack_from("<i", raw, 0)[0] != 348 or raw[344:347] != b"n+1":
        raise ValueError("not a single-file little-endian NIfTI-1 image")
    dim = struct.unpack_from("<8h", raw, 40)
    shape = tuple(dim[1:1 + dim[0]])
    if struct.unpack_from("<h", raw, 70)[0] != _DT_FLOAT32:
        raise ValueError("only float32 images are written by this package")
    off = int(struct.unpack_from("<f", raw, 108)[0])
    data = np.frombuffer(raw, dtype="<f4", count=int(np.prod(shape)), offset=off).reshape(shape, order="F")
    affine = np.eye(4)
    for r in range(3):
        affine[r] = struct.unpack_from("<4f", raw, 280 + 16 * r)
    return data, affine


def save_prediction(path, volume: torch.Tensor, denorm=None):
    """One (C,X,Y,Z) or (1,C,X,Y,Z) prediction volume -> NIfTI-1 file of the channel-last array (X,Y,Z,C)."""
    vol = volume if volume.dim() == 4 else volume[0]
    c, x, y, z = vol.shape
    block = volume_to_nifti_order(vol, denorm).cpu().numpy()
    write_nifti(path, block, (x, y, z, c))
