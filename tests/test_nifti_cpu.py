"""CPU test of the NIfTI-1 container written around the device-produced data block: header
fields at the offsets the NIfTI-1 standard fixes, Fortran storage order, round trip through the reader."""
import gzip
import os
import struct

import numpy as np


def test_header_fields_and_roundtrip(tmp_path):
    from unet_bssfp_b200 import nifti
    rng = np.random.default_rng(0)
    arr = rng.normal(size=(5, 4, 3, 6)).astype(np.float32)                    # logical (X,Y,Z,C)
    block = arr.transpose(3, 2, 1, 0)                                         # storage order: X fastest
    for name in ("v.nii", "v.nii.gz"):
        path = os.path.join(tmp_path, name)
        nifti.write_nifti(path, block, arr.shape)
        raw = (gzip.open if name.endswith(".gz") else open)(path, "rb").read()
        assert len(raw) == 352 + arr.size * 4
        assert struct.unpack_from("<i", raw, 0)[0] == 348                      # sizeof_hdr
        assert struct.unpack_from("<8h", raw, 40) == (4, 5, 4, 3, 6, 1, 1, 1)  # dim
        assert struct.unpack_from("<h", raw, 70)[0] == 16 and struct.unpack_from("<h", raw, 72)[0] == 32
        assert struct.unpack_from("<f", raw, 108)[0] == 352.0                  # vox_offset
        assert struct.unpack_from("<h", raw, 254)[0] == 2                      # sform_code (aligned)
        assert struct.unpack_from("<4f", raw, 280) == (1.0, 0.0, 0.0, 0.0)     # srow_x of the identity affine
        assert raw[344:348] == b"n+1\x00"
        first = np.frombuffer(raw, "<f4", count=5, offset=352)
        np.testing.assert_array_equal(first, arr[:, 0, 0, 0])                  # x runs fastest on disk
        data, affine = nifti.read_nifti(path)
        np.testing.assert_array_equal(data, arr)
        np.testing.assert_array_equal(affine, np.eye(4))
