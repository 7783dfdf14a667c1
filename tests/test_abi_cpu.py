"""CPU tests of the boundary: the shared library loads, exports every symbol include/ub_api.h
declares, the ctypes table covers the header, and argument validation fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ub_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ub_[a-z0-9_]+)\s*\(", src)))


def test_build_and_symbols():
    import __graft_entry__ as ge
    ge.build()
    from unet_bssfp_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ub_api.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.load().ub_version() >= 100


def test_bad_arguments_are_errors():
    from unet_bssfp_b200 import _lib
    lib = _lib.load()
    d = _lib.ConvDesc(0, 1, 8, 16, 8, 24, 24, 0, 0, 32, 32)      # c0p not a multiple of 32
    assert lib.ub_packed_weight_elems(C.byref(d), 0) < 0
    assert b"multiples of 32" in lib.ub_last_error()
    d = _lib.ConvDesc(2, 1, 7, 16, 8, 32, 32, 0, 0, 32, 32)      # odd size for the stride-2 conv
    assert lib.ub_conv_num_tiles(C.byref(d)) < 0
    d = _lib.ConvDesc(0, 2, 8, 32, 16, 32, 32, 0, 0, 64, 64)
    assert lib.ub_conv_num_tiles(C.byref(d)) == 2 * 2 * 2 * 2
    assert lib.ub_packed_weight_elems(C.byref(d), 0) == 27 * 64 * 32
    assert lib.ub_l1_fwd(None, None, 10, None, None, None) < 0
    with pytest.raises(RuntimeError):
        _lib.check(-1, "test")


def test_cpu_tensors_are_rejected():
    import torch
    from unet_bssfp_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pack_ncdhw(torch.zeros(1, 6, 2, 2, 2))


def test_module_tree_matches_oracle():
    import unet_bssfp_b200 as ub
    from oracle import model_oracle as O
    for mod in ("bssfp", "t1w"):
        g, og = ub.Generator(mod), O.Generator(mod)
        d, od = ub.Discriminator(mod), O.Discriminator(mod)
        assert {k: tuple(v.shape) for k, v in g.state_dict().items()} == {k: tuple(v.shape) for k, v in og.state_dict().items()}
        assert {k: tuple(v.shape) for k, v in d.state_dict().items()} == {k: tuple(v.shape) for k, v in od.state_dict().items()}
        assert [n for n, _ in g.named_parameters()] == [n for n, _ in og.named_parameters()]
        g.load_state_dict(og.state_dict()); od.load_state_dict(d.state_dict())
    with pytest.raises(NotImplementedError):
        ub.DownSampleConv(8, 8, kernel=3, strides=1, padding=1)
