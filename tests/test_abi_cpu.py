"""CPU checks of the drop-in boundary: libb2seg.so loads, exports every symbol include/b2seg.h declares, and the ctypes
struct mirrors have exactly the C sizes.  No compute call is made (there is no GPU here); a compute entry point
must fail loudly rather than fall back to the CPU."""
import ctypes
import os
import re

import pytest

from b2seg import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b2seg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b2seg_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "run `python __graft_entry__.py` first"
    lib = ctypes.CDLL(L.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(L.EXPORTED) == declared, (set(declared) ^ set(L.EXPORTED))


def test_ctypes_mirrors_have_c_sizes():
    lib = ctypes.CDLL(L.LIB_PATH)
    for op, desc in L.OP_DESC.items():
        assert lib.b2seg_sizeof_desc(op) == ctypes.sizeof(desc), (op, desc.__name__)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    d = L.AdamDesc()
    rc = lib.b2seg_adam(ctypes.byref(d), None)
    assert rc == -2 and b"no CPU fallback" in lib.b2seg_last_error()
    from b2seg.models2d import unet_model_builder
    import numpy as np
    m = unet_model_builder("UNet", 32, 32, 8, 2, train_mode="from_scratch").ResNet50()
    with pytest.raises(L.B2SegError):
        m.predict(np.zeros((1, 32, 32, 3), np.float32))


def test_builder_api_errors_match_reference():
    from b2seg.models1d import UNet
    from b2seg.models2d import unet_model_builder
    with pytest.raises(ValueError, match="Train Mode"):
        unet_model_builder("UNet", 32, 32, 8, 2, train_mode="nope")
    with pytest.raises(ValueError, match="cannot be less than 1"):
        unet_model_builder("UNet", 32, 32, 8, 0, train_mode="from_scratch")
    with pytest.raises(ValueError, match="from 1 to 5"):
        unet_model_builder("UNet", 32, 32, 8, 6)
    with pytest.raises(ValueError, match="Please Check the Values"):
        unet_model_builder("UNet", 0, 32, 8, 2, train_mode="from_scratch").ResNet50()
    with pytest.raises(ValueError, match="Please Check the Values"):
        UNet(0, 2, 1, 8, 3).UNet()
    with pytest.raises(NameError):
        unet_model_builder("MultiResUNet", 32, 32, 8, 2, lstm=1, train_mode="from_scratch").build_graph()


def test_binding_refuses_a_library_of_another_abi_version(monkeypatch):
    """a stale libb2seg.so must fail loudly at load time, not with a missing symbol in the middle of a step"""
    from b2seg import _lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "ABI_VERSION", L.ABI_VERSION + 1)
    with pytest.raises(L.B2SegError, match="rebuild"):
        L.load()
    monkeypatch.setattr(L, "ABI_VERSION", L.ABI_VERSION - 1)
    assert L.load().b2seg_version() == L.ABI_VERSION
