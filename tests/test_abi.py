"""The built library exports every entry point the header declares (no compute: runs without a GPU)."""

from __future__ import annotations

import ctypes

from minispark_b200 import native


def test_library_exports_every_declared_symbol():
    lib = native.load()
    declared = native.declared_symbols()
    assert len(declared) >= 35
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, missing
    assert set(native._SIGNATURES) == set(declared), set(native._SIGNATURES) ^ set(declared)
    assert lib.msc_abi_version() == native.K["MSC_ABI_VERSION"]


def test_struct_layouts_match_the_header():
    k = native.K
    assert ctypes.sizeof(native.ColBind) == 16
    expect = (8 + 4 + 4 + 16 * (k["MSC_VM_MAX_STAGED"] + k["MSC_VM_MAX_GATHER"]) + 4 + 4 + 4 * k["MSC_VM_MAX_CODE"]
              + 8 * k["MSC_VM_MAX_CONSTS"] + 4 + 4 + 8 * k["MSC_VM_MAX_LUTS"] + 4 + 4 + 4 * k["MSC_VM_MAX_CODE2"] + 8 + 8)  # + nrows_dev, want_jit + pad
    assert ctypes.sizeof(native.ScanDesc) == expect


def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.NativeError):
        native.Context(0)
