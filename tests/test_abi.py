"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares."""
import ctypes
import os
import subprocess

import pytest


def test_header_symbols_all_exported_and_bound(lib_built):
    lib_mod = lib_built._lib
    declared = lib_mod.header_symbols()
    assert "kp_softargmax_render_fwd" in declared and len(declared) >= 8
    handle = lib_mod.load()
    for name in declared:
        assert hasattr(handle, name), "symbol %s declared in include/kp_b200.h is not exported" % name
        assert name in lib_mod.SIGNATURES, "symbol %s has no ctypes signature" % name
    assert set(lib_mod.SIGNATURES) == set(declared)
    assert handle.kp_abi_version() >= 1


def test_library_contains_sm100a_tma_code(lib_built):
    out = subprocess.run(["cuobjdump", "-lelf", lib_built._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", lib_built._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass  # TMA bulk copy feeds the fused keypoint kernel


def test_argument_validation_without_gpu(lib_built):
    """Invalid arguments are rejected before any CUDA call, so this runs on a CPU-only box."""
    L = lib_built._lib
    h = L.load()
    rc = h.kp_softargmax_render_fwd(None, 4, 128, 128, 40, None, None, None, None, 0, 0, 14.3, None)
    assert rc == -1
    assert b"logits" in h.kp_last_error()
    rc = h.kp_render_fwd(None, -1, 40, 32, 32, 14.3, None, None)
    assert rc == -1
    # B == 0 is a valid empty call
    assert h.kp_render_fwd(None, 0, 40, 32, 32, 14.3, None, None) == 0
    with pytest.raises(ValueError):
        L.call("kp_render_fwd", None, 2, 0, 32, 32, 14.3, None, None)


def test_python_mirror_rejects_cpu_tensors(lib_built):
    import torch
    from kp_b200 import model_utils
    with pytest.raises(ValueError):
        model_utils.get_gaussian_maps(torch.zeros(1, 40, 2), [32, 32])
    with pytest.raises(ValueError):
        model_utils.get_coord(torch.zeros(1, 8, 8, 4), 2, 8)
