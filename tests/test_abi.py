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


def test_argument_validation_of_conv_and_bn_entry_points(lib_built):
    """NULL pointers / bad shapes of the convolution-side entry points are rejected with rc -1 and a message naming the
    argument, before any CUDA call (so this runs without a GPU)."""
    import ctypes as C
    L = lib_built._lib
    h = L.load()
    from kp_b200 import tapconv as tc
    # kp_bn_stats_apply: NULL statistics
    rc = h.kp_bn_stats_apply(None, None, None, None, None, 16.0, 1e-5, 0.999, None, None, None, None, None, None, None, 1, 0,
                             1, 4, 4, 16, None, 1, None)
    assert rc == -1 and b"kp_bn_stats_apply" in h.kp_last_error()
    # ... and a batch that does not split into the requested number of statistics segments
    rc = h.kp_bn_stats_apply(None, None, None, None, None, 16.0, 1e-5, 0.999, None, None, None, None, None, None, None, 1, 0,
                             3, 4, 4, 16, None, 2, None)
    assert rc == -1 and b"segments" in h.kp_last_error()
    # kp_upsample2x_bwd: bad shape, then NULL tensors
    assert h.kp_upsample2x_bwd(None, 0, 4, 4, 16, None, None) == -1
    assert h.kp_upsample2x_bwd(None, 1, 4, 4, 16, None, None) == -1
    # kp_pack_weights_batch: NULL job table
    assert h.kp_pack_weights_batch(None, 1, 1, None) == -1
    # kp_pack_job_blocks: malformed descriptor -> 0 blocks, well-formed forward descriptor -> tiles of 64 x 32
    d = tc.PackDesc()
    assert h.kp_pack_job_blocks(C.byref(d)) == 0
    plan, _ = tc.plan_conv_fwd([(1, 8, 8, 64)], 3, 1, 0, 128)
    d = tc.pack_desc(plan, (3, 3, 64, 128))
    assert h.kp_pack_job_blocks(C.byref(d)) == 9 * ((64 + 63) // 64) * ((128 + 31) // 32)
    # kp_tapconv_bf16: a descriptor with an unsupported channel block
    td = plan.desc()
    td.CB = 24
    ptrs = (C.c_void_p * tc.KP_MAX_MAPS)()
    rc = h.kp_tapconv_bf16(C.byref(td), ptrs, None, None, None, None, None, None)
    assert rc == -1
