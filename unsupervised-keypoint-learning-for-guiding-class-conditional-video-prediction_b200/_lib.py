"""ctypes binding of libkp_b200.so (the C ABI in include/kp_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# KP_B200_LIB selects a debug variant of the SAME library (build.py --trace); there is still no fallback of any kind
LIB_PATH = os.environ.get("KP_B200_LIB") or os.path.join(_HERE, "libkp_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "kp_b200.h")

_lib = None

c_f32p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_ll = ctypes.c_longlong
c_vp = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES). Kept in lock-step with include/kp_b200.h;
# tests/test_abi.py parses the header and checks every declared symbol is bound here and exported.
SIGNATURES = {
    "kp_abi_version": [],
    "kp_last_error": [],
    "kp_launch_count": [],
    "kp_softargmax_render_fwd": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_float, c_vp],
    "kp_softargmax_render_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_vp, c_vp, c_vp],
    "kp_render_fwd": [c_vp, c_int, c_int, c_int, c_int, c_float, c_vp, c_vp],
    "kp_render_bwd": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_float, c_vp, c_vp],
    "kp_render_colorize_fwd": [c_vp, c_vp, c_int, c_int, c_int, c_int, c_float, c_vp, c_vp],
    "kp_colorize_fwd": [c_vp, c_vp, c_ll, c_int, c_vp, c_vp],
    "kp_tapconv_bf16": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "kp_tapconv_wgrad_bf16": [c_vp, c_vp, c_vp, c_vp, c_vp],
    "kp_image_prep": [c_vp, c_ll, c_vp, c_vp, c_vp, c_vp, c_vp],
    "kp_image_prep_bwd": [c_vp, c_ll, c_vp, c_vp, c_int, c_vp, c_vp],
    "kp_bn_finalize": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, ctypes.c_double, c_float, c_float, c_vp, c_vp, c_vp, c_vp,
                       c_vp, c_vp, c_vp],
    "kp_bn_stats_apply": [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_double, c_float, c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                          c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp],
    "kp_upsample2x_bwd": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "kp_bn_act_apply": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "kp_bn_act_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                      c_vp, c_vp, c_int, c_int, c_vp],
    "kp_act_mask_bwd": [c_vp, c_vp, c_float, c_ll, c_vp, c_vp],
    "kp_maxpool2x2_fwd": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "kp_maxpool2x2_bwd": [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "kp_mask_compose_fwd": [c_vp, c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_vp],
    "kp_mask_compose_bwd": [c_vp, c_vp, c_vp, c_ll, c_vp, c_vp],
    "kp_pack_channels": [c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_vp, c_vp],
    "kp_unpack_channels": [c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_int, c_vp],
    "kp_l1_pair_fwd_bwd": [c_vp, c_vp, c_ll, c_float, c_vp, c_vp, c_vp],
    "kp_bce_logits_fwd_bwd": [c_vp, c_int, c_float, c_float, c_vp, c_vp, c_vp],
    "kp_adam_tf": [c_vp, c_vp, c_vp, c_vp, c_ll, c_float, c_float, c_float, c_float, c_int, c_float, c_vp, c_vp],
    "kp_channel_sum": [c_vp, c_ll, c_int, c_vp, c_vp],
    "kp_channel_sumsq": [c_vp, c_ll, c_int, c_vp, c_vp],
    "kp_conv1x1_f32": [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp, c_vp],
    "kp_pack_weights": [c_vp, c_vp, c_vp, c_vp, c_vp],
    "kp_pack_job_blocks": [c_vp],
    "kp_pack_weights_batch": [c_vp, c_int, c_int, c_vp],
    "kp_image_prep_unrolled": [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    "kp_image_prep_unrolled_bwd": [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp],
    "kp_augment_plan_host": [c_vp, c_ll, c_int, c_int, c_int, c_int, ctypes.c_double, ctypes.c_double, c_int, c_int, c_int,
                             ctypes.c_double],
    "kp_augment_plan_zero_host": [c_vp],
    "kp_augment_plan_batch_host": [c_vp, c_int] + [c_vp] * 12,
    "kp_augment_frames": [c_vp, c_vp, c_int, c_vp, c_vp],
    "kp_host_register": [c_vp, ctypes.c_ulonglong],
    "kp_host_unregister": [c_vp],
}
_RESTYPES = {"kp_last_error": ctypes.c_char_p, "kp_launch_count": ctypes.c_ulonglong}


class KpError(RuntimeError):
    pass


def header_symbols(path=HEADER_PATH):
    """Function names declared in the public header."""
    with open(path) as fh:
        txt = fh.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kp_[a-z0-9_]+)\s*\(", txt)))


def load():
    """Load (once) and return the ctypes handle. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KpError(
            "libkp_b200.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU/PyTorch fallback for the B200 path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().kp_last_error()
        raise (ValueError if rc == -1 else KpError)(
            "%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def call(name, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)
