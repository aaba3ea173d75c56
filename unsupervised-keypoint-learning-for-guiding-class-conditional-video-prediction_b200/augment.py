"""Device-side frame augmentation: ctypes mirror of ``kp_frame_plan`` and the ``kp_augment_*`` entry points
(include/kp_b200.h, csrc/augment.cu).  Replaces the per-frame Pillow chain of the reference's loaders
(data/image_pair_dataloader.py:95-165, data/keypoint_dataloader.py:66-82, utils/data.py:8-35)."""
import ctypes

import numpy as np
import torch

from . import _lib

IMAGE_SIZE = 128
NO_FILTER = -1


class FramePlan(ctypes.Structure):
    _fields_ = [("src_offset", ctypes.c_longlong), ("src_w", ctypes.c_int), ("src_h", ctypes.c_int),
                ("rotate", ctypes.c_int), ("a", ctypes.c_int * 6), ("filter_id", ctypes.c_int), ("zero", ctypes.c_int),
                ("factor", ctypes.c_float), ("xtab", ctypes.c_short * IMAGE_SIZE), ("ytab", ctypes.c_short * IMAGE_SIZE)]


PLAN_BYTES = ctypes.sizeof(FramePlan)
assert PLAN_BYTES == 568


class PlanTable:
    """Host array of frame plans living in a (pinned, when CUDA is there) uint8 tensor, so that it goes to the device with
    one asynchronous copy."""

    def __init__(self, n_frames, pin=None):
        pin = torch.cuda.is_available() if pin is None else pin
        self.n = int(n_frames)
        self.host = torch.zeros(max(self.n, 1) * PLAN_BYTES, dtype=torch.uint8, pin_memory=pin)
        self._base = self.host.data_ptr()

    def _ptr(self, i):
        if not 0 <= i < self.n:
            raise IndexError(i)
        return ctypes.c_void_p(self._base + i * PLAN_BYTES)

    def set(self, i, src_offset, src_w, src_h, resize_w, resize_h, crop_left, crop_top, angle=0, flip=0,
            filter_id=NO_FILTER, factor=0.0):
        _lib.call("kp_augment_plan_host", self._ptr(i), int(src_offset), int(src_w), int(src_h), int(resize_w),
                  int(resize_h), float(crop_left), float(crop_top), int(angle), int(bool(flip)), int(filter_id), float(factor))

    def set_zero(self, i):
        _lib.call("kp_augment_plan_zero_host", self._ptr(i))

    def set_batch(self, requests, offsets):
        """All plans in ONE library call.  requests: frame-request dicts (data/base_dataloader.py: ``size``, ``resize``,
        ``crop``, ``angle``, ``flip``, ``filter_id``, ``factor``, or ``zero``); offsets: byte offset of each frame."""
        n = len(requests)
        if n != self.n:
            raise ValueError("set_batch: %d requests for a table of %d plans" % (n, self.n))
        live = [(0, 1, 1, 1, 1, 0.0, 0.0, 0, 0, NO_FILTER, 0.0, 1) if r.get("zero") else
                (o, r["size"][0], r["size"][1], r["resize"][0], r["resize"][1], r["crop"][0], r["crop"][1], r["angle"],
                 int(bool(r["flip"])), r["filter_id"], r["factor"], 0) for r, o in zip(requests, offsets)]
        cols = list(zip(*live)) if live else [()] * 12
        kinds = (np.int64, np.int32, np.int32, np.int32, np.int32, np.float64, np.float64, np.int32, np.int32, np.int32,
                 np.float64, np.int32)
        arrs = [np.ascontiguousarray(c, dtype=k) for c, k in zip(cols, kinds)]
        _lib.call("kp_augment_plan_batch_host", ctypes.c_void_p(self._base), n, *[a.ctypes.data for a in arrs])

    def view(self, i):
        """The i-th plan as a ctypes structure (host memory of this table)."""
        return FramePlan.from_address(self._base + i * PLAN_BYTES)


def augment_frames(src, plans, n_frames, out=None, stream=None):
    """src uint8 CUDA tensor (decoded frames), plans uint8 CUDA tensor holding ``n_frames`` kp_frame_plan records ->
    float32 [n_frames, 128, 128, 3] in [-1, 1].  One kernel launch on ``stream`` (default: the current stream)."""
    if not (src.is_cuda and plans.is_cuda):
        raise ValueError("augment_frames: frames and plans must be CUDA tensors (there is no CPU path)")
    if src.dtype != torch.uint8 or plans.dtype != torch.uint8 or plans.numel() < n_frames * PLAN_BYTES:
        raise ValueError("augment_frames: uint8 source / plan buffers expected, plans too short for %d frames" % n_frames)
    if out is None:
        out = torch.empty((n_frames, IMAGE_SIZE, IMAGE_SIZE, 3), dtype=torch.float32, device=src.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
              and out.numel() == n_frames * IMAGE_SIZE * IMAGE_SIZE * 3):
        raise ValueError("augment_frames: out must be a contiguous float32 CUDA tensor of n_frames*128*128*3 elements")
    st = stream if stream is not None else torch.cuda.current_stream(src.device)
    _lib.call("kp_augment_frames", src.data_ptr(), plans.data_ptr(), int(n_frames), out.data_ptr(), st.cuda_stream)
    return out


def frames_to_buffer(frames):
    """Pack decoded uint8 [h, w, 3] arrays into one byte buffer; returns (numpy uint8 buffer, offsets)."""
    offs, total = [], 0
    for f in frames:
        offs.append(total)
        total += int(f.size)
    buf = np.empty(max(total, 1), np.uint8)
    for f, o in zip(frames, offs):
        buf[o:o + f.size] = np.ascontiguousarray(f, dtype=np.uint8).reshape(-1)
    return buf, offs
