"""Shared helpers of the input-pipeline tests (CPU and GPU): the fixture, the data set laid out from its stored JPEG bytes,
and the host-emulated kernel (tests/host_emu)."""
import ctypes
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FIXTURE = os.path.join(HERE, "golden", "data_reference.npz")
SEED = 1234          # make_golden_data.SEED


def fixture():
    return np.load(FIXTURE)


def lay_out_dataset(root):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_data as G
    G.write_dataset(str(root), G.files_from_fixture(fixture()))
    return str(root)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def emu_lib():
    """g++ build of the kernel phases for the host (test infrastructure)."""
    src = os.path.join(HERE, "host_emu", "augment_host.cpp")
    out = os.path.join(HERE, "host_emu", "_build", "libaugment_emu.so")
    pkg = os.path.join(ROOT, "unsupervised-keypoint-learning-for-guiding-class-conditional-video-prediction_b200")
    hdr = os.path.join(pkg, "csrc", "augment_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                               "-I" + os.path.join(pkg, "csrc"), "-o", out, src])
    lib = ctypes.CDLL(out)
    lib.kp_emu_augment_frames.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return lib


def emulate(frames, requests):
    """frames: decoded uint8 arrays; requests: frame-request dicts (data/base_dataloader.py) -> float32 [n,128,128,3] through
    the C plan builder of the product library + the host-compiled kernel phases."""
    from kp_b200 import augment as A
    buf, offs = A.frames_to_buffer([f for f in frames if f is not None])
    table = A.PlanTable(len(requests), pin=False)
    k = 0
    for i, r in enumerate(requests):
        if r.get("zero"):
            table.set_zero(i)
            continue
        h, w = frames[i].shape[:2]
        table.set(i, offs[k], w, h, r["resize"][0], r["resize"][1], r["crop"][0], r["crop"][1], r["angle"], r["flip"],
                  r["filter_id"], r["factor"])
        k += 1
    out = np.empty((len(requests), 128, 128, 3), np.float32)
    emu_lib().kp_emu_augment_frames(buf.ctypes.data, table.host.data_ptr(), len(requests), out.ctypes.data)
    return out


def oracle_frame(img, r):
    """The oracle's value of one frame request (numpy restatement of the Pillow chain)."""
    from oracle import pil_ops as O
    if r.get("zero"):
        return O.to_model_range(np.zeros((128, 128, 3), np.uint8))
    x = img
    if r["angle"]:
        x = O.rotate_nearest(x, r["angle"])
    x = O.resize_nearest(x, r["resize"])
    left, top = r["crop"]
    x = O.crop(x, (left, top, left + 128, top + 128))
    if r["flip"]:
        x = x[:, ::-1].copy()
    if r["filter_id"] >= 0:
        x = O.kernel_filter(x, r["filter_id"]) if r["filter_id"] <= 5 else O.enhance(x, r["filter_id"], r["factor"])
    return O.to_model_range(x)


def random_requests(rng, n, sizes):
    """Seeded frames + requests covering every filter, rotation, flip and out-of-frame crops."""
    frames, reqs = [], []
    for i in range(n):
        w, h = sizes[i % len(sizes)]
        yy, xx = np.mgrid[0:h, 0:w]
        if i % 3 == 0:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        else:
            img = np.stack([127 + 120 * np.sin(xx / (5. + c) + i) * np.cos(yy / (7. - c) - i) for c in range(3)], -1)
            img = np.clip(img + rng.normal(0, 8, img.shape), 0, 255).astype(np.uint8)
        ratio = min(w, h) / 128.0
        W, H = int(w / ratio), int(h / ratio)
        fid = i % 11 - 1
        factor = {6: rng.integers(0, 51), 7: rng.integers(7, 21), 8: rng.integers(0, 51), 9: rng.integers(7, 31)}.get(fid, 0) * 0.1
        if i % 7 == 6:
            crop = (W / 2.0 - 64 + 0.5 * (i % 2), -3 if i % 2 else 0)        # float box, partly outside
        else:
            crop = (int(rng.integers(0, W - 128 + 1)), int(rng.integers(0, H - 128 + 1)))
        frames.append(img)
        reqs.append({"resize": (W, H), "crop": crop, "angle": int(rng.integers(-10, 11)) if i % 4 else 0,
                     "flip": int(rng.integers(0, 2)), "filter_id": fid, "factor": float(factor)})
    return frames, reqs
