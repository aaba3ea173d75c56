"""Generate tests/golden/data_reference.npz by executing the reference's OWN, UNMODIFIED loaders.

    python tests/golden/make_golden_data.py        (needs /root/reference; not run on the GPU box)

What executes: /root/reference/data/image_pair_dataloader.py (ImagePairDataLoader.sample_generator / _get_image_at /
map_fn), data/keypoint_dataloader.py (KeypointDataLoader._get_image_at / map_fn) and utils/data.py, with the installed
Pillow doing the pixel work.  Two stand-ins, both stated here:
  * ``tensorflow`` is a namespace with float32 / int16 / name_scope (the loaders touch nothing else outside get_dataset);
    tf.data's cast of the generator's float64 arrays to the declared tf.float32 is done explicitly before map_fn;
  * the reference pins Pillow 6.2.0 whose ``Image.resize`` defaults to NEAREST; the installed Pillow (>= 7) defaults to
    BICUBIC, so ``Image.Image.resize`` is wrapped to restore the pinned default.  ``rotate`` still defaults to NEAREST.
The data set is synthetic: seeded low-frequency patterns + noise encoded as JPEG here; the encoded BYTES are stored in the
fixture so that every consumer decodes identical files.  Nothing from /root/reference is copied into the repo.
"""
import contextlib
import hashlib
import io
import os
import random
import sys
import tempfile
import types

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("KP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_reference.npz")

VIDEOS = [("frames/0001", (192, 144)), ("frames/0002", (144, 192)), ("frames/0003", (160, 160)),
          ("frames/0004", (256, 144)), ("frames/0005", (157, 131)), ("frames/0006", (131, 157))]
FRAMES_PER_VIDEO = [13, 12, 14, 12, 15, 12]
N_TRAIN, N_FULL = 48, 10
SEED = 1234


def synth_jpeg(rng, w, h, t):
    yy, xx = np.mgrid[0:h, 0:w]
    a = rng.uniform(0.02, 0.09, (3, 2))
    ph = rng.uniform(0, 6.28, 3) + 0.15 * t
    img = np.stack([127 + 110 * np.sin(a[c, 0] * xx + ph[c]) * np.cos(a[c, 1] * yy - ph[c]) for c in range(3)], -1)
    img = np.clip(img + rng.normal(0, 10, img.shape), 0, 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=80)
    return buf.getvalue()


def make_dataset_bytes():
    rng = np.random.default_rng(SEED)
    files = {}
    for (name, (w, h)), n in zip(VIDEOS, FRAMES_PER_VIDEO):
        for t in range(n):
            files["%s/%06d.jpg" % (name, t + 1)] = synth_jpeg(rng, w, h, t)
    return files


def write_dataset(root, files):
    """Lay the stored JPEG bytes out as the data_dir the loaders expect (<video>/<%06d>.jpg + <subset>_set.txt)."""
    for rel, data in files.items():
        os.makedirs(os.path.join(root, os.path.dirname(rel)), exist_ok=True)
        with open(os.path.join(root, rel), "wb") as fh:
            fh.write(data)
    names = sorted({os.path.dirname(r) for r in files})
    for subset in ("train", "test"):
        with open(os.path.join(root, subset + "_set.txt"), "w") as fh:
            fh.write("\n".join("%s %d" % (n, i % 3) for i, n in enumerate(names)))


def files_from_fixture(fx):
    return {str(k): bytes(fx["jpeg_" + str(i)].tobytes()) for i, k in enumerate(fx["jpeg_names"])}


def to_u8(x01):
    u = np.rint(np.asarray(x01) * 255.0).astype(np.uint8)
    assert np.array_equal(u / 255.0, np.asarray(x01)), "not an exact /255 image"
    return u


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int16 = "float32", "int16"
    tf.name_scope = lambda name: contextlib.nullcontext()
    sys.modules["tensorflow"] = tf
    sys.path.insert(0, REF)
    _resize = Image.Image.resize

    def resize_pillow_6_2(self, size, resample=Image.NEAREST, *a, **k):
        return _resize(self, tuple(size), resample, *a, **k)

    Image.Image.resize = resize_pillow_6_2
    from data.image_pair_dataloader import ImagePairDataLoader      # noqa: E402  (reference)
    from data.keypoint_dataloader import KeypointDataLoader         # noqa: E402  (reference)

    files = make_dataset_bytes()
    fx = {"jpeg_names": np.array(sorted(files))}
    for i, k in enumerate(sorted(files)):
        fx["jpeg_" + str(i)] = np.frombuffer(files[k], np.uint8)

    def mapped(loader, sample, keys):
        f32 = dict(sample)
        f32.update({k: np.asarray(sample[k]).astype(np.float32) for k in keys})   # from_generator's cast to tf.float32
        return loader.map_fn(f32)

    with tempfile.TemporaryDirectory() as root:
        write_dataset(root, files)
        # training stream: random order, full augmentation
        np.random.seed(SEED)
        random.seed(SEED)
        ld = ImagePairDataLoader(root, "train", random_order=True, randomness=True)
        full, hashes = [], []

        def repeated():                      # tf.data's repeat() re-invokes the generator when it is exhausted
            while True:
                yield from ld.sample_generator()

        gen = repeated()
        for i in range(N_TRAIN):
            s = next(gen)
            m = mapped(ld, s, ("image", "future_image"))
            pair = np.stack([to_u8(s["image"]), to_u8(s["future_image"])])
            assert m["image"].dtype == np.float32
            hashes.append(sha(np.stack([m["image"], m["future_image"]])))
            if i < N_FULL:
                full.append(pair)
        fx["train_u8"] = np.stack(full)
        fx["train_sha256_f32"] = np.array(hashes)
        # evaluation stream: fixed order, centre crop
        ld = ImagePairDataLoader(root, "test", random_order=False, randomness=False)
        ev = [np.stack([to_u8(s["image"]), to_u8(s["future_image"])]) for s in ld.sample_generator()]
        fx["eval_u8"] = np.stack(ev)
        # keypoint loader: whole videos, zero-padded to 663 frames
        kl = KeypointDataLoader(root, "test")
        lens, idxs, hk, firsts = [], [], [], []
        for v in range(kl.length()):
            s = kl._get_image_at(v)
            m = mapped(kl, s, ("image",))
            assert m["image"].shape == (663, 128, 128, 3)
            u = to_u8(s["image"])
            lens.append(s["len"]); idxs.append(s["idx"])
            hk.append(sha(m["image"]))
            firsts.append(u[[0, s["len"] - 1, s["len"]]])        # first, last real and first padding frame
        fx["kp_len"], fx["kp_idx"] = np.array(lens), np.array(idxs)
        fx["kp_sha256_f32"] = np.array(hk)
        fx["kp_u8"] = np.stack(firsts)
    np.savez_compressed(OUT, **fx)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(files), "jpegs,", sum(map(len, files.values())), "jpeg bytes")


if __name__ == "__main__":
    main()
