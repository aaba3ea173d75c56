"""CPU: the input pipeline without a GPU.

1. the oracle (oracle/pil_ops.py) is pinned bit-for-bit against the installed Pillow called with Pillow 6.2.0's defaults;
2. the oracle reproduces the fixtures the UNMODIFIED reference loaders produced (tests/golden/make_golden_data.py) when it
   is driven by the product loaders' sample descriptions - i.e. the product's random draws, file choice and geometry are
   the reference's;
3. the product's C plan builder (kp_augment_plan_host) equals the oracle's tables / fixed-point coefficients;
4. the kernel's arithmetic (csrc/augment_core.cuh compiled for the host) equals the oracle on every filter / rotation /
   flip / crop case.  The GPU run of the same kernel is tests/test_data_gpu.py.
"""
import random

import numpy as np
import pytest
from PIL import Image, ImageEnhance, ImageFilter

import data_common as DC
from oracle import pil_ops as O

SIZES = [(320, 240), (240, 320), (200, 200), (480, 270), (157, 131), (131, 157), (129, 640)]


def _img(rng, w, h, smooth):
    if not smooth:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.stack([127 + 120 * np.sin(xx / 9. + c) * np.cos(yy / 7. - c) for c in range(3)], -1) + rng.normal(0, 12, (h, w, 3))
    return np.clip(a, 0, 255).astype(np.uint8)


def test_oracle_geometry_matches_pillow():
    rng = np.random.default_rng(0)
    for (w, h) in SIZES + [(128, 128), (1920, 1080)]:
        img = _img(rng, w, h, False)
        pil = Image.fromarray(img)
        for ang in range(-10, 11):
            assert np.array_equal(np.asarray(pil.rotate(ang)), O.rotate_nearest(img, ang)), (w, h, ang)
        W, H, _ = O.resized_size(w, h)
        assert np.array_equal(np.asarray(pil.resize([W, H], Image.NEAREST)), O.resize_nearest(img, (W, H))), (w, h)
    img = _img(rng, 171, 128, False)
    pil = Image.fromarray(img)
    for box in [(21.5, 0, 149.5, 128), (20.5, 0, 148.5, 128), (-10, 0, 118, 128), (60, 0, 188, 128), (0, -3.5, 128, 124.5)]:
        assert np.array_equal(np.asarray(pil.crop(box)), O.crop(img, box)), box


def test_oracle_filters_match_pillow():
    rng = np.random.default_rng(1)
    F = [ImageFilter.DETAIL, ImageFilter.EDGE_ENHANCE, ImageFilter.SMOOTH, ImageFilter.SMOOTH_MORE,
         ImageFilter.EDGE_ENHANCE_MORE, ImageFilter.BLUR]
    E = {6: ImageEnhance.Sharpness, 7: ImageEnhance.Brightness, 8: ImageEnhance.Color, 9: ImageEnhance.Contrast}
    for smooth in (False, True):
        for _ in range(3):
            img = _img(rng, 128, 128, smooth)
            pil = Image.fromarray(img)
            for fid in range(6):
                assert np.array_equal(np.asarray(pil.filter(F[fid])), O.kernel_filter(img, fid)), fid
        img = _img(rng, 128, 128, smooth)
        pil = Image.fromarray(img)
        for fid, (lo, hi) in {6: (0, 50), 7: (7, 20), 8: (0, 50), 9: (7, 30)}.items():
            for r in range(lo, hi + 1):       # every factor apply_random_filter can draw
                assert np.array_equal(np.asarray(E[fid](pil).enhance(r * 0.1)), O.enhance(img, fid, r * 0.1)), (fid, r)


def test_model_range_known_answers():
    t = O.to_model_range(np.array([0, 255, 128, 1], np.uint8))
    assert t.dtype == np.float32 and t[0] == -1.0 and t[1] == 1.0
    assert t[2] == np.float32(np.float32(128 / 255.0) * np.float32(2) - np.float32(1))


def test_plan_builder_matches_oracle(lib_built):
    from kp_b200 import augment as A
    T = A.PlanTable(1, pin=False)
    rng = np.random.default_rng(2)
    for (w, h) in SIZES + [(1920, 1080), (3001, 2000)]:
        W, H, _ = O.resized_size(w, h)
        xt, yt = O.scale_table(w, W), O.scale_table(h, H)
        for ang in list(range(-12, 13)) + [45, 90, 135, 180, 270, 359, -90]:
            flip = int(rng.integers(0, 2))
            left, top = int(rng.integers(-2, W - 128 + 3)), int(rng.integers(-2, H - 128 + 3))
            T.set(0, 77, w, h, W, H, left, top, ang, flip, 5, 0.3)
            p = T.view(0)
            m = O.rotate_matrix(w, h, ang)
            assert (p.rotate == 0) if m is None else (p.rotate == 1 and tuple(p.a) == O.affine_fixed_coeffs(m)), (w, h, ang)
            cols = [127 - x for x in range(128)] if flip else list(range(128))
            ex = [xt[left + x] if 0 <= left + x < W else -1 for x in cols]
            ey = [yt[top + y] if 0 <= top + y < H else -1 for y in range(128)]
            assert list(p.xtab) == ex and list(p.ytab) == ey, (w, h, ang, flip, left, top)
            assert p.src_offset == 77 and p.filter_id == 5 and p.factor == np.float32(0.3)
        # the crop box goes through int(round()): half to even
        for left in (20.5, 21.5, 22.5, -0.5):
            T.set(0, 0, w, h, W, H, left, 0)
            x0 = O.crop_box((left, 0, left + 128, 128))[0]
            assert list(T.view(0).xtab) == [xt[x0 + x] if 0 <= x0 + x < W else -1 for x in range(128)]


def test_batched_plan_builder_equals_single_plans(lib_built):
    from kp_b200 import augment as A
    rng = np.random.default_rng(9)
    frames, reqs = DC.random_requests(rng, 23, SIZES)
    reqs = [dict(r, size=(f.shape[1], f.shape[0])) for f, r in zip(frames, reqs)]
    reqs.insert(5, {"zero": True})
    offs = list(range(0, 1000 * len(reqs), 1000))
    one, many = A.PlanTable(len(reqs), pin=False), A.PlanTable(len(reqs), pin=False)
    many.set_batch(reqs, offs)
    for i, r in enumerate(reqs):
        if r.get("zero"):
            one.set_zero(i)
        else:
            one.set(i, offs[i], r["size"][0], r["size"][1], r["resize"][0], r["resize"][1], r["crop"][0], r["crop"][1], r["angle"],
                    r["flip"], r["filter_id"], r["factor"])
    assert bytes(one.host.numpy()) == bytes(many.host.numpy())
    with pytest.raises(ValueError):
        many.set_batch(reqs[:-1], offs[:-1])


def test_plan_builder_rejects_bad_arguments(lib_built):
    L = lib_built._lib
    h = L.load()
    from kp_b200 import augment as A
    T = A.PlanTable(1, pin=False)
    assert h.kp_augment_plan_host(None, 0, 10, 10, 128, 128, 0.0, 0.0, 0, 0, -1, 0.0) == -1
    assert h.kp_augment_plan_host(T._ptr(0), 0, 0, 10, 128, 128, 0.0, 0.0, 0, 0, -1, 0.0) == -1
    assert h.kp_augment_plan_host(T._ptr(0), 0, 10, 10, 128, 128, 0.0, 0.0, 0, 0, 10, 0.0) == -1
    assert b"filter_id" in h.kp_last_error()
    assert h.kp_augment_plan_host(T._ptr(0), 0, 40000, 10, 128, 128, 0.0, 0.0, 0, 0, -1, 0.0) == -2
    assert h.kp_augment_frames(None, None, -1, None, None) == -1
    assert h.kp_augment_frames(None, None, 0, None, None) == 0           # empty batch
    assert h.kp_augment_frames(None, None, 4, None, None) == -1
    assert h.kp_host_register(None, 16) == -1 and h.kp_host_unregister(None) == -1
    assert h.kp_augment_plan_batch_host(None, 0, *([None] * 12)) == 0 and h.kp_augment_plan_batch_host(None, 2, *([None] * 12)) == -1
    import torch
    with pytest.raises(ValueError):
        A.augment_frames(torch.zeros(16, dtype=torch.uint8), torch.zeros(568, dtype=torch.uint8), 1)


def test_kernel_arithmetic_on_the_host_matches_oracle(lib_built):
    """csrc/augment_core.cuh (the kernel's phases) compiled with g++ and driven by the product's plan builder."""
    rng = np.random.default_rng(3)
    frames, reqs = DC.random_requests(rng, 66, SIZES)
    got = DC.emulate(frames, reqs)
    for i, (f, r) in enumerate(zip(frames, reqs)):
        want = DC.oracle_frame(f, r)
        assert np.array_equal(got[i], want), (i, r, int((got[i] != want).sum()))
    # every enhancement factor the reference can draw, on one frame
    f = frames[1]
    for fid, (lo, hi) in {6: (0, 50), 7: (7, 20), 8: (0, 50), 9: (7, 30)}.items():
        rs = [dict(reqs[1], filter_id=fid, factor=v * 0.1) for v in range(lo, hi + 1)]
        got = DC.emulate([f] * len(rs), rs)
        for g, r in zip(got, rs):
            assert np.array_equal(g, DC.oracle_frame(f, r)), r
    z = DC.emulate([None], [{"zero": True}])
    assert np.all(z == -1.0)


def _decode(path):
    return np.asarray(Image.open(path).convert("RGB"))


def test_pair_loader_descriptions_reproduce_the_reference(lib_built, tmp_path):
    """Seeded like make_golden_data.py, the product loader must choose the reference's files and draw its parameters: the
    oracle applied to its descriptions equals what the unmodified reference produced."""
    from kp_b200 import data
    fx = DC.fixture()
    root = DC.lay_out_dataset(tmp_path)
    np.random.seed(DC.SEED)
    random.seed(DC.SEED)
    ld = data.ImagePairDataLoader(root, "train", random_order=True, randomness=True)
    assert ld.length() == 6 and ld.get_sample_shape()["image"] == [128, 128, 3]
    n, seen_filters = len(fx["train_sha256_f32"]), set()
    k = 0
    while k < n:
        for s in ld.sample_generator():
            if k == n:
                break
            reqs = s["frames"]["image"] + s["frames"]["future_image"]
            seen_filters.add(reqs[0]["filter_id"])
            pair = np.stack([DC.oracle_frame(_decode(r["path"]), r) for r in reqs])
            assert DC.sha(pair) == str(fx["train_sha256_f32"][k]), (k, reqs)
            if k < len(fx["train_u8"]):
                assert np.array_equal(pair, O.to_model_range(fx["train_u8"][k]))
            k += 1
    assert seen_filters == set(range(10))
    ld = data.ImagePairDataLoader(root, "test", random_order=False, randomness=False)
    for k, s in enumerate(ld.sample_generator()):
        reqs = s["frames"]["image"] + s["frames"]["future_image"]
        assert reqs[0]["path"].endswith("000001.jpg") and reqs[1]["path"].endswith("000011.jpg")
        pair = np.stack([DC.oracle_frame(_decode(r["path"]), r) for r in reqs])
        assert np.array_equal(pair, O.to_model_range(fx["eval_u8"][k])), k


def test_keypoint_loader_descriptions_reproduce_the_reference(lib_built, tmp_path):
    from kp_b200 import data
    fx = DC.fixture()
    root = DC.lay_out_dataset(tmp_path)
    kl = data.KeypointDataLoader(root, "test")
    for v, s in enumerate(kl.sample_generator()):
        reqs = s["frames"]["image"]
        assert len(reqs) == 663 and s["extra"]["len"] == int(fx["kp_len"][v]) and s["extra"]["idx"] == int(fx["kp_idx"][v])
        n = s["extra"]["len"]
        real = np.stack([DC.oracle_frame(_decode(r["path"]), r) for r in reqs[:n]])
        assert np.array_equal(real[[0, n - 1]], O.to_model_range(fx["kp_u8"][v][:2]))
        assert all(r.get("zero") for r in reqs[n:]) and np.all(fx["kp_u8"][v][2] == 0)
        # the kernel's arithmetic (host build) over the whole zero-padded video against the reference's hash
        got = DC.emulate([_decode(r["path"]) if not r.get("zero") else None for r in reqs], reqs)
        assert np.array_equal(got[:n], real)
        assert DC.sha(got) == str(fx["kp_sha256_f32"][v])


def test_decode_workers_write_identical_bytes_and_surface_errors(lib_built, tmp_path):
    """The worker processes of the loader (data/_kp_decode_worker.py) decode into a shared staging file exactly what PIL
    decodes in-process; a failing file raises in the parent."""
    import mmap
    from kp_b200.data import base_dataloader as BD
    root = DC.lay_out_dataset(tmp_path)
    paths = [str(tmp_path / "frames" / "0001" / ("%06d.jpg" % i)) for i in (1, 2, 3)] + \
            [str(tmp_path / "frames" / "0002" / "000001.jpg")]
    sizes = [BD.jpeg_size(p) for p in paths]
    offs = np.cumsum([0] + [w * h * 3 for w, h in sizes])
    stage = str(tmp_path / "stage_1")
    with open(stage, "w+b") as fh:
        fh.truncate(int(offs[-1]))
        view = np.frombuffer(mmap.mmap(fh.fileno(), int(offs[-1])), np.uint8)
    W = BD._DecodeWorkers(2)
    try:
        W.run([(stage, int(offs[i]), p) + tuple(sizes[i]) for i, p in enumerate(paths)])
        for i, p in enumerate(paths):
            w, h = sizes[i]
            assert np.array_equal(view[offs[i]:offs[i + 1]].reshape(h, w, 3), BD.decode_rgb(p))
        # two submissions outstanding (the loader keeps two batches in flight): answered in order
        view[:] = 0
        first = W.submit([(stage, int(offs[i]), paths[i]) + tuple(sizes[i]) for i in (0, 1)])
        second = W.submit([(stage, int(offs[i]), paths[i]) + tuple(sizes[i]) for i in (2, 3)])
        first(); second()
        for i, p in enumerate(paths):
            w, h = sizes[i]
            assert np.array_equal(view[offs[i]:offs[i + 1]].reshape(h, w, 3), BD.decode_rgb(p))
        with pytest.raises(RuntimeError, match="decode worker"):
            W.run([(stage, 0, str(tmp_path / "missing.jpg"), 4, 4)])
        with pytest.raises(RuntimeError, match="header said"):
            W.run([(stage, 0, paths[0], 5, 5)])
    finally:
        W.close()


def test_edge_geometries_kernel_arithmetic_vs_oracle_vs_pillow(lib_built):
    """Frames smaller than the crop (up-scaling), square, extreme aspect ratios, crops partly or wholly outside the frame:
    Pillow itself, the oracle and the kernel's phases (host build + the product's plan builder) agree byte for byte."""
    rng = np.random.default_rng(11)
    cases = [(100, 80), (80, 100), (128, 128), (64, 64), (640, 129), (129, 640), (1, 1), (3, 200), (255, 257)]
    frames, reqs, pil_out = [], [], []
    for i, (w, h) in enumerate(cases):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        W, H, _ = O.resized_size(w, h)
        for crop in [(0, 0), (W / 2.0 - 64, 0), (W - 100, H - 90), (-200, -200), (0.5, 1.5)]:
            ang = int(rng.integers(-10, 11))
            flip = int(rng.integers(0, 2))
            fid = int(rng.integers(-1, 10))
            fac = float(rng.integers(0, 51)) * 0.1
            frames.append(img)
            reqs.append({"resize": (W, H), "crop": crop, "angle": ang, "flip": flip, "filter_id": fid, "factor": fac})
            pil = Image.fromarray(img).rotate(ang).resize([W, H], Image.NEAREST)
            pil = pil.crop((crop[0], crop[1], crop[0] + 128, crop[1] + 128))
            if flip:
                pil = pil.transpose(Image.FLIP_LEFT_RIGHT)
            if 0 <= fid <= 5:
                pil = pil.filter([ImageFilter.DETAIL, ImageFilter.EDGE_ENHANCE, ImageFilter.SMOOTH, ImageFilter.SMOOTH_MORE,
                                  ImageFilter.EDGE_ENHANCE_MORE, ImageFilter.BLUR][fid])
            elif fid >= 6:
                pil = [ImageEnhance.Sharpness, ImageEnhance.Brightness, ImageEnhance.Color, ImageEnhance.Contrast][fid - 6](pil).enhance(fac)
            pil_out.append(O.to_model_range(np.asarray(pil)))
    got = DC.emulate(frames, reqs)
    for i, (f, r) in enumerate(zip(frames, reqs)):
        want = DC.oracle_frame(f, r)
        assert np.array_equal(want, pil_out[i]), ("oracle vs Pillow", f.shape, r)
        assert np.array_equal(got[i], want), ("kernel phases vs oracle", f.shape, r)


def test_keypoint_loader_shards_cover_the_video_list_once(lib_built, tmp_path):
    """make_pseudo_labels over N ranks: the loader's sample_range shards (dp.shard_range) are disjoint and complete."""
    from kp_b200 import data, dp
    root = DC.lay_out_dataset(tmp_path)
    kl = data.KeypointDataLoader(root, "test")
    everything = [s["extra"]["idx"] for s in kl.sample_generator()]
    for world in (1, 2, 3, 4, 7):
        seen = []
        for rank in range(world):
            lo, hi = dp.shard_range(kl.length(), rank, world)
            seen += [s["extra"]["idx"] for s in kl.sample_generator(lo, hi)]
        assert seen == everything, world
