"""GPU: the input pipeline through the C ABI - byte-exact against the oracle and against what the
unmodified reference loaders produced (tests/golden/data_reference.npz)."""
import os
import random

import numpy as np
import pytest
import torch

import data_common as DC

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(240)]      # worker processes and pipes: fail, never hang

SIZES = [(320, 240), (240, 320), (200, 200), (480, 270), (157, 131), (131, 157), (129, 640)]


def _run(dev, frames, reqs):
    from kp_b200 import augment as A
    buf, offs = A.frames_to_buffer([f for f in frames if f is not None])
    table = A.PlanTable(len(reqs))
    k = 0
    for i, r in enumerate(reqs):
        if r.get("zero"):
            table.set_zero(i)
            continue
        h, w = frames[i].shape[:2]
        table.set(i, offs[k], w, h, r["resize"][0], r["resize"][1], r["crop"][0], r["crop"][1], r["angle"], r["flip"],
                  r["filter_id"], r["factor"])
        k += 1
    out = A.augment_frames(torch.from_numpy(buf).to(dev), table.host.to(dev), len(reqs))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_kernel_matches_oracle_bit_for_bit(cuda_dev):
    rng = np.random.default_rng(3)
    frames, reqs = DC.random_requests(rng, 132, SIZES)
    frames.append(None)
    reqs.append({"zero": True})
    got = _run(cuda_dev, frames, reqs)
    for i, (f, r) in enumerate(zip(frames, reqs)):
        want = DC.oracle_frame(f, r)
        assert np.array_equal(got[i], want), (i, r, int((got[i] != want).sum()))
    # and the host build of the same phases agrees with the device
    assert np.array_equal(got, DC.emulate(frames, reqs))


def test_every_enhancement_factor(cuda_dev):
    rng = np.random.default_rng(4)
    frames, reqs = DC.random_requests(rng, 2, SIZES[:2])
    f, base = frames[1], reqs[1]
    for fid, (lo, hi) in {6: (0, 50), 7: (7, 20), 8: (0, 50), 9: (7, 30)}.items():
        rs = [dict(base, filter_id=fid, factor=v * 0.1) for v in range(lo, hi + 1)]
        got = _run(cuda_dev, [f] * len(rs), rs)
        for g, r in zip(got, rs):
            assert np.array_equal(g, DC.oracle_frame(f, r)), r


@pytest.mark.parametrize("prefetch,decode", [(False, "thread"), (True, "process"), (False, "process")])
def test_pair_loader_reproduces_the_reference_stream(cuda_dev, tmp_path, prefetch, decode):
    """Seeded like the fixture generator, ImagePairDataLoader.get_dataset yields the reference's batches bit for bit."""
    from kp_b200 import data
    fx = DC.fixture()
    root = DC.lay_out_dataset(tmp_path)
    np.random.seed(DC.SEED)
    random.seed(DC.SEED)
    ld = data.ImagePairDataLoader(root, "train", random_order=True, randomness=True)
    ds = ld.get_dataset(batch_size=4, repeat=True, num_preprocess_threads=4, prefetch=prefetch, device=cuda_dev,
                        decode=decode)
    n, k = len(fx["train_sha256_f32"]), 0
    for batch in ds:
        assert batch["image"].shape == (4, 128, 128, 3) and batch["image"].dtype == torch.float32 and batch["image"].is_cuda
        torch.cuda.synchronize()
        im, fu = batch["image"].cpu().numpy(), batch["future_image"].cpu().numpy()
        for j in range(4):
            assert DC.sha(np.stack([im[j], fu[j]])) == str(fx["train_sha256_f32"][k]), k
            k += 1
        if k >= (n // 4) * 4:
            break
    assert k == (n // 4) * 4
    ld = data.ImagePairDataLoader(root, "test", random_order=False, randomness=False)
    from oracle import pil_ops as O
    ds.close()
    got = list(ld.get_dataset(batch_size=4, prefetch=prefetch, device=cuda_dev, decode=decode))
    assert [b["image"].shape[0] for b in got] == [4, 2]                 # ragged last batch
    torch.cuda.synchronize()
    im = np.concatenate([b["image"].cpu().numpy() for b in got])
    fu = np.concatenate([b["future_image"].cpu().numpy() for b in got])
    assert np.array_equal(np.stack([im, fu], 1), O.to_model_range(fx["eval_u8"]))


def test_keypoint_loader_reproduces_the_reference(cuda_dev, tmp_path):
    from kp_b200 import data
    fx = DC.fixture()
    root = DC.lay_out_dataset(tmp_path)
    kl = data.KeypointDataLoader(root, "test")
    for v, batch in enumerate(kl.get_dataset(batch_size=1, device=cuda_dev)):
        assert batch["image"].shape == (1, 663, 128, 128, 3)
        assert int(batch["len"][0]) == int(fx["kp_len"][v]) and int(batch["idx"][0]) == int(fx["kp_idx"][v])
        torch.cuda.synchronize()
        assert DC.sha(batch["image"][0].cpu().numpy()) == str(fx["kp_sha256_f32"][v]), v
    assert v == 5


def test_loader_feeds_the_train_step(cuda_dev, tmp_path):
    """The batches are what DetectorTranslatorModel.build consumes (the reference: train.py:36-43 -> model.build(inputs))."""
    from kp_b200 import data, models
    root = DC.lay_out_dataset(tmp_path)
    ld = data.ImagePairDataLoader(root, "train", random_order=True, randomness=True)
    it = iter(ld.get_dataset(batch_size=2, repeat=True, device=cuda_dev))
    cfg = {"paths": {"data_dir": "", "vggnet": None, "log_dir": str(tmp_path / "log")},
           "training": {"batch_size": 2, "lr": {"start_val": 1e-4, "step": 20000, "decay": 0.95}},
           "model": {"n_pts": 40, "n_action": 9, "cell_info": [1024, 1024], "vae_dim": 64}}
    model = models.DetectorTranslatorModel(cfg, is_training=True, device=cuda_dev)
    model.build(next(it))
    lD, lG, _, bs = model.test_step()
    assert np.isfinite(lD) and np.isfinite(lG) and bs == 2


def test_pseudo_labels_from_the_keypoint_loader(cuda_dev, tmp_path):
    """make_pseudo_labels.py end to end: KeypointDataLoader -> KeypointModel -> pseudo_labels/{idx:04d}.npy, sharded over two
    (simulated) ranks; each file holds the detector's key points of exactly the `len` real frames of its video."""
    from kp_b200 import data, models
    fx = DC.fixture()
    root = DC.lay_out_dataset(tmp_path / "data")
    cfg = {"paths": {"data_dir": root, "vggnet": None, "log_dir": str(tmp_path / "log")},
           "model": {"n_pts": 40, "n_action": 9, "cell_info": [1024, 1024], "vae_dim": 64}}
    model = models.KeypointModel(cfg, device=cuda_dev)
    kl = data.KeypointDataLoader(root, "test")
    out = str(tmp_path / "pseudo_labels")
    files = []
    for rank in (0, 1):
        files += model.write_pseudo_labels(kl, out, rank=rank, world=2)
    assert sorted(np.unique(files)) == sorted(files) and len(files) == 6
    videos = list(kl.get_dataset(1, device=cuda_dev))
    for v in videos:
        idx, n = int(v["idx"][0]), int(v["len"][0])
        arr = np.load(os.path.join(out, "%04d.npy" % idx))          # what data/sequence_dataloader.py:101 reads
        assert arr.dtype == np.float32 and arr.shape == (n, 40, 2) and n == int(fx["kp_len"][list(fx["kp_idx"]).index(idx)])
        want = model.detect(v["image"][0, :n].contiguous()).cpu().numpy()
        assert np.array_equal(arr, want)
        assert np.all(np.abs(arr) <= 1.0)


def test_abandoned_datasets_release_their_staging_memory(cuda_dev, tmp_path):
    """A data set whose iterator is dropped mid-stream (its last reference then dies on the batch-building thread) must still
    un-register its page-locked staging mapping: a later mapping at the same address registers cleanly and no CUDA error is
    left behind."""
    import gc
    import time
    from kp_b200 import data
    root = DC.lay_out_dataset(tmp_path)
    ld = data.ImagePairDataLoader(root, "train", random_order=True, randomness=True)
    for _ in range(4):
        it = iter(ld.get_dataset(batch_size=2, repeat=True, num_preprocess_threads=2, device=cuda_dev))
        b = next(it)
        assert b["image"].shape == (2, 128, 128, 3)
        del it, b
        gc.collect()
        time.sleep(0.3)                     # the producer notices the stop flag within 0.1 s and lets go of the data set
    torch.cuda.synchronize()
    assert float(torch.empty(4, device=cuda_dev).fill_(1).sum()) == 4.0
