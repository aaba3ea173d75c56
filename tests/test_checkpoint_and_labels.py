"""CPU: checkpoint interchange (TF variable names incl. Adam slots, models/base_model.py:74-92 of the reference) and the
pseudo-label writer (make_pseudo_labels.py:98-101 -> data/sequence_dataloader.py:101).  No kernels run here: the engine's
parameter store lives on the CPU and `detect` is a stub."""
import os

import numpy as np
import torch


class _TinyTrainer:
    """The attributes checkpoint.py touches, over a small parameter store."""

    def __init__(self, seed):
        from kp_b200 import engine as E
        from kp_b200.models.base_model import GlobalStep
        self.ctx = E.Context("cpu")
        self.ctx.G.add("translator/conv_1_0/conv2d/kernel", (3, 3, 8, 4))
        self.ctx.G.add("translator/conv_1_0/conv2d/bias", (4,))
        self.ctx.G.add("translator/b_norm_1_0/gamma", (4,))
        self.ctx.D.add("img_discr/conv_0/conv2d/kernel", (4, 4, 3, 5))
        self.ctx.S.add("translator/b_norm_1_0/moving_mean", (4,))
        self.ctx.V.add("vgg/conv1_1/filter", (3, 3, 3, 6))
        for g in (self.ctx.G, self.ctx.D, self.ctx.S, self.ctx.V):
            g.finalize()
        gen = torch.Generator().manual_seed(seed)
        for g in (self.ctx.G, self.ctx.D, self.ctx.S, self.ctx.V):
            g.data.copy_(torch.randn(g.data.shape, generator=gen))
        for g in (self.ctx.G, self.ctx.D):
            g.m.copy_(torch.randn(g.m.shape, generator=gen))
            g.v.copy_(torch.rand(g.v.shape, generator=gen))
        self.t_D, self.t_G = 7 + seed, 5 + seed
        self.global_step = GlobalStep(5 + seed)


def test_checkpoint_npz_round_trip_with_tf_names(tmp_path, lib_built):
    from kp_b200 import checkpoint
    a, b = _TinyTrainer(0), _TinyTrainer(1)
    v = checkpoint.export_variables(a)
    # TF names: variables, Adam slots of both optimizers, non-slot variables in optimizer creation order (D first), global_step
    for n in ("translator/conv_1_0/conv2d/kernel", "translator/conv_1_0/conv2d/kernel/Adam", "translator/conv_1_0/conv2d/kernel/Adam_1",
              "img_discr/conv_0/conv2d/kernel/Adam", "translator/b_norm_1_0/moving_mean", "vgg/conv1_1/filter",
              "beta1_power", "beta2_power", "beta1_power_1", "beta2_power_1", "global_step"):
        assert n in v, n
    assert "translator/b_norm_1_0/moving_mean/Adam" not in v and "vgg/conv1_1/filter/Adam" not in v
    assert v["translator/conv_1_0/conv2d/kernel"].shape == (3, 3, 8, 4) and v["translator/conv_1_0/conv2d/kernel"].dtype == np.float32
    assert abs(float(v["beta1_power"]) - 0.5 ** 8) < 1e-9 and abs(float(v["beta1_power_1"]) - 0.5 ** 6) < 1e-9    # beta^(t+1)
    path = checkpoint.save_npz(str(tmp_path / "model.ckpt-5.npz"), v)
    back = checkpoint.load_npz(path)
    assert set(back) == set(v)
    loaded = checkpoint.import_variables(b, back)
    assert "translator/conv_1_0/conv2d/kernel/Adam_1" in loaded and "global_step" in loaded
    def same(ga, gb, bufs=("data",)):       # per variable: the flat buffers carry alignment padding between tensors
        return all(torch.equal(ga._view(getattr(ga, bf), n), gb._view(getattr(gb, bf), n)) for n in ga.names() for bf in bufs)
    for ga, gb in ((a.ctx.G, b.ctx.G), (a.ctx.D, b.ctx.D), (a.ctx.S, b.ctx.S), (a.ctx.V, b.ctx.V)):
        assert same(ga, gb)
    for ga, gb in ((a.ctx.G, b.ctx.G), (a.ctx.D, b.ctx.D)):
        assert same(ga, gb, ("m", "v"))
    assert (b.t_D, b.t_G, b.global_step.value) == (a.t_D, a.t_G, a.global_step.value)
    # a dump of a TF checkpoint reader has no kp_b200/* keys: the step counters come from the beta powers; unknown names
    # are ignored, missing ones keep their values (reference restore semantics)
    c = _TinyTrainer(2)
    keep = c.ctx.D.data.clone()
    tf_like = {k: x for k, x in back.items() if not k.startswith("kp_b200/") and not k.startswith("img_discr")}
    tf_like["some/other/variable"] = np.zeros(3, np.float32)
    checkpoint.import_variables(c, tf_like)
    assert (c.t_D, c.t_G) == (a.t_D, a.t_G)
    assert same(c.ctx.G, a.ctx.G) and torch.equal(c.ctx.D.data, keep)


def test_vgg19_npy_loader(tmp_path, lib_built):
    """vgg.py:11 of the reference: np.load(vgg19.npy, encoding='latin1').item() -> {name: [W(3,3,Cin,Cout), b]}."""
    from kp_b200 import engine as E
    from kp_b200.networks import vgg
    rng = np.random.default_rng(0)
    data = {name: [rng.normal(size=(3, 3, ci, co)).astype(np.float32), rng.normal(size=(co,)).astype(np.float32)]
            for name, ci, co in vgg.VGG_LAYERS}
    path = str(tmp_path / "vgg19.npy")
    np.save(path, np.array(data, dtype=object), allow_pickle=True)
    ctx = E.Context("cpu")
    for name, ci, co in vgg.VGG_LAYERS:
        ctx.V.add("vgg/%s/filter" % name, (3, 3, ci, co))
        ctx.V.add("vgg/%s/biases" % name, (co,))
    for g in (ctx.G, ctx.D, ctx.S, ctx.V):
        g.finalize()
    loaded = vgg.load_npy_into(ctx, path)
    assert len(loaded) == 32
    for name, _, _ in vgg.VGG_LAYERS:
        assert np.array_equal(ctx.p("vgg/%s/filter" % name).numpy(), data[name][0])
        assert np.array_equal(ctx.p("vgg/%s/biases" % name).numpy(), data[name][1])


def test_pseudo_label_files_match_the_stage2_reader(tmp_path, lib_built):
    from kp_b200 import pseudo_labels
    rng = np.random.default_rng(3)
    T = 20      # the reference pads every clip to 663 frames; any T works
    videos = []
    for i, n in enumerate([20, 7, 13, 1, 16]):
        im = torch.zeros((1, T, 8, 8, 3))
        im[0, :n] = torch.from_numpy(rng.uniform(-1, 1, (n, 8, 8, 3)).astype(np.float32))
        videos.append({"image": im, "idx": torch.tensor([100 + i]), "len": torch.tensor([n])})

    def detect(frames):       # stand-in for KeypointModel.detect: [F,h,w,3] -> [F,40,2]
        m = frames.mean(dim=(1, 2))
        return torch.stack([m[:, :1].expand(-1, 40), m[:, 1:2].expand(-1, 40)], dim=-1).contiguous()
    files = []
    for rank in range(2):     # two ranks: disjoint contiguous shards, together every video exactly once
        files += pseudo_labels.write_pseudo_labels(detect, videos, str(tmp_path), rank=rank, world=2)
    assert sorted(os.path.basename(f) for f in files) == ["%04d.npy" % (100 + i) for i in range(5)]
    for v in videos:
        n, idx = int(v["len"][0]), int(v["idx"][0])
        # what data/sequence_dataloader.py:101 does: np.load(<...>/pseudo_labels/<video>.npy), then index by frame
        kp = np.load(os.path.join(str(tmp_path), "%04d.npy" % idx))
        assert kp.dtype == np.float32 and kp.shape == (n, 40, 2)
        assert np.array_equal(kp, detect(v["image"][0, :n]).numpy())
        assert np.abs(kp).max() <= 1.0
