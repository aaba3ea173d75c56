"""torch-CPU restatement of the reference's stage-1 networks and losses (TEST INFRASTRUCTURE ONLY).

Follows, function for function:
  encoder / image_encoder / pose_encoder / translator / img_discr
      /root/reference/models/networks/__init__.py:7-102,141-151
  Vgg19.build                     /root/reference/models/networks/vgg.py:13-55
  forward wiring + losses         /root/reference/models/detector_translator_model.py:160-203,246-289
Parameters are a name-keyed dict using the TF variable names the reference's scopes produce
(SURVEY.md §8c item 9), kernels in HWIO.  The wiring is pinned against the reference's own source
executed through oracle/tf_shim (tests/test_oracle_networks.py); TF op semantics are in tf_ops.py.
"""
import numpy as np
import torch

from . import tf_ops as T
from . import k1_torch
from . import precision

VGG_LAYERS = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
              ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256),
              ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512),
              ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512)]
VGG_MEAN = [103.939, 116.779, 123.68]


# ------------------------------------------------------------------------------------------------
# layer spec (shared by init and by tests that want the layer list)
# ------------------------------------------------------------------------------------------------
def encoder_spec(prefix):
    """[(conv_scope, bn_scope, k, stride, cin, cout)] for networks.encoder (filters=32)."""
    spec = [(prefix + "conv_1", prefix + "b_norm_1", 7, 1, 3, 32), (prefix + "conv_2", prefix + "b_norm_2", 3, 1, 32, 32)]
    f = 32
    for i in range(3):
        spec.append((prefix + "conv_%d" % (i * 2 + 3), prefix + "b_norm_%d" % (i * 2 + 3), 3, 2, f, f * 2))
        f *= 2
        spec.append((prefix + "conv_%d" % (i * 2 + 4), prefix + "b_norm_%d" % (i * 2 + 4), 3, 1, f, f))
    return spec


def pose_decoder_spec(n_pts, final_res=128, filters=128):
    """Decoder part of pose_encoder: [(conv, bn|None, k, stride, cin, cout)] in execution order."""
    spec = []
    size, conv_id, cin = 16, 1, 256
    skips = [256, 128, 64, 32]  # block_features[-1 * (i + 1)] channels
    for i in range(4):
        c_in = cin + (skips[i] if i > 0 else 0)
        f = int(filters)
        spec.append(("pose_encoder/conv_%d_0" % conv_id, "pose_encoder/b_norm_%d_0" % conv_id, 3, 1, c_in, f))
        spec.append(("pose_encoder/conv_%d_1" % conv_id, "pose_encoder/b_norm_%d_1" % conv_id, 3, 1, f, f))
        if size == final_res:
            spec.append(("pose_encoder/conv_0", None, 1, 1, f, n_pts))
            break
        spec.append(("pose_encoder/conv_%d_0" % (conv_id + 1), "pose_encoder/b_norm_%d_0" % (conv_id + 1), 3, 1, f, f))
        spec.append(("pose_encoder/conv_%d_1" % (conv_id + 1), "pose_encoder/b_norm_%d_1" % (conv_id + 1), 3, 1, f, f))
        size *= 2
        conv_id += 2
        cin = f
        if filters >= 8:
            filters /= 2
    return spec


def translator_spec(cin=208, start_res=32, final_res=128, filters=256):
    spec = []
    size, conv_id = start_res, 1
    while size <= final_res:
        f = int(filters)
        spec.append(("translator/conv_%d_0" % conv_id, "translator/b_norm_%d_0" % conv_id, 3, 1, cin, f))
        spec.append(("translator/conv_%d_1" % conv_id, "translator/b_norm_%d_1" % conv_id, 3, 1, f, f))
        if size == final_res:
            spec.append(("translator/conv_%d_0" % (conv_id + 1), None, 3, 1, f, 3))
            spec.append(("translator/conv_%d_1" % (conv_id + 1), None, 3, 1, f, 1))
            break
        spec.append(("translator/conv_%d_0" % (conv_id + 1), "translator/b_norm_%d_0" % (conv_id + 1), 3, 1, f, f))
        spec.append(("translator/conv_%d_1" % (conv_id + 1), "translator/b_norm_%d_1" % (conv_id + 1), 3, 1, f, f))
        size *= 2
        conv_id += 2
        cin = f
        if filters >= 8:
            filters /= 2
    return spec


def discr_spec():
    spec = [("img_discr/conv_0", 4, 2, 3, 64, True)]
    ch = 64
    for i in range(1, 6):
        spec.append(("img_discr/conv_%d" % i, 4, 2, ch, ch * 2, True))
        ch *= 2
    spec.append(("img_discr/D_logit", 3, 1, ch, 1, False))
    return spec


def init_params(seed=0, n_pts=40, dtype=torch.float64, with_vgg=True, bias_scale=0.0):
    """Seeded xavier-uniform weights (biases 0 like TF, or small normal if bias_scale>0 to exercise the bias path),
    BN gamma=1 beta=0 moving 0/1; VGG random (He-style) since vgg19.npy is not shipped."""
    rng = np.random.default_rng(seed)
    P = {}

    def add_conv(scope, k, cin, cout, use_bias=True):
        P[scope + "/conv2d/kernel"] = T.xavier_uniform(rng, (k, k, cin, cout), dtype)
        if use_bias:
            b = rng.normal(0, bias_scale, cout) if bias_scale > 0 else np.zeros(cout)
            P[scope + "/conv2d/bias"] = torch.from_numpy(b).to(dtype)

    def add_bn(scope, c):
        P[scope + "/gamma"] = torch.ones(c, dtype=dtype)
        P[scope + "/beta"] = torch.zeros(c, dtype=dtype)
        P[scope + "/moving_mean"] = torch.zeros(c, dtype=dtype)
        P[scope + "/moving_variance"] = torch.ones(c, dtype=dtype)

    for top in ("image_encoder/encoder/", "pose_encoder/encoder/"):
        for conv, bn, k, s, cin, cout in encoder_spec(top):
            add_conv(conv, k, cin, cout)
            add_bn(bn, cout)
    for conv, bn, k, s, cin, cout in pose_decoder_spec(n_pts) + translator_spec(128 + 2 * n_pts):
        add_conv(conv, k, cin, cout)
        if bn is not None:
            add_bn(bn, cout)
    for scope, k, s, cin, cout, use_bias in discr_spec():
        add_conv(scope, k, cin, cout, use_bias)
    if with_vgg:
        for name, cin, cout in VGG_LAYERS:
            std = np.sqrt(2.0 / (9 * cin))
            P["vgg/%s/filter" % name] = torch.from_numpy(rng.normal(0, std, (3, 3, cin, cout))).to(dtype)
            P["vgg/%s/biases" % name] = torch.from_numpy(rng.normal(0, 0.05, cout)).to(dtype)
    return P


def randomize_bn(P, seed=1):
    """Non-trivial BN parameters / moving statistics (so inference-mode folding is really exercised)."""
    rng = np.random.default_rng(seed)
    for k in list(P):
        dt = P[k].dtype
        n = P[k].shape[0]
        if k.endswith("/gamma"):
            P[k] = torch.from_numpy(rng.uniform(0.5, 1.5, n)).to(dt)
        elif k.endswith("/beta"):
            P[k] = torch.from_numpy(rng.normal(0, 0.2, n)).to(dt)
        elif k.endswith("/moving_mean"):
            P[k] = torch.from_numpy(rng.normal(0, 0.2, n)).to(dt)
        elif k.endswith("/moving_variance"):
            P[k] = torch.from_numpy(rng.uniform(0.5, 2.0, n)).to(dt)
    return P


# ------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------
class Ctx:
    """Carries the parameter dict and collects BN moving-average updates (TF's UPDATE_OPS).

    `q` is the precision model (oracle/precision.py): `Exact` (default) is the reference's float arithmetic;
    `Bf16Faithful` additionally rounds to bf16 at the points where the B200 path stores bf16 (weights as the tensor
    cores read them, stored activations, stored activation gradients), which turns a whole-step comparison into a test
    of the wiring instead of a measurement of bf16 drift."""

    def __init__(self, params, q=None, sub=None):
        self.P = params
        self.q = q if q is not None else precision.Exact()
        self.updates = []   # [(name, new_value)] in graph order; two entries per BN for the shared pose_encoder
        self.taps = {}      # optional: named intermediate activations
        # Forward substitution ("teacher forcing", tests only): name -> list of tensors recorded from the implementation
        # under test, consumed in call order.  s(name, x) returns a tensor with the recorded VALUE and x's gradient
        # path, and logs how far x (computed here from the substituted inputs) was from it.  With every stored tensor
        # substituted, ReLU / max-pool / L1-sign decisions and BN statistics are those of the implementation under
        # test, so the backward passes can be compared without the chaotic divergence of two bf16 forward passes.
        self.sub = sub
        self.sub_err = {}

    def s(self, name, x):
        if self.sub is None or name not in self.sub:
            return x
        if not self.sub[name]:
            raise KeyError("substitution list for %r is exhausted" % name)
        v = self.sub[name].pop(0).to(x.dtype)
        if tuple(v.shape) != tuple(x.shape):
            raise ValueError("substitution %r: shape %r vs %r" % (name, tuple(v.shape), tuple(x.shape)))
        xd = x.detach()
        self.sub_err.setdefault(name, []).append(
            (float((xd - v).norm() / (v.norm() + 1e-300)), float((xd - v).abs().max()), float(v.abs().max())))
        return x + (v - xd)

    def conv(self, x, scope, stride=1, pad=0, use_bias=True):
        b = self.P[scope + "/conv2d/bias"] if use_bias else None
        return T.conv2d(x, self.q.w(self.P[scope + "/conv2d/kernel"]), b, stride, pad)

    def bn(self, x, scope, train_mode):
        P = self.P
        if train_mode and not isinstance(self.q, precision.Exact):
            # B200 path: statistics from the fp32 accumulators, applied to the bf16-stored convolution output; the
            # gradient leaving the BN backward is stored as bf16 (one rounding of the complete dy)
            x = self.q.g(x)
            n = x.shape[0] * x.shape[1] * x.shape[2]
            mean = x.mean(dim=(0, 1, 2))
            var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
            cs = getattr(self, "_conv_scope", scope)
            mean_s, rstd_s = self.s(cs + ":mean", mean), self.s(cs + ":rstd", torch.rsqrt(var + 1e-5))
            y = (self.s(cs + ":pre", self.q.w(x)) - mean_s) * rstd_s * P[scope + "/gamma"] + P[scope + "/beta"]
            mm = P[scope + "/moving_mean"] * 0.999 + mean.detach() * (1 - 0.999)
            mv = P[scope + "/moving_variance"] * 0.999 + (var * (n / max(n - 1, 1))).detach() * (1 - 0.999)
        else:
            y, mm, mv = T.batch_norm(x, P[scope + "/gamma"], P[scope + "/beta"], P[scope + "/moving_mean"],
                                     P[scope + "/moving_variance"], train_mode)
        if train_mode:
            self.updates.append((scope + "/moving_mean", mm))
            self.updates.append((scope + "/moving_variance", mv))
        return y

    def cbr(self, x, conv_scope, bn_scope, train_mode, stride=1, store=True):
        """conv + BN + ReLU.  store=False: the caller applies an x2 resize before the result is stored."""
        c = self.conv(x, conv_scope, stride)
        self.taps[conv_scope + ":pre"] = c
        self._conv_scope = conv_scope
        y = torch.relu(self.bn(c, bn_scope, train_mode))
        if store:
            y = self.s(conv_scope, self.q.a(y))
        self.taps[conv_scope] = y
        return y

    def upsample2x(self, x, conv_scope):
        """tf.image.resize_images x2 of a cbr(..., store=False) result; rounding points of the fused B200 kernel: the
        gradient arriving at the un-resized activation and the stored resized activation."""
        size = x.shape[1]
        y = self.s(conv_scope, self.q.a(T.resize_bilinear_legacy(self.q.g(x), 2 * size, 2 * size)))
        self.taps[conv_scope + ":up"] = y
        return y

    def conv_act(self, x, scope, act, stride=1, pad=0, use_bias=True):
        """conv + bias + pointwise activation stored as bf16 (VGG, img_discr): the masked gradient is stored too."""
        return self.s(scope, self.q.a(act(self.q.g(self.conv(x, scope, stride, pad, use_bias)))))


def encoder(ctx, x, train_mode, prefix):
    """networks/__init__.py:7-26."""
    feats = []
    p = prefix + "encoder/"
    x = ctx.q.a(x)
    x = ctx.cbr(x, p + "conv_1", p + "b_norm_1", train_mode)
    x = ctx.cbr(x, p + "conv_2", p + "b_norm_2", train_mode)
    feats.append(x)
    for i in range(3):
        x = ctx.cbr(x, p + "conv_%d" % (i * 2 + 3), p + "b_norm_%d" % (i * 2 + 3), train_mode, stride=2)
        x = ctx.cbr(x, p + "conv_%d" % (i * 2 + 4), p + "b_norm_%d" % (i * 2 + 4), train_mode)
        feats.append(x)
    return feats


def image_encoder(ctx, x, train_mode):
    """networks/__init__.py:29-33."""
    return [x] + encoder(ctx, x, train_mode, "image_encoder/")


def pose_encoder_logits(ctx, x, n_pts, train_mode, final_res=128, filters=128):
    """networks/__init__.py:36-66 (everything before get_coord)."""
    feats = encoder(ctx, x, train_mode, "pose_encoder/")
    x = feats[-1]
    size = x.shape[1]
    conv_id = 1
    s = "pose_encoder/"
    for i in range(4):
        if i > 0:
            x = torch.cat([x, feats[-1 * (i + 1)]], dim=-1)
        x = ctx.cbr(x, s + "conv_%d_0" % conv_id, s + "b_norm_%d_0" % conv_id, train_mode)
        x = ctx.cbr(x, s + "conv_%d_1" % conv_id, s + "b_norm_%d_1" % conv_id, train_mode)
        if size == final_res:
            x = ctx.s(s + "conv_0", ctx.q.g(ctx.conv(x, s + "conv_0")))
            break
        x = ctx.cbr(x, s + "conv_%d_0" % (conv_id + 1), s + "b_norm_%d_0" % (conv_id + 1), train_mode)
        x = ctx.cbr(x, s + "conv_%d_1" % (conv_id + 1), s + "b_norm_%d_1" % (conv_id + 1), train_mode, store=False)
        x = ctx.upsample2x(x, s + "conv_%d_1" % (conv_id + 1))
        size = x.shape[1]
        conv_id += 2
        if filters >= 8:
            filters /= 2
    return x


def pose_encoder(ctx, x, n_pts, train_mode, final_res=128, filters=128):
    """networks/__init__.py:36-72 -> mu [B,n_pts,2]."""
    logits = pose_encoder_logits(ctx, x, n_pts, train_mode, final_res, filters)
    return ctx.s("mu", k1_torch.soft_argmax(logits))


def translator(ctx, x, train_mode, final_res=128, filters=256):
    """networks/__init__.py:75-102 -> (crude [B,128,128,3], mask [B,128,128,1])."""
    size = x.shape[1]
    conv_id = 1
    s = "translator/"
    while size <= final_res:
        x = ctx.cbr(x, s + "conv_%d_0" % conv_id, s + "b_norm_%d_0" % conv_id, train_mode)
        x = ctx.cbr(x, s + "conv_%d_1" % conv_id, s + "b_norm_%d_1" % conv_id, train_mode)
        if size == final_res:
            crude = ctx.s(s + "conv_%d_0" % (conv_id + 1), ctx.q.g(ctx.conv(x, s + "conv_%d_0" % (conv_id + 1))))
            mask = ctx.s(s + "conv_%d_1:sigmoid" % (conv_id + 1), torch.sigmoid(ctx.q.g(ctx.conv(x, s + "conv_%d_1" % (conv_id + 1)))))
            break
        x = ctx.cbr(x, s + "conv_%d_0" % (conv_id + 1), s + "b_norm_%d_0" % (conv_id + 1), train_mode)
        x = ctx.cbr(x, s + "conv_%d_1" % (conv_id + 1), s + "b_norm_%d_1" % (conv_id + 1), train_mode, store=False)
        x = ctx.upsample2x(x, s + "conv_%d_1" % (conv_id + 1))
        size = x.shape[1]
        conv_id += 2
        if filters >= 8:
            filters /= 2
    return crude, mask


def img_discr(ctx, x):
    """networks/__init__.py:141-151 -> logit [B,6,6,1] at 128x128 input."""
    def lrelu(t):
        return T.leaky_relu(t, 0.01)
    x = ctx.conv_act(ctx.q.a(x), "img_discr/conv_0", lrelu, stride=2, pad=1)
    for i in range(1, 6):
        x = ctx.conv_act(x, "img_discr/conv_%d" % i, lrelu, stride=2, pad=1)
    return ctx.s("img_discr/D_logit", ctx.q.g(ctx.conv(x, "img_discr/D_logit", stride=1, pad=1, use_bias=False)))


def vgg19(ctx, rgb):
    """vgg.py:13-43: rgb in [0,255] -> [conv1_2, conv2_2, conv3_4, conv4_4, conv5_4]."""
    P = ctx.P
    mean = torch.tensor(VGG_MEAN, dtype=rgb.dtype)
    bgr = torch.stack([rgb[..., 2] - mean[0], rgb[..., 1] - mean[1], rgb[..., 0] - mean[2]], dim=-1)

    q = ctx.q

    def cl(x, name):
        return ctx.s("vgg/" + name, q.a(torch.relu(q.g(T.conv2d(x, q.w(P["vgg/%s/filter" % name]), P["vgg/%s/biases" % name], 1, 0)))))

    x = cl(q.a(bgr), "conv1_1"); c12 = cl(x, "conv1_2"); x = T.max_pool_2x2(c12)
    x = cl(x, "conv2_1"); c22 = cl(x, "conv2_2"); x = T.max_pool_2x2(c22)
    x = cl(x, "conv3_1"); x = cl(x, "conv3_2"); x = cl(x, "conv3_3"); c34 = cl(x, "conv3_4"); x = T.max_pool_2x2(c34)
    x = cl(x, "conv4_1"); x = cl(x, "conv4_2"); x = cl(x, "conv4_3"); c44 = cl(x, "conv4_4"); x = T.max_pool_2x2(c44)
    x = cl(x, "conv5_1"); x = cl(x, "conv5_2"); x = cl(x, "conv5_3"); c54 = cl(x, "conv5_4")
    return [c12, c22, c34, c44, c54]


# ------------------------------------------------------------------------------------------------
# stage-1 model (detector_translator_model.py)
# ------------------------------------------------------------------------------------------------
def forward_pass(ctx, im, future_im, n_pts, is_training):
    """_define_forward_pass (:160-184)."""
    emb = image_encoder(ctx, im, is_training)
    cur_pt = pose_encoder(ctx, im, n_pts, is_training)
    fut_pt = pose_encoder(ctx, future_im, n_pts, is_training)
    cur_map = ctx.s("maps", k1_torch.get_gaussian_maps(cur_pt, [32, 32]))
    fut_map = ctx.s("maps", k1_torch.get_gaussian_maps(fut_pt, [32, 32]))
    joint = ctx.s("joint", ctx.q.a(torch.cat([emb[-2], cur_map, fut_map], dim=-1)))
    crude, mask = translator(ctx, joint, is_training)
    final = ctx.s("final", im * mask + crude * (1 - mask))
    return {"final_output": final, "crude_output": crude, "mask": mask, "current_pt": cur_pt, "future_pt": fut_pt,
            "current_map": cur_map, "future_map": fut_map, "embedding": emb[-2]}


def loss_D(ctx, future_im_pred, future_im):
    """_compute_loss_D (:246-259)."""
    real_ = img_discr(ctx, future_im)
    fake_ = img_discr(ctx, future_im_pred)
    real_loss = T.sigmoid_cross_entropy_with_logits(real_, torch.ones_like(real_)).mean()
    fake_loss = T.sigmoid_cross_entropy_with_logits(fake_, torch.zeros_like(fake_)).mean()
    return real_loss + fake_loss, real_loss, fake_loss


def perceptual_loss(ctx, gt_image, pred_image):
    """_compute_perceptual_loss (:274-289)."""
    ims = torch.cat([gt_image, pred_image], dim=0)
    feats = vgg19(ctx, ims)
    losses = []
    for f in feats:
        half = f.shape[0] // 2
        losses.append((f[:half] - f[half:]).abs().mean())
    return torch.stack(losses).mean()


def loss_G(ctx, future_im_pred, future_im):
    """_compute_loss_G (:261-272)."""
    recon = perceptual_loss(ctx, (future_im + 1) / 2.0 * 255.0, (future_im_pred + 1) / 2.0 * 255.0)
    fake_ = img_discr(ctx, future_im_pred)
    adv = T.sigmoid_cross_entropy_with_logits(fake_, torch.ones_like(fake_)).mean()
    return recon + adv, recon, adv
