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
    """Carries the parameter dict and collects BN moving-average updates (TF's UPDATE_OPS)."""

    def __init__(self, params):
        self.P = params
        self.updates = []   # [(name, new_value)] in graph order; two entries per BN for the shared pose_encoder
        self.taps = {}      # optional: named intermediate activations

    def conv(self, x, scope, stride=1, pad=0, use_bias=True):
        b = self.P[scope + "/conv2d/bias"] if use_bias else None
        return T.conv2d(x, self.P[scope + "/conv2d/kernel"], b, stride, pad)

    def bn(self, x, scope, train_mode):
        P = self.P
        y, mm, mv = T.batch_norm(x, P[scope + "/gamma"], P[scope + "/beta"], P[scope + "/moving_mean"],
                                 P[scope + "/moving_variance"], train_mode)
        if train_mode:
            self.updates.append((scope + "/moving_mean", mm))
            self.updates.append((scope + "/moving_variance", mv))
        return y

    def cbr(self, x, conv_scope, bn_scope, train_mode, stride=1):
        y = torch.relu(self.bn(self.conv(x, conv_scope, stride), bn_scope, train_mode))
        self.taps[conv_scope] = y
        return y


def encoder(ctx, x, train_mode, prefix):
    """networks/__init__.py:7-26."""
    feats = []
    p = prefix + "encoder/"
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
            x = ctx.conv(x, s + "conv_0")
            break
        x = ctx.cbr(x, s + "conv_%d_0" % (conv_id + 1), s + "b_norm_%d_0" % (conv_id + 1), train_mode)
        x = ctx.cbr(x, s + "conv_%d_1" % (conv_id + 1), s + "b_norm_%d_1" % (conv_id + 1), train_mode)
        x = T.resize_bilinear_legacy(x, 2 * size, 2 * size)
        size = x.shape[1]
        conv_id += 2
        if filters >= 8:
            filters /= 2
    return x


def pose_encoder(ctx, x, n_pts, train_mode, final_res=128, filters=128):
    """networks/__init__.py:36-72 -> mu [B,n_pts,2]."""
    logits = pose_encoder_logits(ctx, x, n_pts, train_mode, final_res, filters)
    return k1_torch.soft_argmax(logits)


def translator(ctx, x, train_mode, final_res=128, filters=256):
    """networks/__init__.py:75-102 -> (crude [B,128,128,3], mask [B,128,128,1])."""
    size = x.shape[1]
    conv_id = 1
    s = "translator/"
    while size <= final_res:
        x = ctx.cbr(x, s + "conv_%d_0" % conv_id, s + "b_norm_%d_0" % conv_id, train_mode)
        x = ctx.cbr(x, s + "conv_%d_1" % conv_id, s + "b_norm_%d_1" % conv_id, train_mode)
        if size == final_res:
            crude = ctx.conv(x, s + "conv_%d_0" % (conv_id + 1))
            mask = torch.sigmoid(ctx.conv(x, s + "conv_%d_1" % (conv_id + 1)))
            break
        x = ctx.cbr(x, s + "conv_%d_0" % (conv_id + 1), s + "b_norm_%d_0" % (conv_id + 1), train_mode)
        x = ctx.cbr(x, s + "conv_%d_1" % (conv_id + 1), s + "b_norm_%d_1" % (conv_id + 1), train_mode)
        x = T.resize_bilinear_legacy(x, 2 * size, 2 * size)
        size = x.shape[1]
        conv_id += 2
        if filters >= 8:
            filters /= 2
    return crude, mask


def img_discr(ctx, x):
    """networks/__init__.py:141-151 -> logit [B,6,6,1] at 128x128 input."""
    x = T.leaky_relu(ctx.conv(x, "img_discr/conv_0", stride=2, pad=1), 0.01)
    for i in range(1, 6):
        x = T.leaky_relu(ctx.conv(x, "img_discr/conv_%d" % i, stride=2, pad=1), 0.01)
    return ctx.conv(x, "img_discr/D_logit", stride=1, pad=1, use_bias=False)


def vgg19(ctx, rgb):
    """vgg.py:13-43: rgb in [0,255] -> [conv1_2, conv2_2, conv3_4, conv4_4, conv5_4]."""
    P = ctx.P
    mean = torch.tensor(VGG_MEAN, dtype=rgb.dtype)
    bgr = torch.stack([rgb[..., 2] - mean[0], rgb[..., 1] - mean[1], rgb[..., 0] - mean[2]], dim=-1)

    def cl(x, name):
        return torch.relu(T.conv2d(x, P["vgg/%s/filter" % name], P["vgg/%s/biases" % name], 1, 0))

    x = cl(bgr, "conv1_1"); c12 = cl(x, "conv1_2"); x = T.max_pool_2x2(c12)
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
    cur_map = k1_torch.get_gaussian_maps(cur_pt, [32, 32])
    fut_map = k1_torch.get_gaussian_maps(fut_pt, [32, 32])
    joint = torch.cat([emb[-2], cur_map, fut_map], dim=-1)
    crude, mask = translator(ctx, joint, is_training)
    final = im * mask + crude * (1 - mask)
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
