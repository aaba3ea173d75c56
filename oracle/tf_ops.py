"""TensorFlow-1.12 op semantics restated on torch CPU tensors (TEST INFRASTRUCTURE ONLY).

The reference's arithmetic lives in the un-vendored dependency tensorflow-gpu==1.12.0
(/root/reference/requirements.txt:16), which cannot be installed here.  These functions restate the
TF-1.12 behaviour of the ops the stage-1 path calls; each cites the reference call site it serves.
All tensors are NHWC like the reference's; dtype follows the input (float64 for golden values, float32
for the timed CPU baseline).  Parity status of THESE semantics: unpinned (no TF to run), see
oracle/__init__.py.
"""
import math

import torch
import torch.nn.functional as F


def same_pad(in_size, k, s):
    """TF 'SAME': out=ceil(in/s); total=max((out-1)*s+k-in,0); before=total//2; after=total-before."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return before, total - before


def conv2d(x, kernel, bias=None, stride=1, pad=0):
    """layers.conv (models/networks/layers.py:4-10): tf.pad(pad) then tf.layers.conv2d(padding='same').

    x [B,H,W,Cin]; kernel HWIO [k,k,Cin,Cout]; bias [Cout] or None.
    """
    kh, kw = kernel.shape[0], kernel.shape[1]
    xc = x.permute(0, 3, 1, 2)
    if pad:
        xc = F.pad(xc, (pad, pad, pad, pad))
    pt, pb = same_pad(xc.shape[2], kh, stride)
    pl, pr = same_pad(xc.shape[3], kw, stride)
    xc = F.pad(xc, (pl, pr, pt, pb))
    y = F.conv2d(xc, kernel.permute(3, 2, 0, 1).contiguous(), bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def batch_norm(x, gamma, beta, moving_mean, moving_var, train_mode, eps=1e-5, decay=0.999):
    """layers.batch_norm (layers.py:13-14): tf.contrib.layers.batch_norm(fused), NHWC.

    train: normalise with batch mean / BIASED variance; moving_var is updated with the UNBIASED one;
    moving <- moving*decay + batch*(1-decay).  Returns (y, new_moving_mean, new_moving_var).
    """
    if train_mode:
        n = x.shape[0] * x.shape[1] * x.shape[2]
        mean = x.mean(dim=(0, 1, 2))
        var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
        y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
        unbiased = var * (n / max(n - 1, 1))
        new_mm = moving_mean * decay + mean.detach() * (1 - decay)
        new_mv = moving_var * decay + unbiased.detach() * (1 - decay)
        return y, new_mm, new_mv
    y = (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta
    return y, moving_mean, moving_var


def resize_bilinear_legacy(x, out_h, out_w):
    """tf.image.resize_images default in TF 1.12 (networks/__init__.py:63,98): bilinear,
    align_corners=False, NO half-pixel centres: src = dst * (in/out)."""
    B, H, W, C = x.shape

    def axis(n_in, n_out):
        scale = n_in / n_out
        src = torch.arange(n_out, dtype=x.dtype) * scale
        lo = src.floor().long().clamp(max=n_in - 1)
        hi = (lo + 1).clamp(max=n_in - 1)
        frac = (src - lo.to(x.dtype))
        return lo, hi, frac

    lo_h, hi_h, fh = axis(H, out_h)
    lo_w, hi_w, fw = axis(W, out_w)
    top = x[:, lo_h]
    bot = x[:, hi_h]
    rows = top + (bot - top) * fh.view(1, -1, 1, 1)
    left = rows[:, :, lo_w]
    right = rows[:, :, hi_w]
    return left + (right - left) * fw.view(1, 1, -1, 1)


def max_pool_2x2(x):
    """tf.nn.max_pool 2x2 s2 SAME (vgg.py:45-46); even sizes only on this path."""
    xc = x.permute(0, 3, 1, 2)
    H, W = xc.shape[2], xc.shape[3]
    if H % 2 or W % 2:
        xc = F.pad(xc, (0, W % 2, 0, H % 2), value=-math.inf)
    return F.max_pool2d(xc, 2, 2).permute(0, 2, 3, 1)


def sigmoid_cross_entropy_with_logits(logits, labels):
    """max(x,0) - x*z + log1p(exp(-|x|)) (detector_translator_model.py:249-254,265-267)."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))


def leaky_relu(x, alpha):
    return torch.where(x >= 0, x, x * alpha)


def xavier_uniform(rng, shape, dtype):
    """tf.contrib.layers.xavier_initializer (layers.py:8): U(+-sqrt(6/(fan_in+fan_out))), fan = k*k*C."""
    k1, k2, cin, cout = shape
    limit = math.sqrt(6.0 / (k1 * k2 * cin + k1 * k2 * cout))
    import numpy as np
    return torch.from_numpy(rng.uniform(-limit, limit, size=shape).astype(np.float64)).to(dtype)


def adam_tf(param, grad, m, v, t, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update (detector_translator_model.py:198-202); t is the 1-based step."""
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m = m + (grad - m) * (1 - beta1)
    v = v + (grad * grad - v) * (1 - beta2)
    param = param - lr_t * m / (v.sqrt() + eps)
    return param, m, v


def exponential_decay(lr0, step, decay_steps, decay_rate):
    """tf.train.exponential_decay, staircase=False (detector_translator_model.py:193-195)."""
    return lr0 * decay_rate ** (step / decay_steps)
