"""Layer primitives — mirror of /root/reference/models/networks/layers.py:4-14 on the B200 engine.

``conv`` and ``batch_norm`` exist for API parity with the reference's call sites; the network builders use the
fused ``engine.conv_layer`` (convolution + BN statistics + normalise + ReLU [+ x2 resize]) directly, because on
this hardware the separate ops would each round-trip the activation through HBM.
"""
import torch

from .. import engine as E
from .. import ops
from .. import tapconv as tc


def _ctx():
    from . import get_context
    return get_context()


def conv(x, channels, kernel=4, stride=2, pad=0, use_bias=True, scope='conv_0'):
    """tf.pad(pad) + tf.layers.conv2d(padding='same') (reference layers.py:4-10).  x: bf16 NHWC (or a float32
    3-channel image); variables `<scope>/conv2d/{kernel,bias}` must exist in the current context."""
    ctx = _ctx()
    if x.dtype == torch.float32:
        x = ops.image_prep(x) if x.shape[-1] == 3 else x.to(torch.bfloat16)
    w = ctx.p(scope + "/conv2d/kernel")
    if w.shape[3] != channels or w.shape[0] != kernel:
        raise ValueError("conv %s: variable shape %s does not match channels=%d kernel=%d" %
                         (scope, tuple(w.shape), channels, kernel))
    return E.conv_layer(ctx, [x], scope + "/conv2d/kernel", scope + "/conv2d/bias" if use_bias else None, kernel, stride,
                        pad, act=tc.ACT_NONE)


def batch_norm(x, train_mode, scope='batch_norm'):
    """tf.contrib.layers.batch_norm(eps=1e-5, center, scale) (reference layers.py:13-14) as a stand-alone op:
    bf16 NHWC in, bf16 out (no activation).  Training mode uses batch statistics (biased variance)."""
    ctx = _ctx()
    N, H, W, C = x.shape
    gamma, beta = ctx.p(scope + "/gamma"), ctx.p(scope + "/beta")
    if train_mode:
        # stand-alone path (the fused builders take the statistics from the convolution epilogue instead)
        stats = torch.zeros((2, C), device=x.device, dtype=torch.float32)
        ssum, ssq = ops.channel_sum(x, stats[0]), ops.channel_sum(x, stats[1], squares=True)
        mm = ctx.p(scope + "/moving_mean") if ctx.update_moving else None
        mv = ctx.p(scope + "/moving_variance") if ctx.update_moving else None
        scale, shift, _, _ = ops.bn_finalize(ssum, ssq, None, gamma, beta, N * H * W, mm, mv)
    else:
        scale = gamma * torch.rsqrt(ctx.p(scope + "/moving_variance") + 1e-5)
        shift = beta - ctx.p(scope + "/moving_mean") * scale
    return ops.bn_act_apply(x, scale.contiguous(), shift.contiguous(), relu=False, upsample=False)
