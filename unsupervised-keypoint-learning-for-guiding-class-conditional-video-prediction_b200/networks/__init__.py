"""Network builders — the operator surface of the reference's ``models/networks/__init__.py`` on the B200 engine.

Same function names, argument order and returned structures as /root/reference/models/networks/__init__.py
(encoder :7-26, image_encoder :29-33, pose_encoder :36-72, translator :75-102, img_discr :141-151).  The
reference builds a TF graph inside variable scopes; here the functions execute eagerly against the current
``engine.Context`` (set with ``set_context``, the analogue of TF's default graph), whose parameters carry the
TF variable names those scopes produce.  Images cross the boundary as NHWC float32 CUDA tensors in [-1,1];
activations inside are bf16 NHWC; keypoints / maps / heads are float32.
"""
import torch

from .. import engine as E
from .. import ops
from .. import tapconv as tc
from ..utils import model as model_utils
from . import layers  # noqa: F401
from .vgg import Vgg19  # noqa: F401

_CTX = None


def set_context(ctx):
    global _CTX
    _CTX = ctx
    return ctx


def get_context():
    if _CTX is None:
        raise RuntimeError("networks: no engine.Context set (call networks.set_context(ctx) first)")
    return _CTX


def build_parameters(ctx, n_pts=40, with_vgg=True):
    """Declare every variable of the stage-1 graph under its TF name (kernels HWIO) and allocate the flat buffers."""
    def conv(group, scope, k, cin, cout, bias=True):
        group.add(scope + "/conv2d/kernel", (k, k, cin, cout))
        if bias:
            group.add(scope + "/conv2d/bias", (cout,))

    def bn(scope, c):
        ctx.G.add(scope + "/gamma", (c,))
        ctx.G.add(scope + "/beta", (c,))
        ctx.S.add(scope + "/moving_mean", (c,))
        ctx.S.add(scope + "/moving_variance", (c,))

    for top in ("image_encoder/encoder/", "pose_encoder/encoder/"):
        conv(ctx.G, top + "conv_1", 7, 3, 32); bn(top + "b_norm_1", 32)
        conv(ctx.G, top + "conv_2", 3, 32, 32); bn(top + "b_norm_2", 32)
        f = 32
        for i in range(3):
            conv(ctx.G, top + "conv_%d" % (i * 2 + 3), 3, f, f * 2); bn(top + "b_norm_%d" % (i * 2 + 3), f * 2)
            f *= 2
            conv(ctx.G, top + "conv_%d" % (i * 2 + 4), 3, f, f); bn(top + "b_norm_%d" % (i * 2 + 4), f)
    # pose_encoder decoder (networks/__init__.py:41-66)
    skips = [256, 128, 64, 32]
    cin, conv_id, filters, size = 256, 1, 128, 16
    for i in range(4):
        f = int(filters)
        s = "pose_encoder/"
        conv(ctx.G, s + "conv_%d_0" % conv_id, 3, cin + (skips[i] if i > 0 else 0), f); bn(s + "b_norm_%d_0" % conv_id, f)
        conv(ctx.G, s + "conv_%d_1" % conv_id, 3, f, f); bn(s + "b_norm_%d_1" % conv_id, f)
        if size == 128:
            conv(ctx.G, s + "conv_0", 1, f, n_pts)
            break
        conv(ctx.G, s + "conv_%d_0" % (conv_id + 1), 3, f, f); bn(s + "b_norm_%d_0" % (conv_id + 1), f)
        conv(ctx.G, s + "conv_%d_1" % (conv_id + 1), 3, f, f); bn(s + "b_norm_%d_1" % (conv_id + 1), f)
        size *= 2; conv_id += 2; cin = f
        filters /= 2
    # translator (networks/__init__.py:75-102)
    cin, conv_id, filters, size = 128 + 2 * n_pts, 1, 256, 32
    while size <= 128:
        f = int(filters)
        s = "translator/"
        conv(ctx.G, s + "conv_%d_0" % conv_id, 3, cin, f); bn(s + "b_norm_%d_0" % conv_id, f)
        conv(ctx.G, s + "conv_%d_1" % conv_id, 3, f, f); bn(s + "b_norm_%d_1" % conv_id, f)
        if size == 128:
            conv(ctx.G, s + "conv_%d_0" % (conv_id + 1), 3, f, 3)
            conv(ctx.G, s + "conv_%d_1" % (conv_id + 1), 3, f, 1)
            break
        conv(ctx.G, s + "conv_%d_0" % (conv_id + 1), 3, f, f); bn(s + "b_norm_%d_0" % (conv_id + 1), f)
        conv(ctx.G, s + "conv_%d_1" % (conv_id + 1), 3, f, f); bn(s + "b_norm_%d_1" % (conv_id + 1), f)
        size *= 2; conv_id += 2; cin = f
        filters /= 2
    # img_discr (networks/__init__.py:141-151)
    conv(ctx.D, "img_discr/conv_0", 4, 3, 64)
    ch = 64
    for i in range(1, 6):
        conv(ctx.D, "img_discr/conv_%d" % i, 4, ch, ch * 2)
        ch *= 2
    conv(ctx.D, "img_discr/D_logit", 3, ch, 1, bias=False)
    if with_vgg:
        from .vgg import VGG_LAYERS
        for name, ci, co in VGG_LAYERS:
            ctx.V.add("vgg/%s/filter" % name, (3, 3, ci, co))
            ctx.V.add("vgg/%s/biases" % name, (co,))
    for g in (ctx.G, ctx.D, ctx.S, ctx.V):
        g.finalize()
    for n in ctx.S.names():
        if n.endswith("moving_variance"):
            ctx.S.p(n).fill_(1.0)
    for n in ctx.G.names():
        if n.endswith("/gamma"):
            ctx.G.p(n).fill_(1.0)
    return ctx


def _prep(x):
    """float32 NHWC image in [-1,1] -> bf16 [B,H,W,16] (channels 3..15 zero); bf16 inputs pass through."""
    if x.dtype == torch.bfloat16:
        return x
    return ops.image_prep(x)


def _prep7(x):
    """Encoder input: float32 NHWC image -> W-unrolled bf16 [B,H,W,32] (7 horizontal taps x 3 channels, SAME pad 3),
    so that the 7x7x3 conv_1 runs as a 7x1 convolution over 21 channels (7 taps instead of 49)."""
    if x.dtype == torch.bfloat16:
        return x
    return ops.image_prep_unrolled(x, 7, 3, 32)


def _cbr(ctx, srcs, conv_scope, bn_scope, train_mode, k=3, stride=1, upsample=False, need_input_grad=True, segments=1):
    return E.conv_layer(ctx, srcs, conv_scope + "/conv2d/kernel", conv_scope + "/conv2d/bias", k, stride, 0, bn=bn_scope,
                        train_mode=train_mode, upsample=upsample, need_input_grad=need_input_grad,
                        segments=segments if train_mode else 1)


def encoder(x, train_mode, filters=32, _scope="encoder/", _n_blocks=4, _segments=1):
    """reference networks/__init__.py:7-26.  x: prepared (W-unrolled, see _prep7) bf16 image.  Returns the 4 block features.
    `_n_blocks=3` skips conv_7/conv_8, whose output the stage-1 graph never consumes: TF prunes
    them from the D run; the G run still executes them for their moving-average updates."""
    ctx = get_context()
    p = _scope
    block_features = []
    sg = _segments if train_mode else 1
    x = E.conv_layer(ctx, [x], p + "conv_1/conv2d/kernel", p + "conv_1/conv2d/bias", (7, 1), 1, 0, bn=p + "b_norm_1",
                     train_mode=train_mode, need_input_grad=False, wshape=(7, 1, 21, filters), segments=sg)
    x = _cbr(ctx, [x], p + "conv_2", p + "b_norm_2", train_mode, segments=sg)
    block_features.append(x)
    for i in range(_n_blocks - 1):
        x = _cbr(ctx, [x], p + "conv_%d" % (i * 2 + 3), p + "b_norm_%d" % (i * 2 + 3), train_mode, stride=2, segments=sg)
        x = _cbr(ctx, [x], p + "conv_%d" % (i * 2 + 4), p + "b_norm_%d" % (i * 2 + 4), train_mode, segments=sg)
        block_features.append(x)
    return block_features


def image_encoder(x, train_mode):
    """reference networks/__init__.py:29-33: returns [x] + block features; consumers use [-2] (32x32x128)."""
    return [x] + encoder(_prep7(x), train_mode, _scope="image_encoder/encoder/")


def pose_encoder_logits(x, n_pts, train_mode, final_res=128, filters=128, _segments=1):
    """Everything of pose_encoder before get_coord (reference networks/__init__.py:36-66): fp32 logits [B,128,128,n_pts].
    _segments: the batch holds that many independent calls side by side (each normalised with its own batch statistics)."""
    ctx = get_context()
    sg = _segments
    block_features = encoder(_prep7(x), train_mode, _scope="pose_encoder/encoder/", _segments=sg)
    x = block_features[-1]
    size = x.shape[1]
    conv_id = 1
    s = "pose_encoder/"
    for i in range(4):
        srcs = [x, block_features[-1 * (i + 1)]] if i > 0 else [x]
        x = _cbr(ctx, srcs, s + "conv_%d_0" % conv_id, s + "b_norm_%d_0" % conv_id, train_mode, segments=sg)
        x = _cbr(ctx, [x], s + "conv_%d_1" % conv_id, s + "b_norm_%d_1" % conv_id, train_mode, segments=sg)
        if size == final_res:
            # 1x1 head: fp32 output (bf16 logits would move mu by up to 2.7e-3)
            x = E.conv_layer(ctx, [x], s + "conv_0/conv2d/kernel", s + "conv_0/conv2d/bias", 1, 1, 0, out_f32=True)
            break
        x = _cbr(ctx, [x], s + "conv_%d_0" % (conv_id + 1), s + "b_norm_%d_0" % (conv_id + 1), train_mode, segments=sg)
        x = _cbr(ctx, [x], s + "conv_%d_1" % (conv_id + 1), s + "b_norm_%d_1" % (conv_id + 1), train_mode, upsample=True, segments=sg)
        size = x.shape[1]
        conv_id += 2
        if filters >= 8:
            filters /= 2
    return x


def _keypoints(ctx, logits, map_hw):
    """get_coord x2 + stack (+ get_gaussian_maps) in the fused K1 kernel, with its tape entry."""
    from .. import k1
    mu, px, py, maps = k1.softargmax_render_fwd(logits, map_hw, want_prob=True)
    if ctx.trace is not None:
        ctx.trace.append(dict(scope="k1", kind="k1", mu=mu, maps=maps))
    if ctx.tape is not None:
        tape = ctx.tape
        H, W = logits.shape[1], logits.shape[2]

        def bwd():
            d_maps = tape.grad(maps) if maps is not None else None
            d_mu = tape.grad(mu)
            if d_maps is None and d_mu is None:
                return
            d_logits = k1.softargmax_render_bwd(d_maps, d_mu, mu, px, py, H, W)
            tape.set_grad(logits, ops.pack_channels([d_logits], d_logits.shape[-1]))   # f32 -> bf16 for the head conv
        tape.record(bwd)
    return mu, maps


def pose_encoder(x, n_pts, train_mode, final_res=128, filters=128):
    """reference networks/__init__.py:36-72: image -> keypoints mu [B,n_pts,2] (x,y) float32."""
    logits = pose_encoder_logits(x, n_pts, train_mode, final_res, filters)
    mu, _ = _keypoints(get_context(), logits, None)
    return mu


def pose_encoder_with_maps(x, n_pts, train_mode, map_hw=(32, 32)):
    """pose_encoder + get_gaussian_maps(mu, map_hw) with ONE pass over the logits (fused K1 kernel)."""
    logits = pose_encoder_logits(x, n_pts, train_mode)
    return _keypoints(get_context(), logits, map_hw)


def pose_encoder_pair_with_maps(im, future_im, n_pts, train_mode, map_hw=(32, 32)):
    """pose_encoder(im) and pose_encoder(future_im) (detector_translator_model.py:166-167: the SAME variables, two calls,
    each with its own batch-norm statistics and moving-average update) as ONE pass over [im; future_im]: every layer is a
    single launch with per-segment statistics, which halves the launch count of the detector and doubles the tiles per
    launch on its under-filled 16x16 / 32x32 layers.  Returns ((mu_cur, maps_cur), (mu_fut, maps_fut))."""
    ctx = get_context()
    B = im.shape[0]
    logits = pose_encoder_logits(torch.cat([im, future_im], dim=0), n_pts, train_mode, _segments=2)
    mu, maps = _keypoints(ctx, logits, map_hw)
    cur_map, fut_map = maps[:B], maps[B:]
    if ctx.tape is not None:
        # the two halves receive their gradients separately (joint_embedding's backward): they are views of one buffer
        d_maps = torch.empty_like(maps)
        ctx.tape.set_grad(maps, d_maps)
        ctx.tape.set_grad(cur_map, d_maps[:B])
        ctx.tape.set_grad(fut_map, d_maps[B:])
    return (mu[:B], cur_map), (mu[B:], fut_map)


def joint_embedding(embedding, cur_map, fut_map):
    """tf.concat([embeddings[-2], current_pt_map, future_pt_map], -1) (detector_translator_model.py:170) as one
    bf16 tensor whose channels are zero-padded to a multiple of 64 (208 -> 256)."""
    ctx = get_context()
    ctot = embedding.shape[-1] + cur_map.shape[-1] + fut_map.shape[-1]
    joint = ops.pack_channels([embedding, cur_map, fut_map], tc.round_up(ctot, 64))
    if ctx.trace is not None:
        ctx.trace.append(dict(scope="joint", kind="joint", out=joint[..., :ctot]))
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            g = tape.grad(joint)
            if g is None:
                return
            d_emb, _ = tape.acquire(embedding)
            d_cur, _ = tape.acquire(cur_map)
            d_fut, _ = tape.acquire(fut_map)
            ops.unpack_channels(g, [d_emb, d_cur, d_fut])
        tape.record(bwd)
    return joint


def translator_heads(x, train_mode, final_res=128, filters=256):
    """translator body + both heads fused into one conv: f32 [B,128,128,4] = (crude rgb, sigmoid(mask))."""
    ctx = get_context()
    size = x.shape[1]
    conv_id = 1
    s = "translator/"
    while size <= final_res:
        x = _cbr(ctx, [x], s + "conv_%d_0" % conv_id, s + "b_norm_%d_0" % conv_id, train_mode)
        x = _cbr(ctx, [x], s + "conv_%d_1" % conv_id, s + "b_norm_%d_1" % conv_id, train_mode)
        if size == final_res:
            n0, n1 = s + "conv_%d_0" % (conv_id + 1), s + "conv_%d_1" % (conv_id + 1)
            return E.conv_layer(ctx, [x], [n0 + "/conv2d/kernel", n1 + "/conv2d/kernel"],
                                [n0 + "/conv2d/bias", n1 + "/conv2d/bias"], 3, 1, 0, act=tc.ACT_SIGMOID_LAST, out_f32=True)
        x = _cbr(ctx, [x], s + "conv_%d_0" % (conv_id + 1), s + "b_norm_%d_0" % (conv_id + 1), train_mode)
        x = _cbr(ctx, [x], s + "conv_%d_1" % (conv_id + 1), s + "b_norm_%d_1" % (conv_id + 1), train_mode, upsample=True)
        size = x.shape[1]
        conv_id += 2
        if filters >= 8:
            filters /= 2
    raise ValueError("translator: input resolution %d exceeds final_res %d" % (size, final_res))


def translator(x, train_mode, final_res=128, filters=256):
    """reference networks/__init__.py:75-102: joint embedding -> (crude_output [B,128,128,3], mask [B,128,128,1])."""
    heads = translator_heads(x, train_mode, final_res, filters)
    return heads[..., :3], heads[..., 3:4]


def compose(im, heads, clip=False, want_parts=False):
    """final_output = im*mask + crude*(1-mask) (detector_translator_model.py:174; final_model.py:96-99 with clip)."""
    ctx = get_context()
    final, crude, mask = ops.compose_fwd(heads, im, clip=clip, want_parts=want_parts)
    if ctx.trace is not None:
        ctx.trace.append(dict(scope="final", kind="final", out=final))
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            g = tape.grad(final)
            if g is None:
                return
            tape.set_grad(heads, ops.compose_bwd(g, heads, im))
        tape.record(bwd)
    return (final, crude, mask) if want_parts else final


def _prep_with_grad(ctx, x, prep, unroll=None):
    """image_prep of a tensor that needs a gradient (the generated frame entering VGG / the discriminator).
    unroll = (kw, pad_left, cpad) selects the W-unrolled layout."""
    xp = ops.image_prep(x, prep) if unroll is None else ops.image_prep_unrolled(x, unroll[0], unroll[1], unroll[2], prep)
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            g = tape.grad(xp)
            if g is None:
                return
            dx, acc = tape.acquire(x)
            if unroll is None:
                ops.image_prep_bwd(g, dx, prep, accumulate=acc)
            else:
                ops.image_prep_unrolled_bwd(g, dx, unroll[0], unroll[1], prep, accumulate=acc)
        tape.record(bwd)
    return xp


def img_discr(x, need_input_grad=False, marker=None):
    """reference networks/__init__.py:141-151: image [B,128,128,3] -> logit [B,6,6,1] float32.
    marker(scope): called before each layer is recorded; the data-parallel train step hangs the all-reduce of a gradient
    bucket on it (a tape entry recorded there runs, in the backward pass, right after that layer's gradients)."""
    ctx = get_context()
    h = _prep_with_grad(ctx, x, ops.IDENT_PREP) if need_input_grad else ops.image_prep(x)
    h = E.conv_layer(ctx, [h], "img_discr/conv_0/conv2d/kernel", "img_discr/conv_0/conv2d/bias", 4, 2, 1, act=tc.ACT_LEAKY,
                     alpha=0.01, need_input_grad=need_input_grad)
    for i in range(1, 6):
        sc = "img_discr/conv_%d" % i
        if marker is not None:
            marker(sc + "/")
        h = E.conv_layer(ctx, [h], sc + "/conv2d/kernel", sc + "/conv2d/bias", 4, 2, 1, act=tc.ACT_LEAKY, alpha=0.01)
    return E.conv_layer(ctx, [h], "img_discr/D_logit/conv2d/kernel", None, 3, 1, 1, out_f32=True)


def maxpool(x, grad_rows=None):
    """grad_rows=(lo, hi): only that run of images carries a gradient (see engine.conv_layer)."""
    ctx = get_context()
    y = ops.maxpool_fwd(x)
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            g = tape.grad(y)
            if g is None:
                return
            dx, acc = tape.acquire(x)
            if grad_rows is None:
                ops.maxpool_bwd(g, x, dx, relu_mask=False, accumulate=acc)
            else:
                lo, hi = grad_rows
                ops.maxpool_bwd(g[lo:hi], x[lo:hi], dx[lo:hi], relu_mask=False, accumulate=acc)
        tape.record(bwd)
    return y
