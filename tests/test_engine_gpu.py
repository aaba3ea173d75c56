"""GPU: every forward/backward building block of the engine against fp64 autograd of the oracle ops, on IDENTICAL
(bf16-rounded) inputs, so that rounding noise of a deep stack cannot hide a wrong formula.

Tolerance: BASELINE.json's "bf16 conv outputs <= 1e-2 relative to fp32" — applied as max|err| <= 1e-2 * max|ref|
for bf16 tensors; f32 outputs and f32-accumulated weight gradients are held to 2e-3.
"""
import numpy as np
import pytest
import torch

from oracle import tf_ops as T

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _close(got, ref, tol, what):
    """Forward tensors: max|err| <= tol*max|ref|.  Gradients ("d..." names) are sums over many pixels whose ReLU
    masks are decided on bf16-rounded pre-activations: a unit whose pre-activation lies within the bf16 rounding
    error of zero flips its mask and moves the sum by its whole gradient.  With ~0.3 % of the units affected the
    expected relative L2 error of such a sum is a few per cent, so gradients are held to 5*tol relative L2 and a
    cosine of at least 0.998 (a wrong formula lands far outside both)."""
    got, ref = got.double().cpu(), ref.double().cpu()
    if what.lower().startswith("d") or "grad" in what:
        err = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
        cos = (torch.dot(got.flatten(), ref.flatten()) / (got.norm() * ref.norm() + 1e-30)).item()
        assert err <= 5 * tol and cos >= 0.998, "%s: relative L2 error %.3e (limit %.1e), cos %.5f" % (what, err, 5 * tol, cos)
        return
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item() / scale
    assert err <= tol, "%s: max err / max|ref| = %.3e > %.1e" % (what, err, tol)


def _mk_ctx(dev, specs, bn=None):
    """specs: {name: tensor}; bn: scope -> channels."""
    from kp_b200 import engine as E
    ctx = E.Context(dev)
    for n, v in specs.items():
        ctx.G.add(n, tuple(v.shape))
    for scope, c in (bn or {}).items():
        ctx.G.add(scope + "/gamma", (c,)); ctx.G.add(scope + "/beta", (c,))
        ctx.S.add(scope + "/moving_mean", (c,)); ctx.S.add(scope + "/moving_variance", (c,))
    for g in (ctx.G, ctx.D, ctx.S, ctx.V):
        g.finalize()
    for n, v in specs.items():
        ctx.G.p(n).copy_(v.to(dev))
    ctx.params_changed()
    return ctx


CONV_BN_CASES = [
    # (N,H,W,[C],k,stride,cout,upsample)
    (2, 16, 16, [64], 3, 1, 64, False),
    (2, 16, 16, [32], 3, 1, 32, True),
    (2, 32, 32, [16], 3, 1, 16, False),
    (2, 16, 16, [64, 64], 3, 1, 32, False),
    (2, 32, 32, [32], 3, 2, 64, False),
    (1, 32, 32, [16], 7, 1, 32, False),
    (3, 8, 8, [128], 3, 1, 128, True),
    # output extent not a multiple of the pixel tile (20x20 -> 4x32 tiles): rows outside the image still see real input
    # through their taps and must stay out of the batch statistics (conv_tc.cu epilogue)
    (2, 20, 20, [32], 3, 1, 32, False),
    (2, 20, 20, [16], 3, 1, 16, False),
    (3, 12, 20, [64], 3, 2, 64, False),
]


@pytest.mark.parametrize("case", CONV_BN_CASES)
def test_conv_bn_relu_layer_forward_backward(cuda_dev, case):
    from kp_b200 import engine as E
    N, H, W, Cs, k, s, cout, up = case
    rng = np.random.default_rng(sum(case[:3]) + cout)
    cin = sum(Cs)
    xs = [torch.from_numpy(rng.normal(size=(N, H, W, C)).astype(np.float32)).to(BF) for C in Cs]
    w = torch.from_numpy((rng.normal(size=(k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)).to(BF).float()
    b = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
    gamma = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    beta = torch.from_numpy(rng.normal(0, 0.3, cout).astype(np.float32))
    ctx = _mk_ctx(cuda_dev, {"t/conv2d/kernel": w, "t/conv2d/bias": b}, {"bn": cout})
    ctx.G.p("bn/gamma").copy_(gamma.to(cuda_dev)); ctx.G.p("bn/beta").copy_(beta.to(cuda_dev))
    ctx.S.p("bn/moving_variance").fill_(1.0)
    ctx.params_changed()
    ctx.tape, ctx.train_G, ctx.update_moving = E.Tape(), True, True
    srcs = [x.to(cuda_dev) for x in xs]
    out = E.conv_layer(ctx, srcs, "t/conv2d/kernel", "t/conv2d/bias", k, s, 0, bn="bn", train_mode=True, upsample=up)
    dout = torch.from_numpy(rng.normal(size=tuple(out.shape)).astype(np.float32)).to(BF)
    ctx.tape.set_grad(out, dout.to(cuda_dev))
    tape = ctx.tape
    grads_x = None

    # capture dx before the tape clears: run backward manually
    tape.run_closures()
    grads_x = [tape.grad(sx) for sx in srcs]

    # oracle: fp64 autograd on the same rounded operands
    x64 = [x.double().requires_grad_(True) for x in xs]
    w64 = w.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    g64 = gamma.double().requires_grad_(True)
    be64 = beta.double().requires_grad_(True)
    y = T.conv2d(torch.cat(x64, dim=-1), w64, b64, s, 0)
    z, mm, mv = T.batch_norm(y, g64, be64, torch.zeros(cout, dtype=torch.float64), torch.ones(cout, dtype=torch.float64), True)
    a = torch.relu(z)
    if up:
        a = T.resize_bilinear_legacy(a, 2 * a.shape[1], 2 * a.shape[2])
    a.backward(dout.double())
    _close(out, a.detach(), 1e-2, "forward")
    _close(ctx.S.p("bn/moving_mean"), mm, 2e-3, "moving_mean")
    _close(ctx.S.p("bn/moving_variance"), mv, 2e-3, "moving_variance")
    _close(ctx.G.g("bn/gamma"), g64.grad, 1e-2, "dgamma")
    _close(ctx.G.g("bn/beta"), be64.grad, 1e-2, "dbeta")
    _close(ctx.G.g("t/conv2d/kernel"), w64.grad, 1e-2, "dW")
    for gx, x in zip(grads_x, x64):
        if k == 7:
            continue   # first-layer convention: no input gradient requested in the networks; still produced here
        _close(gx, x.grad, 1.5e-2, "dX")


@pytest.mark.parametrize("act,cout,out_f32", [("relu", 64, False), ("leaky", 32, False), ("none", 40, True),
                                               ("sigmoid_last", 4, True), ("none", 1, True)])
def test_conv_bias_act_layer_forward_backward(cuda_dev, act, cout, out_f32):
    from kp_b200 import engine as E, tapconv as tc
    rng = np.random.default_rng(cout)
    N, H, W, C, k = 2, 16, 16, 64, 3
    x = torch.from_numpy(rng.normal(size=(N, H, W, C)).astype(np.float32)).to(BF)
    w = torch.from_numpy((rng.normal(size=(k, k, C, cout)) / np.sqrt(k * k * C)).astype(np.float32)).to(BF).float()
    b = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
    ctx = _mk_ctx(cuda_dev, {"t/conv2d/kernel": w, "t/conv2d/bias": b})
    ctx.tape, ctx.train_G = E.Tape(), True
    code = {"relu": tc.ACT_RELU, "leaky": tc.ACT_LEAKY, "none": tc.ACT_NONE, "sigmoid_last": tc.ACT_SIGMOID_LAST}[act]
    xd = x.to(cuda_dev)
    y = E.conv_layer(ctx, [xd], "t/conv2d/kernel", "t/conv2d/bias", k, 1, 0, act=code, alpha=0.01, out_f32=out_f32)
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    pre = T.conv2d(x64, w64, b64, 1, 0)
    if act == "relu":
        ref = torch.relu(pre)
    elif act == "leaky":
        ref = T.leaky_relu(pre, 0.01)
    elif act == "sigmoid_last":
        ref = torch.cat([pre[..., :-1], torch.sigmoid(pre[..., -1:])], dim=-1)
    else:
        ref = pre
    _close(y, ref.detach(), 2e-3 if out_f32 else 1e-2, "forward")
    # gradient w.r.t. the PRE-activation for f32-output layers (that is what their consumers hand back, channel-padded)
    cpad = -(-cout // 8) * 8
    if out_f32:
        dpre = torch.from_numpy(rng.normal(size=(N, H, W, cout)).astype(np.float32)).to(BF)
        dy = torch.zeros((N, H, W, cpad), dtype=BF)
        dy[..., :cout] = dpre
        ctx.tape.set_grad(y, dy.to(cuda_dev))
        pre.backward(dpre.double())
    else:
        dout = torch.from_numpy(rng.normal(size=(N, H, W, cout)).astype(np.float32)).to(BF)
        ctx.tape.set_grad(y, dout.to(cuda_dev))
        ref.backward(dout.double())
    ctx.tape.run_closures()
    _close(ctx.G.g("t/conv2d/kernel"), w64.grad, 1e-2, "dW")
    _close(ctx.G.g("t/conv2d/bias"), b64.grad, 1e-2, "dbias")
    _close(ctx.tape.grad(xd), x64.grad, 1.5e-2, "dX")


# Shapes of the stage-1 graph at BASELINE sizes, where the launch heuristics take other branches than at toy sizes
# (pixel-tile choice, channel-tile halving for under-filled launches, split-K, halo kernel, parity views of stride 2):
#   (N, H, W, C_stored, C_real, k, stride, pad, cout, act, out_f32, bias)
SHAPE_CASES = [
    # img_discr: 4x4 s2 with explicit pad 1 + SAME -> pads (2,2)/(2,3); odd (65) and even (34) input classes
    (4, 128, 128, 16, 3, 4, 2, 1, 64, "leaky", False, True),       # conv_0: 128 -> 65
    (4, 65, 65, 64, 64, 4, 2, 1, 128, "leaky", False, True),       # conv_1: 65 -> 34, pad (2,3)
    (4, 34, 34, 128, 128, 4, 2, 1, 256, "leaky", False, True),     # conv_2: 34 -> 18
    (8, 10, 10, 512, 512, 4, 2, 1, 1024, "leaky", False, True),    # conv_4: 10 -> 6 (few tiles, long K)
    (64, 4, 4, 2048, 2048, 3, 1, 1, 1, "none", True, False),       # D_logit on [real; fake]: split-K path
    # BASELINE batch: 32 images at 128x128
    (32, 128, 128, 16, 16, 3, 1, 0, 16, "relu", False, True),      # pose_encoder conv_7_1 geometry (halo kernel)
    (32, 128, 128, 64, 64, 3, 1, 0, 64, "relu", False, True),      # VGG conv1_2 / translator conv_5_1 geometry
    (64, 8, 8, 512, 512, 3, 1, 0, 512, "relu", False, True),       # VGG conv5_x on 64 images: channel-tile halving
    (32, 32, 32, 256, 208, 3, 1, 0, 256, "relu", False, True),     # translator conv_1_0: 208 real of 256 stored channels
    (32, 128, 128, 16, 16, 1, 1, 0, 40, "none", True, True),       # 1x1 head, fp32 logits
]


@pytest.mark.parametrize("case", SHAPE_CASES, ids=lambda c: "N%d_%dx%d_C%d_k%ds%dp%d_to%d" % (c[0], c[1], c[2], c[4], c[5], c[6], c[7], c[8]))
def test_conv_layer_at_graph_shapes(cuda_dev, case):
    """Forward, data gradient, weight gradient and bias gradient of one convolution at the shapes the stage-1 graph runs
    at BASELINE batch (reference geometry: models/networks/__init__.py:141-151 for img_discr, :75-102 translator,
    vgg.py:13-43).  Oracle: torch CPU float32 autograd of tf_ops.conv2d on the same bf16-rounded operands."""
    from kp_b200 import engine as E, tapconv as tc
    N, H, W, Cst, Cre, k, stride, pad, cout, act, out_f32, use_bias = case
    rng = np.random.default_rng(N * 1000 + H + cout)
    x = torch.zeros((N, H, W, Cst), dtype=torch.float32)
    x[..., :Cre] = torch.from_numpy(rng.normal(size=(N, H, W, Cre)).astype(np.float32))
    x = x.to(BF)
    w = torch.from_numpy((rng.normal(size=(k, k, Cre, cout)) / np.sqrt(k * k * Cre)).astype(np.float32)).to(BF).float()
    b = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
    specs = {"t/conv2d/kernel": w}
    if use_bias:
        specs["t/conv2d/bias"] = b
    ctx = _mk_ctx(cuda_dev, specs)
    ctx.tape, ctx.train_G = E.Tape(), True
    code = {"relu": tc.ACT_RELU, "leaky": tc.ACT_LEAKY, "none": tc.ACT_NONE}[act]
    xd = x.to(cuda_dev)
    y = E.conv_layer(ctx, [xd], "t/conv2d/kernel", "t/conv2d/bias" if use_bias else None, k, stride, pad, act=code, alpha=0.01,
                     out_f32=out_f32)
    x32 = x[..., :Cre].float().requires_grad_(True)
    w32 = w.clone().requires_grad_(True)
    b32 = b.clone().requires_grad_(True)
    pre = T.conv2d(x32, w32, b32 if use_bias else None, stride, pad)
    ref = torch.relu(pre) if act == "relu" else T.leaky_relu(pre, 0.01) if act == "leaky" else pre
    assert tuple(y.shape) == tuple(ref.shape)
    _close(y, ref.detach(), 2e-3 if out_f32 else 1e-2, "forward")
    cpad = -(-cout // 8) * 8
    dout = torch.from_numpy(rng.normal(size=tuple(ref.shape)).astype(np.float32)).to(BF)
    dy = dout
    if out_f32:
        dy = torch.zeros(tuple(ref.shape[:3]) + (cpad,), dtype=BF)
        dy[..., :cout] = dout
    ctx.tape.set_grad(y, dy.to(cuda_dev))
    ref.backward(dout.float())
    ctx.tape.run_closures()
    _close(ctx.G.g("t/conv2d/kernel"), w32.grad, 1e-2, "dW")
    if use_bias:
        _close(ctx.G.g("t/conv2d/bias"), b32.grad, 1e-2, "dbias")
    dx = ctx.tape.grad(xd)
    _close(dx[..., :Cre], x32.grad, 1.5e-2, "dX")


def test_gradient_accumulation_two_consumers(cuda_dev):
    """A tensor consumed by two convolutions receives the SUM of both data gradients (epilogue accumulate)."""
    from kp_b200 import engine as E
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.normal(size=(2, 16, 16, 32)).astype(np.float32)).to(BF)
    w1 = torch.from_numpy((rng.normal(size=(3, 3, 32, 32)) / 17).astype(np.float32)).to(BF).float()
    w2 = torch.from_numpy((rng.normal(size=(3, 3, 32, 64)) / 17).astype(np.float32)).to(BF).float()
    ctx = _mk_ctx(cuda_dev, {"a/conv2d/kernel": w1, "b/conv2d/kernel": w2})
    ctx.tape, ctx.train_G = E.Tape(), True
    xd = x.to(cuda_dev)
    y1 = E.conv_layer(ctx, [xd], "a/conv2d/kernel", None, 3, 1, 0)
    y2 = E.conv_layer(ctx, [xd], "b/conv2d/kernel", None, 3, 2, 0)
    d1 = torch.from_numpy(rng.normal(size=tuple(y1.shape)).astype(np.float32)).to(BF)
    d2 = torch.from_numpy(rng.normal(size=tuple(y2.shape)).astype(np.float32)).to(BF)
    ctx.tape.set_grad(y1, d1.to(cuda_dev)); ctx.tape.set_grad(y2, d2.to(cuda_dev))
    ctx.tape.run_closures()
    x64 = x.double().requires_grad_(True)
    (T.conv2d(x64, w1.double(), None, 1, 0) * d1.double()).sum().add((T.conv2d(x64, w2.double(), None, 2, 0) * d2.double()).sum()).backward()
    _close(ctx.tape.grad(xd), x64.grad, 1.5e-2, "accumulated dX")


def test_maxpool_compose_pack_prep_losses(cuda_dev):
    from kp_b200 import ops
    rng = np.random.default_rng(5)
    dev = cuda_dev
    # max pool fwd/bwd (+accumulate)
    x = torch.from_numpy(rng.normal(size=(2, 8, 8, 16)).astype(np.float32)).to(BF)
    x64 = x.double().requires_grad_(True)
    ref = T.max_pool_2x2(x64)
    y = ops.maxpool_fwd(x.to(dev))
    assert torch.equal(y.cpu().double(), ref.detach())
    dy = torch.from_numpy(rng.normal(size=tuple(y.shape)).astype(np.float32)).to(BF)
    ref.backward(dy.double())
    dx = torch.empty_like(x, device=dev)
    ops.maxpool_bwd(dy.to(dev), x.to(dev), dx)
    assert torch.equal(dx.cpu().double(), x64.grad)
    base = torch.from_numpy(rng.normal(size=tuple(x.shape)).astype(np.float32)).to(BF)
    dx2 = base.clone().to(dev)
    ops.maxpool_bwd(dy.to(dev), x.to(dev), dx2, accumulate=True)
    _close(dx2, x64.grad + base.double(), 1e-2, "maxpool accumulate")
    # compose fwd/bwd
    P = (2, 8, 8)
    heads = torch.from_numpy(rng.normal(size=P + (4,)).astype(np.float32))
    heads[..., 3] = torch.sigmoid(heads[..., 3])
    im = torch.from_numpy(rng.uniform(-1, 1, P + (3,)).astype(np.float32))
    final, crude, mask = ops.compose_fwd(heads.to(dev), im.to(dev), clip=False, want_parts=True)
    pre = heads.double().clone()
    pre[..., 3] = torch.logit(heads[..., 3].double())
    pre.requires_grad_(True)
    m = torch.sigmoid(pre[..., 3:4])
    ref = im.double() * m + pre[..., :3] * (1 - m)
    _close(final, ref.detach(), 1e-6, "compose")
    g = torch.from_numpy(rng.normal(size=P + (3,)).astype(np.float32))
    ref.backward(g.double())
    dh = ops.compose_bwd(g.to(dev), heads.to(dev), im.to(dev))
    _close(dh[..., :4], pre.grad, 1e-2, "compose bwd")
    assert dh[..., 4:].abs().max().item() == 0
    fc, cc, _ = ops.compose_fwd(heads.to(dev) * 3, im.to(dev), clip=True, want_parts=True)
    assert fc.abs().max().item() <= 1.0 and cc.abs().max().item() <= 1.0
    # pack / unpack channels
    a = torch.from_numpy(rng.normal(size=(3, 4, 4, 128)).astype(np.float32)).to(BF)
    b_ = torch.from_numpy(rng.normal(size=(3, 4, 4, 40)).astype(np.float32))
    c = torch.from_numpy(rng.normal(size=(3, 4, 4, 40)).astype(np.float32))
    j = ops.pack_channels([a.to(dev), b_.to(dev), c.to(dev)], 256)
    ref = torch.cat([a.float(), b_, c, torch.zeros(3, 4, 4, 48)], dim=-1).to(BF)
    assert torch.equal(j.cpu(), ref)
    da = torch.empty_like(a, device=dev); db = torch.empty_like(b_, device=dev); dc = torch.empty_like(c, device=dev)
    ops.unpack_channels(j, [da, db, dc])
    assert torch.equal(da.cpu(), a) and torch.equal(db.cpu(), b_.to(BF).float()) and torch.equal(dc.cpu(), c.to(BF).float())
    # image prep (VGG preprocessing) and its adjoint
    img = torch.from_numpy(rng.uniform(-1, 1, (2, 8, 8, 3)).astype(np.float32))
    xp = ops.image_prep(img.to(dev), ops.VGG_PREP)
    rgb = (img.double() + 1) / 2 * 255
    ref = torch.stack([rgb[..., 2] - 103.939, rgb[..., 1] - 116.779, rgb[..., 0] - 123.68], dim=-1)
    _close(xp[..., :3], ref, 5e-3, "vgg prep")
    assert xp[..., 3:].abs().max().item() == 0
    gp = torch.from_numpy(rng.normal(size=(2, 8, 8, 16)).astype(np.float32)).to(BF)
    dimg = torch.zeros_like(img, device=dev)
    ops.image_prep_bwd(gp.to(dev), dimg, ops.VGG_PREP)
    ref = torch.stack([gp[..., 2], gp[..., 1], gp[..., 0]], dim=-1).double() * 127.5
    _close(dimg, ref, 1e-6, "prep bwd")
    ops.image_prep_bwd(gp.to(dev), dimg, ops.VGG_PREP, accumulate=True)
    _close(dimg, 2 * ref, 1e-6, "prep bwd accumulate")
    # L1 pair
    fg = torch.from_numpy(rng.normal(size=(2, 4, 4, 64)).astype(np.float32)).to(BF)
    fp = torch.from_numpy(rng.normal(size=(2, 4, 4, 64)).astype(np.float32)).to(BF)
    loss = torch.zeros(1, device=dev)
    d = torch.empty_like(fp, device=dev)
    ops.l1_pair(fg.to(dev), fp.to(dev), 0.2, loss, d)
    fp64 = fp.double().requires_grad_(True)
    ref = 0.2 * (fg.double() - fp64).abs().mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-7
    _close(d, fp64.grad, 1e-2, "l1 grad")
    # BCE with logits
    lg = torch.from_numpy(rng.normal(0, 2, (4, 6, 6, 1)).astype(np.float32))
    for label in (0.0, 1.0):
        loss = torch.zeros(1, device=dev)
        dl = ops.bce_logits(lg.to(dev), label, 1.0, loss, True)
        l64 = lg.double().requires_grad_(True)
        ref = T.sigmoid_cross_entropy_with_logits(l64, torch.full_like(l64, label)).mean()
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 1e-5
        _close(dl[..., :1], l64.grad, 1e-2, "bce grad")
        assert dl[..., 1:].abs().max().item() == 0


def test_adam_tf_matches_oracle(cuda_dev):
    from kp_b200 import ops
    rng = np.random.default_rng(9)
    n = 10007
    p = torch.from_numpy(rng.normal(size=n).astype(np.float32))
    m = torch.zeros(n); v = torch.zeros(n)
    pd, md, vd = p.clone().to(cuda_dev), m.clone().to(cuda_dev), v.clone().to(cuda_dev)
    p64, m64, v64 = p.double(), m.double(), v.double()
    for t in range(1, 4):
        g = torch.from_numpy(rng.normal(size=n).astype(np.float32))
        ops.adam_tf(pd, (g * 4).to(cuda_dev), md, vd, 1e-4, t, grad_scale=0.25)
        p64, m64, v64 = T.adam_tf(p64, g.double(), m64, v64, t, 1e-4)
    assert (pd.cpu().double() - p64).abs().max().item() < 1e-6


def test_detector_inference_matches_oracle(cuda_dev):
    """KeypointModel (BN folded, fused K1) vs the oracle with non-trivial BN statistics: mu within 1e-4."""
    from kp_b200 import models
    from oracle import networks as ON
    from oracle import k1_torch
    rng = np.random.default_rng(0)
    P = ON.init_params(0, dtype=torch.float32, with_vgg=False, bias_scale=0.02)
    ON.randomize_bn(P, 1)
    cfg = {"paths": {"log_dir": "/tmp/kp"}, "model": {"n_pts": 40}, "training": {}}
    km = models.KeypointModel(cfg, device=cuda_dev)
    km.ctx.load_state_dict(P)
    im = torch.from_numpy(rng.uniform(-1, 1, (1, 6, 128, 128, 3)).astype(np.float32))
    km.build({"image": im.to(cuda_dev), "idx": torch.tensor([3]), "len": torch.tensor([6])})
    out = km.run()
    assert out["pts"].shape == (1, 6, 40, 2)
    octx = ON.Ctx(P)
    ref = ON.pose_encoder(octx, im[0], 40, False)
    assert (out["pts"][0].cpu() - ref).abs().max().item() <= 1e-4


def test_pack_weights_kernel_matches_recipe(cuda_dev):
    """kp_pack_weights (one launch) == the numpy recipe the CPU lowering tests pin against the conv oracle (bit-exact
    after the bf16 cast), for forward (concat segments, channel tails, BN-fold scale) and data-gradient layouts."""
    from kp_b200 import conv, tapconv as tc
    rng = np.random.default_rng(21)
    cases = [([(2, 8, 8, 16), (2, 8, 8, 32)], 3, 1, 0, 24), ([(1, 8, 8, 40)], 3, 1, 0, 16), ([(1, 16, 16, 64)], 4, 2, 1, 128),
             ([(1, 8, 8, 16)], 7, 1, 0, 32), ([(1, 8, 8, 128), (1, 8, 8, 40), (1, 8, 8, 40)], 3, 1, 0, 256)]
    for shapes, k, s, pad, cout in cases:
        cin = sum(sh[3] for sh in shapes)
        w = rng.normal(size=(k, k, cin, cout)).astype(np.float32)
        wd = torch.from_numpy(w).to(cuda_dev)
        plan, _ = tc.plan_conv_fwd(shapes, k, s, pad, cout)
        ref = torch.from_numpy(tc.pack_weights_np(plan, w)).to(torch.bfloat16)
        assert torch.equal(conv.pack_weights(plan, wd).cpu(), ref)
        assert torch.equal(conv.pack_weights_torch(plan, wd).cpu(), ref)
        scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        ref_s = torch.from_numpy(tc.pack_weights_np(plan, w * scale.reshape(1, 1, 1, cout))).to(torch.bfloat16)
        assert torch.equal(conv.pack_weights(plan, wd, row_scale=torch.from_numpy(scale).to(cuda_dev)).cpu(), ref_s)
        c0 = 0
        for sh in shapes:
            for p in tc.plan_conv_dgrad(sh, k, s, pad, cout if cout % 8 == 0 else cout, cin_slice=(c0, c0 + sh[3], cin)):
                ref_d = torch.from_numpy(tc.pack_weights_np(p, w)).to(torch.bfloat16)
                assert torch.equal(conv.pack_weights(p, wd).cpu(), ref_d)
            c0 += sh[3]


def test_batched_pack_matches_single_launch_pack(cuda_dev):
    """kp_pack_weights_batch (one launch for a whole job table, in place) == kp_pack_weights per job, bit-exact: forward
    (concat segments, tails, unpadded cin/cout) and data-gradient layouts (vector path cout % 8 == 0 and scalar path)."""
    from kp_b200 import conv, tapconv as tc
    rng = np.random.default_rng(33)
    jobs, refs = [], []
    cases = [([(2, 8, 8, 16), (2, 8, 8, 32)], 3, 1, 0, 24), ([(1, 8, 8, 40)], 3, 1, 0, 16), ([(1, 16, 16, 64)], 4, 2, 1, 128),
             ([(1, 8, 8, 16)], 7, 1, 0, 32), ([(1, 8, 8, 128), (1, 8, 8, 40), (1, 8, 8, 40)], 3, 1, 0, 256),
             ([(1, 8, 8, 64)], 3, 1, 0, 4), ([(1, 8, 8, 2048)], 3, 1, 1, 1), ([(1, 8, 8, 256)], 3, 1, 0, 512)]
    for shapes, k, s, pad, cout in cases:
        cin = sum(sh[3] for sh in shapes)
        wd = torch.from_numpy(rng.normal(size=(k, k, cin, cout)).astype(np.float32)).to(cuda_dev)
        plan, _ = tc.plan_conv_fwd(shapes, k, s, pad, cout)
        plans = [plan]
        c0 = 0
        cpad = (cout + 7) // 8 * 8
        for sh in shapes:
            plans += list(tc.plan_conv_dgrad(sh, k, s, pad, cpad, cin_slice=(c0, c0 + sh[3], cin)))
            c0 += sh[3]
        for pl in plans:
            refs.append(conv.pack_weights(pl, wd))
            jobs.append((pl, wd, torch.full_like(refs[-1], float("nan"))))
    table = conv.PackTable(jobs, cuda_dev)
    table.run()
    torch.cuda.synchronize()
    for (pl, _, out), ref in zip(jobs, refs):
        assert torch.equal(out.view(torch.int16), ref.view(torch.int16)), (pl.pack["mode"], pl.rows_pad, pl.Ktot)


@pytest.mark.parametrize("k,cpad,cout,with_bn", [(7, 32, 32, True), (3, 16, 64, False)])
def test_w_unrolled_first_layer_matches_square_conv(cuda_dev, k, cpad, cout, with_bn):
    """The 3-channel first layers run as k x 1 convolutions over a W-unrolled image (kp_image_prep_unrolled): forward,
    weight gradient (written straight into the [k,k,3,cout] variable) and image gradient must equal the k x k conv."""
    from kp_b200 import engine as E, ops, tapconv as tc
    rng = np.random.default_rng(k)
    N, H, W = 2, 24, 20
    img = torch.from_numpy(rng.uniform(-1, 1, (N, H, W, 3)).astype(np.float32))
    w = torch.from_numpy((rng.normal(size=(k, k, 3, cout)) / np.sqrt(k * k * 3)).astype(np.float32)).to(BF).float()
    b = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
    ctx = _mk_ctx(cuda_dev, {"t/conv2d/kernel": w, "t/conv2d/bias": b}, {"bn": cout} if with_bn else None)
    if with_bn:
        ctx.G.p("bn/gamma").fill_(1.0)
        ctx.S.p("bn/moving_variance").fill_(1.0)
    ctx.tape, ctx.train_G = E.Tape(), True
    prep = ops.VGG_PREP if not with_bn else ops.IDENT_PREP
    imd = img.to(cuda_dev)
    xp = ops.image_prep_unrolled(imd, k, (k - 1) // 2, cpad, prep)
    y = E.conv_layer(ctx, [xp], "t/conv2d/kernel", "t/conv2d/bias", (k, 1), 1, 0, bn="bn" if with_bn else None,
                     train_mode=with_bn, act=tc.ACT_NONE if with_bn else tc.ACT_RELU, wshape=(k, 1, 3 * k, cout))
    # oracle: the square conv on the bf16-rounded preprocessed image
    a_, b_, perm = prep
    pre_img = torch.stack([img[..., perm[c]] * a_[c] + b_[c] for c in range(3)], dim=-1).to(BF).double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    z = T.conv2d(pre_img, w64, b.double(), 1, 0)
    if with_bn:
        z, _, _ = T.batch_norm(z, torch.ones(cout, dtype=torch.float64), torch.zeros(cout, dtype=torch.float64),
                               torch.zeros(cout, dtype=torch.float64), torch.ones(cout, dtype=torch.float64), True)
    ref = torch.relu(z)
    _close(y, ref.detach(), 1e-2, "forward")
    dout = torch.from_numpy(rng.normal(size=tuple(y.shape)).astype(np.float32)).to(BF)
    ctx.tape.set_grad(y, dout.to(cuda_dev))
    ref.backward(dout.double())
    ctx.tape.run_closures()
    _close(ctx.G.g("t/conv2d/kernel"), w64.grad, 1e-2, "dW")
    # image gradient through the adjoint of the unrolling (and of the affine preprocessing)
    g = ctx.tape.grad(xp)
    dimg = torch.zeros_like(imd)
    ops.image_prep_unrolled_bwd(g, dimg, k, (k - 1) // 2, prep)
    ref_dimg = torch.zeros_like(img, dtype=torch.float64)
    for c in range(3):
        ref_dimg[..., perm[c]] += pre_img.grad[..., c] * a_[c]
    _close(dimg, ref_dimg, 1.5e-2, "d_image")


HALO_CASES = [
    # (N,H,W,[C],k,cout,out_f32,stats)  - every branch of csrc/conv_halo.cu
    (2, 32, 32, [64], 3, 64, False, False),        # halves=2, weights resident
    (2, 32, 32, [256], 3, 256, False, True),       # streamed weight stages, two channel tiles, BN statistics
    (2, 16, 16, [128], 3, 384, False, False),      # halves=1, BN=128 x 3 channel tiles
    (3, 24, 20, [16, 32], 3, 16, False, True),     # ragged H/W, virtual concat, statistics with invalid pixels
    (1, 32, 32, [8], 3, 8, False, False),          # 8-channel slot (zero-filled second plane), dgrad of a 1-ch head
    (2, 32, 24, [40], 3, 40, True, False),         # 32+8 channel slots, fp32 output, partial last chunk
    (1, 64, 64, [32], (7, 1), 32, False, False),   # 7x1 taps (W-unrolled first layer), 38-row halo
    (2, 128, 128, [16], 3, 16, False, False),      # the detector's 128x128 16-channel layers
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_halo_tile_kernel_matches_oracle_and_tap_kernel(cuda_dev, case, monkeypatch):
    """The halo-tile kernel is normally chosen for narrow outputs only; force it for every eligible shape and compare
    with the fp64 conv oracle (<= 1e-2 of max|ref|, the bf16 conv tolerance) and with the TMA-tap kernel on the same
    operands (both accumulate bf16 products in fp32; a different summation order can flip the bf16 rounding of an output by
    one ulp = 2^-7 relative: <= 8e-3 of max|ref|)."""
    from kp_b200 import conv, tapconv as tc
    N, H, W, Cs, k, cout, out_f32, stats = case
    kh, kw = (k, k) if isinstance(k, int) else k
    rng = np.random.default_rng(N * 1000 + H + cout)
    cin = sum(Cs)
    xs = [torch.from_numpy(rng.normal(size=(N, H, W, C)).astype(np.float32)).to(BF) for C in Cs]
    w = torch.from_numpy((rng.normal(size=(kh, kw, cin, cout)) / np.sqrt(kh * kw * cin)).astype(np.float32)).to(BF).float()
    b = torch.from_numpy(rng.normal(0, 0.2, cout).astype(np.float32))
    plan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape) for x in xs], k, 1, 0, cout)
    srcs = [x.to(cuda_dev) for x in xs]
    wp = conv.pack_weights(plan, w.to(cuda_dev))
    bias = conv.pad_vec(b.to(cuda_dev), plan.rows_pad)
    odt = torch.float32 if out_f32 else BF

    def run(halo):
        monkeypatch.setenv("KP_TAPCONV_HALO2", "0")          # this test pins the cp.async halo kernel (csrc/conv_halo.cu)
        monkeypatch.setenv("KP_TAPCONV_HALO", "1" if halo else "0")
        monkeypatch.setenv("KP_HALO_MAX_COUT", "4096")
        out = torch.full((n, ho, wo, cout), float("nan"), device=cuda_dev, dtype=odt)
        st = (torch.zeros(plan.rows_pad, device=cuda_dev), torch.zeros(plan.rows_pad, device=cuda_dev)) if stats else None
        conv.run_plan(plan, srcs, wp, None if stats else bias, out, act=tc.ACT_NONE if stats else tc.ACT_RELU, stats=st)
        torch.cuda.synchronize()
        return out.float().cpu(), (None if st is None else (st[0].cpu(), st[1].cpu()))
    y_halo, st_halo = run(True)
    y_tap, st_tap = run(False)
    ref = T.conv2d(torch.cat([x.double() for x in xs], dim=-1), w.double(), None if stats else b.double(), 1, 0)
    if not stats:
        ref = torch.relu(ref)
    scale = ref.abs().max().item()
    assert torch.isfinite(y_halo).all()
    assert (y_halo.double() - ref).abs().max().item() <= 1e-2 * scale
    assert (y_halo - y_tap).abs().max().item() <= 8e-3 * scale
    if stats:
        s_ref, q_ref = ref.sum(dim=(0, 1, 2)), (ref * ref).sum(dim=(0, 1, 2))
        for got in (st_halo, st_tap):
            assert (got[0][:cout].double() - s_ref).abs().max().item() <= 2e-3 * (s_ref.abs().max().item() + ref.abs().sum(dim=(0, 1, 2)).max().item() * 1e-2)
            assert (got[1][:cout].double() - q_ref).abs().max().item() <= 2e-3 * q_ref.abs().max().item()

HALO2_CASES = [
    # (N,H,W,[C],k,cout,out_f32,stats)  - every branch of csrc/conv_halo2.cu (TMA-staged, swizzled halo)
    (2, 32, 32, [64], 3, 64, False, False),        # 128-byte rows, halves=2, weights resident, one CTA per SM
    (2, 32, 32, [32], 3, 32, False, True),         # 64-byte rows, two CTAs per SM, BN statistics
    (2, 32, 32, [16], 3, 16, False, False),        # 32-byte rows
    (2, 32, 32, [256], 3, 256, False, True),       # four slots, streamed weight stages, two channel tiles
    (2, 16, 16, [128], 3, 384, False, False),      # halves=1, BN=128 x 3 channel tiles
    (3, 24, 20, [16, 32], 3, 16, False, True),     # ragged H/W, virtual concat of a 32-byte and a 64-byte source
    (2, 40, 24, [96], 3, 48, True, False),         # 64+32 channel slots of ONE source (two tensor maps), fp32 output
    (1, 64, 64, [32], (7, 1), 32, False, False),   # 7x1 taps (W-unrolled first layer), 38-row halo, pitch 8
    (1, 64, 64, [16], (3, 1), 64, False, False),   # 3x1 taps (VGG conv1_1 layout)
    (2, 128, 128, [16], 3, 16, False, False),      # the detector's 128x128 16-channel layers
    (2, 64, 64, [128], 3, 128, False, False),      # translator 64x64: streamed weights, 512 TMEM columns, one CTA per SM
    (4, 128, 128, [64, 64], 3, 32, False, True),   # pose_encoder conv_5_0-like concat at 128x128
]


@pytest.mark.parametrize("case", HALO2_CASES)
def test_tma_halo_kernel_matches_oracle_and_tap_kernel(cuda_dev, case, monkeypatch):
    """csrc/conv_halo2.cu: the input halo arrives by ONE swizzled TMA box per channel slot and every tap is a start-address
    offset of the shared-memory descriptor.  Forced on for every eligible shape (KP_TAPCONV_HALO2=2) and compared with the
    fp64 conv oracle (<= 1e-2 of max|ref|) and with the TMA-tap kernel on the same operands (<= one bf16 ulp of max|ref|)."""
    from kp_b200 import conv, tapconv as tc
    N, H, W, Cs, k, cout, out_f32, stats = case
    kh, kw = (k, k) if isinstance(k, int) else k
    rng = np.random.default_rng(N * 1000 + H + cout)
    cin = sum(Cs)
    xs = [torch.from_numpy(rng.normal(size=(N, H, W, C)).astype(np.float32)).to(BF) for C in Cs]
    w = torch.from_numpy((rng.normal(size=(kh, kw, cin, cout)) / np.sqrt(kh * kw * cin)).astype(np.float32)).to(BF).float()
    b = torch.from_numpy(rng.normal(0, 0.2, cout).astype(np.float32))
    plan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape) for x in xs], k, 1, 0, cout)
    srcs = [x.to(cuda_dev) for x in xs]
    wp = conv.pack_weights(plan, w.to(cuda_dev))
    bias = conv.pad_vec(b.to(cuda_dev), plan.rows_pad)
    odt = torch.float32 if out_f32 else BF
    from kp_b200 import _lib
    lib = _lib.load()

    def run(halo2):
        monkeypatch.setenv("KP_TAPCONV_HALO2", "2" if halo2 else "0")
        monkeypatch.setenv("KP_TAPCONV_HALO", "0")
        out = torch.full((n, ho, wo, cout), float("nan"), device=cuda_dev, dtype=odt)
        st = (torch.zeros(plan.rows_pad, device=cuda_dev), torch.zeros(plan.rows_pad, device=cuda_dev)) if stats else None
        conv.run_plan(plan, srcs, wp, None if stats else bias, out, act=tc.ACT_NONE if stats else tc.ACT_RELU, stats=st)
        torch.cuda.synchronize()
        return out.float().cpu(), (None if st is None else (st[0].cpu(), st[1].cpu()))
    y_halo, st_halo = run(True)
    y_tap, st_tap = run(False)
    ref = T.conv2d(torch.cat([x.double() for x in xs], dim=-1), w.double(), None if stats else b.double(), 1, 0)
    if not stats:
        ref = torch.relu(ref)
    scale = ref.abs().max().item()
    assert torch.isfinite(y_halo).all()
    assert (y_halo.double() - ref).abs().max().item() <= 1e-2 * scale
    assert (y_halo - y_tap).abs().max().item() <= 8e-3 * scale
    if stats:
        s_ref, q_ref = ref.sum(dim=(0, 1, 2)), (ref * ref).sum(dim=(0, 1, 2))
        for got in (st_halo, st_tap):
            assert (got[0][:cout].double() - s_ref).abs().max().item() <= 2e-3 * (s_ref.abs().max().item() + ref.abs().sum(dim=(0, 1, 2)).max().item() * 1e-2)
            assert (got[1][:cout].double() - q_ref).abs().max().item() <= 2e-3 * q_ref.abs().max().item()

WGRAD2_CASES = [
    # (N,H,W,Cx,c0,cin_total,cout,k)  - every branch of csrc/conv_wgrad2.cu (halo-tile weight gradient)
    (2, 32, 32, 128, 0, 128, 128, 3),      # two 64-channel slots per tap, 9 accumulators of 128 columns in 3 tap groups
    (2, 32, 32, 256, 0, 256, 256, 3),      # two ci blocks x two co blocks
    (2, 32, 32, 64, 0, 64, 64, 3),         # Cin 64: two taps of a kernel row per M=128 MMA (LBO = one pixel row)
    (2, 32, 32, 32, 0, 32, 32, 3),         # Cin 32: a whole 3-wide kernel row per MMA, 64-byte swizzle
    (2, 32, 32, 16, 0, 16, 16, 3),         # Cin 16: 32-byte swizzle
    (3, 24, 20, 32, 0, 32, 16, 3),         # ragged H/W (zero-filled tile borders), Cout 16
    (2, 64, 64, 32, 0, 32, 32, (7, 1)),    # 7x1 taps (W-unrolled first layer): one tap per MMA group
    (2, 32, 32, 16, 0, 16, 64, (3, 1)),    # VGG conv1_1 layout
    (2, 16, 16, 128, 0, 128, 64, 3),       # one tile row of tiles, BN 64
    (2, 32, 32, 64, 64, 128, 32, 3),       # second source of a virtual concat: rows [64,128) of a 128-row kernel
    (4, 128, 128, 64, 0, 64, 64, 3),       # 128x128 geometry, many pixel splits
]


@pytest.mark.parametrize("case", WGRAD2_CASES)
def test_halo_weight_gradient_matches_oracle_and_tap_kernel(cuda_dev, case, monkeypatch):
    """csrc/conv_wgrad2.cu against the fp64 oracle (autograd of tf_ops.conv2d on the same bf16 operands) and against the
    per-tap kernel (csrc/conv_wgrad.cu, KP_WGRAD_HALO=0)."""
    from kp_b200 import conv, tapconv as tc
    N, H, W, Cx, c0, cin_total, cout, k = case
    kh, kw = (k, k) if isinstance(k, int) else k
    rng = np.random.default_rng(N * 100 + H + Cx + cout)
    x = torch.from_numpy(rng.normal(size=(N, H, W, Cx)).astype(np.float32)).to(BF)
    dy = torch.from_numpy(rng.normal(size=(N, H, W, cout)).astype(np.float32)).to(BF)
    plan = tc.plan_conv_wgrad((N, H, W, Cx), k, 1, 0, cout, cin_slice=(c0, c0 + Cx, cin_total))
    xd, dyd = x.to(cuda_dev), dy.to(cuda_dev)

    def run(halo):
        monkeypatch.setenv("KP_WGRAD_HALO", "2" if halo else "0")
        dw = torch.zeros((kh, kw, cin_total, cout), device=cuda_dev)
        conv.run_wgrad(plan, xd, dyd, dw)
        torch.cuda.synchronize()
        return dw.cpu()
    got, old = run(True), run(False)
    w64 = torch.zeros((kh, kw, Cx, cout), dtype=torch.float64, requires_grad=True)
    y = T.conv2d(x.double(), w64, None, 1, 0)
    y.backward(dy.double())
    ref = torch.zeros((kh, kw, cin_total, cout), dtype=torch.float64)
    ref[:, :, c0:c0 + Cx] = w64.grad
    scale = ref.abs().max().item()
    assert torch.isfinite(got).all()
    assert (got.double() - ref).abs().max().item() <= 2e-3 * scale, (got.double() - ref).abs().max().item() / scale
    assert (got - old).abs().max().item() <= 2e-3 * scale
    untouched = torch.ones(cin_total, dtype=torch.bool)
    untouched[c0:c0 + Cx] = False
    assert got[:, :, untouched].abs().max().item() == 0.0 if untouched.any() else True


@pytest.mark.parametrize("geom", [(2, 32, 32, 16, 32, 16, 3), (2, 16, 16, 64, 128, 64, 3), (3, 24, 20, 32, 64, 8, 1)])
def test_two_layer_chain_forward_backward(cuda_dev, geom):
    """conv1 + BN + ReLU -> conv2 + BN + ReLU through the tape: every gradient against the fp64 autograd oracle on the same
    bf16-rounded operands (the data gradient of layer 2 feeds the batch-norm backward of layer 1)."""
    from kp_b200 import engine as E
    N, H, W, C0, C1, C2, k2 = geom
    rng = np.random.default_rng(C0 + C1 + C2)
    x = torch.from_numpy(rng.normal(size=(N, H, W, C0)).astype(np.float32)).to(BF)
    w1 = torch.from_numpy((rng.normal(size=(3, 3, C0, C1)) / np.sqrt(9 * C0)).astype(np.float32)).to(BF).float()
    w2 = torch.from_numpy((rng.normal(size=(k2, k2, C1, C2)) / np.sqrt(k2 * k2 * C1)).astype(np.float32)).to(BF).float()
    b1 = torch.from_numpy(rng.normal(0, 0.1, C1).astype(np.float32))
    b2 = torch.from_numpy(rng.normal(0, 0.1, C2).astype(np.float32))
    g1 = torch.from_numpy(rng.uniform(0.5, 1.5, C1).astype(np.float32)); be1 = torch.from_numpy(rng.normal(0, 0.3, C1).astype(np.float32))
    g2 = torch.from_numpy(rng.uniform(0.5, 1.5, C2).astype(np.float32)); be2 = torch.from_numpy(rng.normal(0, 0.3, C2).astype(np.float32))
    ctx = _mk_ctx(cuda_dev, {"a/conv2d/kernel": w1, "a/conv2d/bias": b1, "b/conv2d/kernel": w2, "b/conv2d/bias": b2},
                  {"bna": C1, "bnb": C2})
    for n_, v in (("bna/gamma", g1), ("bna/beta", be1), ("bnb/gamma", g2), ("bnb/beta", be2)):
        ctx.G.p(n_).copy_(v.to(cuda_dev))
    ctx.S.p("bna/moving_variance").fill_(1.0); ctx.S.p("bnb/moving_variance").fill_(1.0)
    ctx.params_changed()
    ctx.begin_run()
    ctx.tape, ctx.train_G = E.Tape(), True
    xd = x.to(cuda_dev)
    h1 = E.conv_layer(ctx, [xd], "a/conv2d/kernel", "a/conv2d/bias", 3, 1, 0, bn="bna", train_mode=True)
    h2 = E.conv_layer(ctx, [h1], "b/conv2d/kernel", "b/conv2d/bias", k2, 1, 0, bn="bnb", train_mode=True)
    dout = torch.from_numpy(rng.normal(size=tuple(h2.shape)).astype(np.float32)).to(BF)
    ctx.tape.set_grad(h2, dout.to(cuda_dev))
    tape = ctx.tape
    tape.run_closures()
    dx = tape.grad(xd)
    # oracle
    x64 = x.double().requires_grad_(True)
    P = {n_: v.double().requires_grad_(True) for n_, v in (("w1", w1), ("w2", w2), ("b1", b1), ("b2", b2), ("g1", g1), ("be1", be1),
                                                           ("g2", g2), ("be2", be2))}
    y1 = T.conv2d(x64, P["w1"], P["b1"], 1, 0)
    z1, _, _ = T.batch_norm(y1, P["g1"], P["be1"], torch.zeros(C1, dtype=torch.float64), torch.ones(C1, dtype=torch.float64), True)
    a1 = torch.relu(z1)
    y2 = T.conv2d(a1, P["w2"], P["b2"], 1, 0)
    z2, _, _ = T.batch_norm(y2, P["g2"], P["be2"], torch.zeros(C2, dtype=torch.float64), torch.ones(C2, dtype=torch.float64), True)
    a2 = torch.relu(z2)
    a2.backward(dout.double())
    _close(h2, a2.detach(), 2e-2, "forward")
    # layer-1 quantities sit behind two ReLU masks and two batch-statistics renormalisations: 1.5x the single-layer budget
    _close(ctx.G.g("bna/gamma"), P["g1"].grad, 1.5e-2, "dgamma layer 1")
    _close(ctx.G.g("bna/beta"), P["be1"].grad, 1.5e-2, "dbeta layer 1")
    _close(ctx.G.g("bnb/gamma"), P["g2"].grad, 1e-2, "dgamma layer 2")
    _close(ctx.G.g("a/conv2d/kernel"), P["w1"].grad, 1.5e-2, "dW layer 1")
    _close(ctx.G.g("b/conv2d/kernel"), P["w2"].grad, 1e-2, "dW layer 2")
    _close(dx, x64.grad, 1.5e-2, "dX")


@pytest.mark.parametrize("C,train", [(16, True), (64, True), (24, True), (32, False)])
def test_standalone_batch_norm_layer(cuda_dev, C, train):
    """networks.layers.batch_norm (reference layers.py:13-14) as a stand-alone op: batch statistics from kp_channel_sum /
    kp_channel_sumsq (no torch reductions), biased variance for the normalisation, unbiased for the moving average."""
    from kp_b200 import networks
    from kp_b200.networks import layers
    rng = np.random.default_rng(C)
    N, H, W = 3, 10, 12
    x = torch.from_numpy(rng.normal(0.5, 2.0, (N, H, W, C)).astype(np.float32)).to(BF)
    gamma = torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32))
    beta = torch.from_numpy(rng.normal(0, 0.3, C).astype(np.float32))
    mm0 = torch.from_numpy(rng.normal(0, 0.2, C).astype(np.float32))
    mv0 = torch.from_numpy(rng.uniform(0.5, 2.0, C).astype(np.float32))
    ctx = _mk_ctx(cuda_dev, {}, bn={"bn": C})
    ctx.G.p("bn/gamma").copy_(gamma); ctx.G.p("bn/beta").copy_(beta)
    ctx.S.p("bn/moving_mean").copy_(mm0); ctx.S.p("bn/moving_variance").copy_(mv0)
    ctx.update_moving = True
    networks.set_context(ctx)
    y = layers.batch_norm(x.to(cuda_dev), train, scope="bn")
    torch.cuda.synchronize()
    xd = x.double()
    if train:
        mean, var = xd.mean(dim=(0, 1, 2)), xd.var(dim=(0, 1, 2), unbiased=False)
        n = N * H * W
        assert torch.allclose(ctx.S.p("bn/moving_mean").cpu().double(), mm0.double() * 0.999 + mean * 0.001, atol=1e-5)
        assert torch.allclose(ctx.S.p("bn/moving_variance").cpu().double(), mv0.double() * 0.999 + var * n / (n - 1) * 0.001,
                              atol=1e-5)
    else:
        mean, var = mm0.double(), mv0.double()
    ref = (xd - mean) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    _close(y, ref, 1e-2, "batch_norm out")
