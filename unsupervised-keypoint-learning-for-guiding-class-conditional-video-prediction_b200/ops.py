"""Thin torch-facing wrappers of the memory-bound C-ABI entry points (include/kp_b200.h).  No arithmetic here."""
import ctypes

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v])


def _i3(v):
    return (ctypes.c_int * 3)(*[int(x) for x in v])


IDENT_PREP = ((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (0, 1, 2))
# (x+1)/2*255 then RGB->BGR minus VGG_MEAN (detector_translator_model.py:262-263, vgg.py:15-19)
VGG_MEAN = (103.939, 116.779, 123.68)
VGG_PREP = ((127.5, 127.5, 127.5), tuple(127.5 - m for m in VGG_MEAN), (2, 1, 0))


def image_prep(x, prep=IDENT_PREP):
    """f32 [..., 3] -> bf16 [..., 16]."""
    assert x.dtype == F32 and x.shape[-1] == 3 and x.is_cuda
    x = x.contiguous()
    out = torch.empty(tuple(x.shape[:-1]) + (16,), device=x.device, dtype=BF16)
    a, b, perm = prep
    _lib.call("kp_image_prep", _p(x), x.numel() // 3, _f3(a), _f3(b), _i3(perm), _p(out), _st())
    return out


def image_prep_bwd(g, dx, prep=IDENT_PREP, accumulate=False):
    a, _, perm = prep
    _lib.call("kp_image_prep_bwd", _p(g), g.numel() // 16, _f3(a), _i3(perm), 1 if accumulate else 0, _p(dx), _st())
    return dx


def image_prep_unrolled(x, kw, pad_left, cpad, prep=IDENT_PREP, out=None):
    """f32 [N,H,W,3] -> bf16 [N,H,W,cpad] with out[..., j*3+c] = prep(x[..., w+j-pad_left, c]) (zero outside): turns a
    KHxKW first-layer convolution into a KHx1 one over KW*3 channels.  `out`: optional destination (a batch slice of a
    larger buffer)."""
    assert x.dtype == F32 and x.dim() == 4 and x.shape[-1] == 3 and x.is_cuda
    x = x.contiguous()
    N, H, W, _ = x.shape
    if out is None:
        out = torch.empty((N, H, W, cpad), device=x.device, dtype=BF16)
    assert out.is_contiguous() and tuple(out.shape) == (N, H, W, cpad) and out.dtype == BF16
    a, b, perm = prep
    _lib.call("kp_image_prep_unrolled", _p(x), N, H, W, int(kw), int(pad_left), int(cpad), _f3(a), _f3(b), _i3(perm), _p(out),
              _st())
    return out


def image_prep_unrolled_bwd(g, dx, kw, pad_left, prep=IDENT_PREP, accumulate=False):
    a, _, perm = prep
    N, H, W, cpad = g.shape
    _lib.call("kp_image_prep_unrolled_bwd", _p(g), N, H, W, int(kw), int(pad_left), int(cpad), _f3(a), _i3(perm),
              1 if accumulate else 0, _p(dx), _st())
    return dx


def bn_finalize(ssum, ssq, bias, gamma, beta, count, moving_mean, moving_var, eps=1e-5, decay=0.999):
    C = gamma.shape[0]
    dev = gamma.device
    scale = torch.empty(C, device=dev, dtype=F32)
    shift = torch.empty(C, device=dev, dtype=F32)
    mean = torch.empty(C, device=dev, dtype=F32)
    rstd = torch.empty(C, device=dev, dtype=F32)
    _lib.call("kp_bn_finalize", _p(ssum), _p(ssq), _p(bias), _p(gamma), _p(beta), C, float(count), eps, decay,
              _p(moving_mean), _p(moving_var), _p(scale), _p(shift), _p(mean), _p(rstd), _st())
    return scale, shift, mean, rstd


def bn_stats_apply(x, ssum, ssq, bias, gamma, beta, count, moving_mean, moving_var, relu=True, upsample=False, eps=1e-5,
                   decay=0.999, segments=1):
    """bn_finalize + bn_act_apply in one launch -> (out, scale, shift, mean, rstd).  segments > 1: the batch is that many
    equal runs of images normalised with their own statistics (ssum / ssq and the four returned vectors are
    [segments*C]; `count` = pixels of ONE segment)."""
    N, H, W, C = x.shape
    dev = x.device
    par = torch.empty((4, segments * C), device=dev, dtype=F32)
    out = torch.empty((N, 2 * H, 2 * W, C) if upsample else (N, H, W, C), device=dev, dtype=BF16)
    _lib.call("kp_bn_stats_apply", _p(ssum), _p(ssq), _p(bias), _p(gamma), _p(beta), float(count), eps, decay,
              _p(moving_mean), _p(moving_var), _p(par[0]), _p(par[1]), _p(par[2]), _p(par[3]), _p(x), 1 if relu else 0,
              1 if upsample else 0, N, H, W, C, _p(out), segments, _st())
    return out, par[0], par[1], par[2], par[3]


def bn_act_apply(x, scale, shift, relu=True, upsample=False):
    N, H, W, C = x.shape
    out = torch.empty((N, 2 * H, 2 * W, C) if upsample else (N, H, W, C), device=x.device, dtype=BF16)
    _lib.call("kp_bn_act_apply", _p(x), _p(scale), _p(shift), 1 if relu else 0, 1 if upsample else 0, N, H, W, C, _p(out),
              _st())
    return out


def bn_act_bwd(dout, x, scale, shift, mean, rstd, relu=True, upsample=False, gbeta_acc=None, ggamma_acc=None,
               zeroed=None, segments=1):
    """`zeroed`: optional pre-zeroed f32 buffer of >= 2*segments*C elements for this call's dbeta/dgamma sums (saves two
    memsets); gbeta_acc / ggamma_acc: parameter-gradient views that receive += dbeta / dgamma inside the kernel."""
    N, H, W, C = x.shape
    SC = segments * C
    if zeroed is not None:
        dbeta, dgamma, pre = zeroed[:SC], zeroed[SC:2 * SC], 1
    else:
        dbeta = torch.empty(SC, device=x.device, dtype=F32)
        dgamma = torch.empty(SC, device=x.device, dtype=F32)
        pre = 0
    dx = torch.empty_like(x)
    if upsample:
        # adjoint of the x2 resize once (kp_upsample2x_bwd), then the plain backward: the fused variant gathered the 3x3
        # neighbourhood of dout in BOTH of its passes
        dact = torch.empty_like(x)
        _lib.call("kp_upsample2x_bwd", _p(dout), N, H, W, C, _p(dact), _st())
        dout = dact
    _lib.call("kp_bn_act_bwd", _p(dout), _p(x), _p(scale), _p(shift), _p(mean), _p(rstd), 1 if relu else 0,
              0, N, H, W, C, _p(dbeta), _p(dgamma), _p(dx), _p(gbeta_acc), _p(ggamma_acc), pre, segments, _st())
    return dx, dgamma, dbeta


def conv1x1_f32(x, w, bias):
    """x bf16 [..., Cin] (Cin = 16), w f32 [1,1,Cin,Cout] / [Cin,Cout], bias f32 [Cout] or None -> f32 [..., Cout]."""
    cin, cout = w.shape[-2], w.shape[-1]
    out = torch.empty(tuple(x.shape[:-1]) + (cout,), device=x.device, dtype=F32)
    _lib.call("kp_conv1x1_f32", _p(x), _p(w), _p(bias), x.numel() // cin, cin, cout, _p(out), _st())
    return out


def act_mask_bwd(dy, y, alpha):
    g = torch.empty_like(y)
    _lib.call("kp_act_mask_bwd", _p(dy), _p(y), float(alpha), y.numel(), _p(g), _st())
    return g


def maxpool_fwd(x):
    N, H, W, C = x.shape
    out = torch.empty((N, H // 2, W // 2, C), device=x.device, dtype=BF16)
    _lib.call("kp_maxpool2x2_fwd", _p(x), N, H, W, C, _p(out), _st())
    return out


def maxpool_bwd(dy, x, dx, relu_mask=False, accumulate=False):
    N, H, W, C = x.shape
    _lib.call("kp_maxpool2x2_bwd", _p(dy), _p(x), (1 if relu_mask else 0) | (2 if accumulate else 0), N, H, W, C, _p(dx),
              _st())
    return dx


def compose_fwd(heads, im, clip=False, want_parts=False):
    P = im.numel() // 3
    final = torch.empty_like(im)
    crude = torch.empty_like(im) if want_parts else None
    mask = torch.empty(tuple(im.shape[:-1]) + (1,), device=im.device, dtype=F32) if want_parts else None
    _lib.call("kp_mask_compose_fwd", _p(heads), _p(im), P, 1 if clip else 0, _p(final), _p(crude), _p(mask), _st())
    return final, crude, mask


def compose_bwd(d_final, heads, im):
    P = im.numel() // 3
    out = torch.empty(tuple(im.shape[:-1]) + (16,), device=im.device, dtype=BF16)
    _lib.call("kp_mask_compose_bwd", _p(d_final), _p(heads), _p(im), P, _p(out), _st())
    return out


def pack_channels(srcs, Ctot):
    """cat(srcs, -1) cast to bf16, zero padded to Ctot channels."""
    n = len(srcs)
    srcs = [s.contiguous() for s in srcs]
    P = srcs[0].numel() // srcs[0].shape[-1]
    out = torch.empty(tuple(srcs[0].shape[:-1]) + (Ctot,), device=srcs[0].device, dtype=BF16)
    ptrs = (ctypes.c_void_p * 3)(*[s.data_ptr() for s in srcs], *([None] * (3 - n)))
    Cs = (ctypes.c_int * 3)(*[s.shape[-1] for s in srcs], *([0] * (3 - n)))
    f32 = (ctypes.c_int * 3)(*[1 if s.dtype == F32 else 0 for s in srcs], *([0] * (3 - n)))
    _lib.call("kp_pack_channels", ptrs, Cs, f32, n, P, Ctot, _p(out), _st())
    return out


def unpack_channels(g, dsts):
    n = len(dsts)
    Ctot = g.shape[-1]
    P = g.numel() // Ctot
    ptrs = (ctypes.c_void_p * 3)(*[d.data_ptr() for d in dsts], *([None] * (3 - n)))
    Cs = (ctypes.c_int * 3)(*[d.shape[-1] for d in dsts], *([0] * (3 - n)))
    f32 = (ctypes.c_int * 3)(*[1 if d.dtype == F32 else 0 for d in dsts], *([0] * (3 - n)))
    _lib.call("kp_unpack_channels", _p(g), P, Ctot, ptrs, Cs, f32, n, _st())
    return dsts


def l1_pair(feat_gt, feat_pred, weight, loss, d_pred):
    _lib.call("kp_l1_pair_fwd_bwd", _p(feat_gt), _p(feat_pred), feat_pred.numel(), float(weight), _p(loss), _p(d_pred),
              _st())


def bce_logits(logits, label, weight, loss, want_grad):
    n = logits.numel()
    d = torch.empty(tuple(logits.shape[:-1]) + (8,), device=logits.device, dtype=BF16) if want_grad else None
    _lib.call("kp_bce_logits_fwd_bwd", _p(logits), n, float(label), float(weight), _p(loss), _p(d), _st())
    return d


def adam_lr_t(lr, t, beta1=0.5, beta2=0.999):
    """TF's bias-corrected step size lr*sqrt(1-beta2^t)/(1-beta1^t)."""
    import math
    return lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)


def adam_tf(p, g, m, v, lr, t, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0, lr_t_dev=None):
    _lib.call("kp_adam_tf", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), beta1, beta2, eps, int(t), float(grad_scale),
              _p(lr_t_dev), _st())


def channel_sum(g, out, squares=False):
    """out[c] += sum over pixels of g[...,c] (squares: of g[...,c]**2)."""
    C = g.shape[-1]
    _lib.call("kp_channel_sumsq" if squares else "kp_channel_sum", _p(g), g.numel() // C, C, _p(out), _st())
    return out
