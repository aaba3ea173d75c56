"""Torch-facing wrappers (and autograd nodes) for the K1 keypoint kernels.

PyTorch is plumbing here: it owns device memory and the stream; every FLOP happens in libkp_b200.so.
"""
import torch

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_cuda_f32(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (the B200 path has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32 (got %s)" % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims (got shape %s)" % (name, ndim, tuple(t.shape)))
    return t.contiguous()


def softargmax_render_fwd(logits, map_hw=None, inv_std=14.3, want_prob=True):
    """logits [B,H,W,K] -> (mu [B,K,2], prob_x [B,W,K]|None, prob_y [B,H,K]|None, maps [B,h,w,K]|None)."""
    logits = _check_cuda_f32(logits, "logits", 4)
    B, H, W, K = logits.shape
    dev = logits.device
    mu = torch.empty((B, K, 2), device=dev, dtype=torch.float32)
    px = torch.empty((B, W, K), device=dev, dtype=torch.float32) if want_prob else None
    py = torch.empty((B, H, K), device=dev, dtype=torch.float32) if want_prob else None
    maps = None
    hm = wm = 0
    if map_hw is not None:
        hm, wm = int(map_hw[0]), int(map_hw[1])
        maps = torch.empty((B, hm, wm, K), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.call("kp_softargmax_render_fwd", _ptr(logits), B, H, W, K, _ptr(mu), _ptr(px), _ptr(py), _ptr(maps),
                  hm, wm, float(inv_std), _stream())
    return mu, px, py, maps


def softargmax_render_bwd(d_maps, d_mu, mu, px, py, H, W, inv_std=14.3):
    """Gradients on maps [B,h,w,K] and/or mu [B,K,2] -> d_logits [B,H,W,K]."""
    B, K, _ = mu.shape
    dev = mu.device
    hm = wm = 0
    if d_maps is not None:
        d_maps = _check_cuda_f32(d_maps, "d_maps", 4)
        hm, wm = d_maps.shape[1], d_maps.shape[2]
    if d_mu is not None:
        d_mu = _check_cuda_f32(d_mu, "d_mu", 3)
    d_logits = torch.empty((B, H, W, K), device=dev, dtype=torch.float32)
    scratch = torch.empty((B, K, 2), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.call("kp_softargmax_render_bwd", _ptr(d_maps), _ptr(d_mu), _ptr(mu), _ptr(px), _ptr(py), B, H, W, K,
                  hm, wm, float(inv_std), _ptr(d_logits), _ptr(scratch), _stream())
    return d_logits


def render_fwd(mu, shape_hw, inv_std=14.3):
    mu = _check_cuda_f32(mu, "mu", 3)
    B, K, two = mu.shape
    if two != 2:
        raise ValueError("mu must be [B,K,2]")
    h, w = int(shape_hw[0]), int(shape_hw[1])
    maps = torch.empty((B, h, w, K), device=mu.device, dtype=torch.float32)
    with torch.cuda.device(mu.device):
        _lib.call("kp_render_fwd", _ptr(mu), B, K, h, w, float(inv_std), _ptr(maps), _stream())
    return maps


def render_bwd(d_maps, mu, inv_std=14.3):
    d_maps = _check_cuda_f32(d_maps, "d_maps", 4)
    mu = _check_cuda_f32(mu, "mu", 3)
    B, h, w, K = d_maps.shape
    d_mu = torch.empty_like(mu)
    with torch.cuda.device(mu.device):
        _lib.call("kp_render_bwd", _ptr(d_maps), None, _ptr(mu), B, K, h, w, float(inv_std), _ptr(d_mu), _stream())
    return d_mu


def render_colorize(mu, colors, shape_hw, inv_std=14.3):
    mu = _check_cuda_f32(mu, "mu", 3)
    B, K, _ = mu.shape
    colors = torch.as_tensor(colors, dtype=torch.float32, device=mu.device).reshape(K, 3).contiguous()
    h, w = int(shape_hw[0]), int(shape_hw[1])
    out = torch.empty((B, h, w, 3), device=mu.device, dtype=torch.float32)
    with torch.cuda.device(mu.device):
        _lib.call("kp_render_colorize_fwd", _ptr(mu), _ptr(colors), B, K, h, w, float(inv_std), _ptr(out), _stream())
    return out


def colorize(maps, colors):
    maps = _check_cuda_f32(maps, "maps")
    K = maps.shape[-1]
    colors = torch.as_tensor(colors, dtype=torch.float32, device=maps.device).reshape(K, 3).contiguous()
    P = maps.numel() // K
    out = torch.empty(tuple(maps.shape[:-1]) + (3,), device=maps.device, dtype=torch.float32)
    with torch.cuda.device(maps.device):
        _lib.call("kp_colorize_fwd", _ptr(maps), _ptr(colors), P, K, _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------------------------
# autograd nodes
# ----------------------------------------------------------------------------------------------
class SoftArgmaxRender(torch.autograd.Function):
    """logits -> (mu, maps).  One kernel forward, one kernel backward."""

    @staticmethod
    def forward(ctx, logits, map_h, map_w, inv_std):
        mu, px, py, maps = softargmax_render_fwd(logits, (map_h, map_w), inv_std, want_prob=True)
        ctx.save_for_backward(mu, px, py)
        ctx.hw = (logits.shape[1], logits.shape[2])
        ctx.inv_std = inv_std
        return mu, maps

    @staticmethod
    def backward(ctx, d_mu, d_maps):
        mu, px, py = ctx.saved_tensors
        H, W = ctx.hw
        d_logits = softargmax_render_bwd(d_maps, d_mu, mu, px, py, H, W, ctx.inv_std)
        return d_logits, None, None, None


class SoftArgmax(torch.autograd.Function):
    """logits -> mu (get_coord x2 + stack)."""

    @staticmethod
    def forward(ctx, logits):
        mu, px, py, _ = softargmax_render_fwd(logits, None, want_prob=True)
        ctx.save_for_backward(mu, px, py)
        ctx.hw = (logits.shape[1], logits.shape[2])
        return mu

    @staticmethod
    def backward(ctx, d_mu):
        mu, px, py = ctx.saved_tensors
        H, W = ctx.hw
        return softargmax_render_bwd(None, d_mu, mu, px, py, H, W)


class GaussianMaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, h, w, inv_std):
        ctx.save_for_backward(mu)
        ctx.inv_std = inv_std
        return render_fwd(mu, (h, w), inv_std)

    @staticmethod
    def backward(ctx, d_maps):
        (mu,) = ctx.saved_tensors
        return render_bwd(d_maps, mu, ctx.inv_std), None, None, None
