"""Keypoint math utilities — drop-in for the reference's ``utils/model.py`` on torch CUDA tensors.

Same function names, argument order and returned structures as
/root/reference/utils/model.py:13-70; the arithmetic runs in the fused sm_100a kernels of
``csrc/k1_keypoints.cu`` through the C ABI (include/kp_b200.h).  No CPU path.
"""
import random

import torch

from .. import k1 as _k1


# ---- colour helpers (host side, reference utils/model.py:13-39; unseeded like the reference) --------
def get_random_color(pastel_factor=0.5):
    return [(x + pastel_factor) / (1.0 + pastel_factor) for x in [random.uniform(0, 1.0) for _ in (1, 2, 3)]]


def color_distance(c1, c2):
    return sum(abs(a - b) for a, b in zip(c1, c2))


def generate_new_color(existing_colors, pastel_factor=0.5):
    max_distance = None
    best_color = None
    for _ in range(100):
        color = get_random_color(pastel_factor=pastel_factor)
        if not existing_colors:
            return color
        best_distance = min(color_distance(color, c) for c in existing_colors)
        if not max_distance or best_distance > max_distance:
            max_distance = best_distance
            best_color = color
    return best_color


def get_n_colors(n, pastel_factor=0.9):
    # the reference ignores its pastel_factor argument and always uses 0.9 (utils/model.py:35-39)
    colors = []
    for _ in range(n):
        colors.append(generate_new_color(colors, pastel_factor=0.9))
    return colors


# ---- device ops ---------------------------------------------------------------------------------
class _GetCoord(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, axis_idx):
        mu, px, py, _ = _k1.softargmax_render_fwd(x, None, want_prob=True)
        ctx.save_for_backward(mu, px, py)
        ctx.hw = (x.shape[1], x.shape[2])
        ctx.axis_idx = axis_idx
        coord = mu[:, :, axis_idx].contiguous()
        prob = px if axis_idx == 0 else py
        ctx.mark_non_differentiable(prob)
        return coord, prob

    @staticmethod
    def backward(ctx, d_coord, _d_prob):
        mu, px, py = ctx.saved_tensors
        d_mu = torch.zeros_like(mu)
        d_mu[:, :, ctx.axis_idx] = d_coord
        H, W = ctx.hw
        return _k1.softargmax_render_bwd(None, d_mu, mu, px, py, H, W), None


def get_coord(x, other_axis, axis_size):
    """Marginal soft-argmax along one image axis (reference utils/model.py:63-70).

    x: [B,H,W,K] float32 CUDA.  ``other_axis`` is the axis averaged away (2 -> returns the y
    coordinate over H, 1 -> the x coordinate over W), exactly like the reference call sites
    models/networks/__init__.py:69-70.  Returns (coord [B,K], prob [B,axis_size,K]).
    """
    if x.dim() != 4:
        raise ValueError("get_coord expects [B,H,W,K]")
    if other_axis not in (1, 2):
        raise ValueError("other_axis must be 1 or 2")
    kept = 1 if other_axis == 2 else 2
    if x.shape[kept] != axis_size:
        raise ValueError("axis_size %d does not match the kept axis (%d)" % (axis_size, x.shape[kept]))
    return _GetCoord.apply(x, 1 if other_axis == 2 else 0)


def get_gaussian_maps(mu, shape_hw, inv_std=14.3):
    """mu [B,K,2] (x,y) -> [B,H,W,K] un-normalised Gaussians (reference utils/model.py:49-60)."""
    return _k1.GaussianMaps.apply(mu, int(shape_hw[0]), int(shape_hw[1]), float(inv_std))


def colorize_point_maps(maps, colors):
    """max_k maps[...,k] * colour_k -> [...,3] (reference utils/model.py:42-46)."""
    return _k1.colorize(maps, colors)


def soft_argmax(x):
    """Both get_coord calls + stack((x,y)) of models/networks/__init__.py:68-71 in one kernel."""
    return _k1.SoftArgmax.apply(x)


def soft_argmax_and_maps(x, shape_hw, inv_std=14.3):
    """Fused detector tail: logits -> (mu [B,K,2], maps [B,h,w,K]) in ONE pass over the logits."""
    return _k1.SoftArgmaxRender.apply(x, int(shape_hw[0]), int(shape_hw[1]), float(inv_std))


def gaussian_maps_colorized(mu, colors, shape_hw, inv_std=14.3):
    """get_gaussian_maps + colorize_point_maps without materialising the [B,H,W,K] maps."""
    return _k1.render_colorize(mu, colors, shape_hw, inv_std)
