"""numpy restatement of the reference keypoint math (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/utils/model.py op for op:
  get_coord           utils/model.py:63-70
  get_gaussian_maps   utils/model.py:49-60
  colorize_point_maps utils/model.py:42-46
and the call site models/networks/__init__.py:68-71 (x from axis 1 reduction,
y from axis 2 reduction, stacked as (x, y)).

Backward formulas are the analytic derivatives; they are
cross-checked against finite differences in tests/test_oracle_k1.py.

Parity status: pinned against the reference's own utils/model.py executed
through oracle/tf_shim (tests/golden/k1_*.npz), plus analytic known answers.
"""
import numpy as np


def _linspace(n, dtype):
    # tf.linspace(-1.0, 1.0, n) in fp32: start + i * ((stop-start)/(n-1))
    if n == 1:
        return np.array([-1.0], dtype=dtype)
    step = dtype(2.0) / dtype(n - 1)
    return (dtype(-1.0) + np.arange(n, dtype=dtype) * step).astype(dtype)


def softmax(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def get_coord(x, other_axis, axis_size):
    """utils/model.py:63-70.  x: [B,H,W,K].  Returns (coord [B,K], prob [B,N,K])."""
    dt = x.dtype.type
    g_c_prob = np.mean(x, axis=other_axis, dtype=x.dtype)
    g_c_prob = softmax(g_c_prob, axis=1)
    coord_pt = _linspace(axis_size, dt).reshape(1, axis_size, 1)
    g_c = np.sum(g_c_prob * coord_pt, axis=1, dtype=x.dtype)
    return g_c, g_c_prob


def soft_argmax(x):
    """models/networks/__init__.py:68-71: mu[b,k] = (x, y)."""
    H, W = x.shape[1], x.shape[2]
    gauss_y, prob_y = get_coord(x, 2, H)
    gauss_x, prob_x = get_coord(x, 1, W)
    mu = np.stack([gauss_x, gauss_y], axis=2)
    return mu, prob_x, prob_y


def get_gaussian_maps(mu, shape_hw, inv_std=14.3):
    """utils/model.py:49-60.  mu: [B,K,2] (x,y).  Returns [B,H,W,K]."""
    dt = mu.dtype.type
    mu_x, mu_y = mu[:, :, 0:1], mu[:, :, 1:2]
    y = _linspace(shape_hw[0], dt)
    x = _linspace(shape_hw[1], dt)
    mu_y, mu_x = mu_y[..., None], mu_x[..., None]
    y = y.reshape(1, 1, shape_hw[0], 1)
    x = x.reshape(1, 1, 1, shape_hw[1])
    g_y = np.square(y - mu_y)
    g_x = np.square(x - mu_x)
    dist = (g_y + g_x) * dt(inv_std ** 2)
    g_yx = np.transpose(np.exp(-dist), (0, 2, 3, 1))
    return np.ascontiguousarray(g_yx)


def colorize_point_maps(maps, colors):
    """utils/model.py:42-46: max over k of maps[...,k] * colour_k -> [B,H,W,3]."""
    colors = np.asarray(colors, dtype=maps.dtype)
    out = None
    for i in range(maps.shape[-1]):
        h = maps[..., i:i + 1] * colors[i].reshape(1, 1, 1, 3)
        out = h if out is None else np.maximum(out, h)
    return out


# ----------------------------------------------------------------------------
# analytic backward
# ----------------------------------------------------------------------------
def gaussian_maps_bwd(d_maps, mu, inv_std=14.3):
    """d_maps [B,h,w,K], mu [B,K,2] -> d_mu [B,K,2]."""
    dt = mu.dtype.type
    h, w = d_maps.shape[1], d_maps.shape[2]
    s = dt(inv_std ** 2)
    G = get_gaussian_maps(mu, [h, w], inv_std)
    cy = _linspace(h, dt).reshape(1, h, 1, 1)
    cx = _linspace(w, dt).reshape(1, 1, w, 1)
    mu_x = mu[:, :, 0].reshape(-1, 1, 1, mu.shape[1])
    mu_y = mu[:, :, 1].reshape(-1, 1, 1, mu.shape[1])
    dG = d_maps * G * (dt(2.0) * s)
    d_mu_x = np.sum(dG * (cx - mu_x), axis=(1, 2))
    d_mu_y = np.sum(dG * (cy - mu_y), axis=(1, 2))
    return np.stack([d_mu_x, d_mu_y], axis=2)


def soft_argmax_bwd(d_mu, mu, prob_x, prob_y, H, W):
    """d_mu [B,K,2] -> d_logits [B,H,W,K] (row-vector + column-vector structure)."""
    dt = mu.dtype.type
    cW = _linspace(W, dt).reshape(1, W, 1)
    cH = _linspace(H, dt).reshape(1, H, 1)
    dq = prob_x * (cW - mu[:, None, :, 0]) * d_mu[:, None, :, 0]   # [B,W,K]
    dr = prob_y * (cH - mu[:, None, :, 1]) * d_mu[:, None, :, 1]   # [B,H,K]
    return dr[:, :, None, :] / dt(W) + dq[:, None, :, :] / dt(H)


def softargmax_render_fwd(logits, map_hw, inv_std=14.3):
    mu, px, py = soft_argmax(logits)
    maps = get_gaussian_maps(mu, map_hw, inv_std)
    return mu, px, py, maps


def softargmax_render_bwd(d_maps, d_mu_extra, mu, px, py, H, W, inv_std=14.3):
    d_mu = np.zeros_like(mu)
    if d_maps is not None:
        d_mu = d_mu + gaussian_maps_bwd(d_maps, mu, inv_std)
    if d_mu_extra is not None:
        d_mu = d_mu + d_mu_extra
    return soft_argmax_bwd(d_mu, mu, px, py, H, W)
