"""torch (autograd-able) twin of oracle/k1_numpy.py, used inside the network oracle (TEST INFRASTRUCTURE ONLY).

Same op order as /root/reference/utils/model.py:49-70 and models/networks/__init__.py:68-71.
"""
import torch


def _linspace(n, dtype):
    if n == 1:
        return torch.tensor([-1.0], dtype=dtype)
    step = torch.tensor(2.0, dtype=dtype) / torch.tensor(float(n - 1), dtype=dtype)
    return torch.tensor(-1.0, dtype=dtype) + torch.arange(n, dtype=dtype) * step


def get_coord(x, other_axis, axis_size):
    g_c_prob = torch.mean(x, dim=other_axis)
    g_c_prob = torch.softmax(g_c_prob, dim=1)
    coord_pt = _linspace(axis_size, x.dtype).reshape(1, axis_size, 1)
    g_c = torch.sum(g_c_prob * coord_pt, dim=1)
    return g_c, g_c_prob


def soft_argmax(x):
    gauss_y, _ = get_coord(x, 2, x.shape[1])
    gauss_x, _ = get_coord(x, 1, x.shape[2])
    return torch.stack([gauss_x, gauss_y], dim=2)


def get_gaussian_maps(mu, shape_hw, inv_std=14.3):
    mu_x, mu_y = mu[:, :, 0:1], mu[:, :, 1:2]
    y = _linspace(shape_hw[0], mu.dtype)
    x = _linspace(shape_hw[1], mu.dtype)
    mu_y, mu_x = mu_y.unsqueeze(-1), mu_x.unsqueeze(-1)
    y = y.reshape(1, 1, shape_hw[0], 1)
    x = x.reshape(1, 1, 1, shape_hw[1])
    g_y = torch.square(y - mu_y)
    g_x = torch.square(x - mu_x)
    dist = (g_y + g_x) * inv_std ** 2
    return torch.exp(-dist).permute(0, 2, 3, 1)
