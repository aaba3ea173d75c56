"""CPU: the numpy K1 oracle against analytic known answers and finite differences."""
import numpy as np
import pytest

from oracle import k1_numpy as o


def test_constant_heatmap_gives_centre():
    x = np.full((2, 128, 128, 40), 3.25, dtype=np.float32)
    mu, px, py = o.soft_argmax(x)
    assert np.abs(mu).max() < 1e-6
    np.testing.assert_allclose(px, 1.0 / 128, rtol=1e-6)
    np.testing.assert_allclose(py, 1.0 / 128, rtol=1e-6)


def test_spike_gives_grid_coordinate():
    x = np.zeros((1, 128, 128, 40), dtype=np.float64)
    h0, w0 = 17, 93
    x[0, h0, w0, :] = 128 * 60.0
    mu, _, _ = o.soft_argmax(x)
    lin = np.linspace(-1, 1, 128)
    np.testing.assert_allclose(mu[0, :, 0], lin[w0], atol=1e-9)   # x comes first
    np.testing.assert_allclose(mu[0, :, 1], lin[h0], atol=1e-9)


def test_gaussian_peak_value_and_separability():
    mu = np.zeros((1, 40, 2), dtype=np.float64)
    G = o.get_gaussian_maps(mu, [32, 32])
    assert G.shape == (1, 32, 32, 40)
    np.testing.assert_allclose(G.max(), np.exp(-2 * (1 / 31) ** 2 * 14.3 ** 2), rtol=1e-12)
    assert abs(G.max() - 0.653392) < 1e-6
    rng = np.random.default_rng(0)
    mu = rng.uniform(-0.9, 0.9, (3, 40, 2))
    G = o.get_gaussian_maps(mu, [32, 24])
    lin_y, lin_x = np.linspace(-1, 1, 32), np.linspace(-1, 1, 24)
    gy = np.exp(-(lin_y[None, :, None] - mu[:, None, :, 1]) ** 2 * 14.3 ** 2)
    gx = np.exp(-(lin_x[None, :, None] - mu[:, None, :, 0]) ** 2 * 14.3 ** 2)
    np.testing.assert_allclose(G, gy[:, :, None, :] * gx[:, None, :, :], rtol=1e-10, atol=1e-300)
    assert (G > 0).all() or (G >= 0).all()
    assert G.max() <= 1.0


def test_head_commutes_with_marginal_mean():
    rng = np.random.default_rng(1)
    feat = rng.normal(size=(1, 16, 16, 8))
    Wt = rng.normal(size=(8, 5))
    logits = feat @ Wt
    a = logits.mean(axis=2)
    b = feat.mean(axis=2) @ Wt
    np.testing.assert_allclose(a, b, atol=1e-12)


def test_colorize_matches_definition():
    rng = np.random.default_rng(2)
    maps = rng.uniform(0, 1, (2, 5, 6, 7))
    colors = rng.uniform(0, 1, (7, 3))
    out = o.colorize_point_maps(maps, colors)
    ref = (maps[..., :, None] * colors[None, None, None]).max(axis=3)
    np.testing.assert_allclose(out, ref)


def _fd(f, x, eps=1e-6):
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        g[i] = (f(xp) - f(xm)) / (2 * eps)
    return g


def test_backward_formulas_match_finite_differences():
    rng = np.random.default_rng(3)
    B, H, W, K, hm, wm = 1, 6, 5, 3, 4, 7
    x = rng.normal(0, 2.0, (B, H, W, K))
    cot_maps = rng.normal(size=(B, hm, wm, K))
    cot_mu = rng.normal(size=(B, K, 2))

    def loss(xx):
        mu, px, py, maps = o.softargmax_render_fwd(xx, [hm, wm])
        return float((maps * cot_maps).sum() + (mu * cot_mu).sum())

    mu, px, py, _ = o.softargmax_render_fwd(x, [hm, wm])
    ana = o.softargmax_render_bwd(cot_maps, cot_mu, mu, px, py, H, W)
    num = _fd(loss, x)
    np.testing.assert_allclose(ana, num, rtol=2e-5, atol=1e-7)


def test_fp32_oracle_close_to_fp64():
    rng = np.random.default_rng(4)
    x64 = rng.normal(0, 5.0, (2, 128, 128, 40))
    mu64, _, _ = o.soft_argmax(x64)
    mu32, _, _ = o.soft_argmax(x64.astype(np.float32))
    assert np.abs(mu64 - mu32).max() < 5e-6
    m64 = o.get_gaussian_maps(mu64, [32, 32])
    m32 = o.get_gaussian_maps(mu64.astype(np.float32), [32, 32])
    assert np.abs(m64 - m32).max() < 5e-6
