"""GPU: K1 CUDA kernels (through the C ABI) against the numpy oracle.

Tolerances are BASELINE.json's: keypoints <= 1e-4 (normalised units), rendered maps <= 1e-5 absolute
(judged on identical mu), gradients to 1e-4 relative of the gradient scale.
"""
import numpy as np
import pytest
import torch

from oracle import k1_numpy as o

pytestmark = pytest.mark.gpu

MU_TOL = 1e-4
MAP_TOL = 1e-5


def _dev(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.mark.parametrize("sigma", [1.0, 5.0, 30.0])
@pytest.mark.parametrize("B", [1, 3, 8])
def test_fused_forward_penn_shape(cuda_dev, sigma, B):
    from kp_b200 import k1
    rng = np.random.default_rng(int(sigma * 10) + B)
    x = rng.normal(0, sigma, (B, 128, 128, 40)).astype(np.float32)
    mu64, px64, py64 = o.soft_argmax(x.astype(np.float64))
    mu, px, py, maps = k1.softargmax_render_fwd(_dev(x, cuda_dev), (32, 32))
    mu, px, py, maps = mu.cpu().numpy(), px.cpu().numpy(), py.cpu().numpy(), maps.cpu().numpy()
    assert np.abs(mu - mu64).max() <= MU_TOL
    assert np.abs(px - px64).max() <= 1e-5 and np.abs(py - py64).max() <= 1e-5
    # maps are judged on identical mu (the kernel's own mu)
    ref_maps = o.get_gaussian_maps(mu.astype(np.float64), [32, 32])
    assert np.abs(maps - ref_maps).max() <= MAP_TOL


def test_known_answers_on_gpu(cuda_dev):
    from kp_b200 import k1
    x = np.full((2, 128, 128, 40), 1.5, dtype=np.float32)
    mu, px, py, maps = k1.softargmax_render_fwd(_dev(x, cuda_dev), (32, 32))
    assert mu.abs().max().item() < 1e-6
    assert abs(maps.max().item() - 0.653392) < 1e-5
    x = np.zeros((1, 128, 128, 40), dtype=np.float32)
    x[0, 17, 93, :] = 128 * 60.0
    mu = k1.softargmax_render_fwd(_dev(x, cuda_dev), None)[0].cpu().numpy()
    lin = np.linspace(-1, 1, 128)
    assert np.abs(mu[0, :, 0] - lin[93]).max() < 1e-6
    assert np.abs(mu[0, :, 1] - lin[17]).max() < 1e-6


@pytest.mark.parametrize("shape", [(2, 5, 7, 3), (1, 128, 128, 8), (3, 64, 128, 40), (2, 16, 16, 40),
                                   (1, 1, 1, 1), (2, 200, 128, 40), (1, 31, 128, 40)])
def test_forward_ragged_shapes(cuda_dev, shape):
    """Off-fast-path shapes (generic kernel) and fast-path shapes with H != 128."""
    from kp_b200 import k1
    rng = np.random.default_rng(sum(shape))
    x = rng.normal(0, 3.0, shape).astype(np.float32)
    mu64, px64, py64 = o.soft_argmax(x.astype(np.float64))
    hm, wm = 9, 13
    mu, px, py, maps = k1.softargmax_render_fwd(_dev(x, cuda_dev), (hm, wm))
    mu = mu.cpu().numpy()
    assert np.abs(mu - mu64).max() <= MU_TOL
    assert np.abs(px.cpu().numpy() - px64).max() <= 1e-5
    assert np.abs(py.cpu().numpy() - py64).max() <= 1e-5
    ref_maps = o.get_gaussian_maps(mu.astype(np.float64), [hm, wm])
    assert np.abs(maps.cpu().numpy() - ref_maps).max() <= MAP_TOL


def test_empty_batch(cuda_dev):
    from kp_b200 import k1
    x = torch.empty((0, 128, 128, 40), device=cuda_dev)
    mu, px, py, maps = k1.softargmax_render_fwd(x, (32, 32))
    assert mu.shape == (0, 40, 2) and maps.shape == (0, 32, 32, 40)


@pytest.mark.parametrize("hw", [(32, 32), (128, 128), (64, 64), (5, 9)])
def test_render_standalone(cuda_dev, hw):
    from kp_b200 import model_utils
    rng = np.random.default_rng(7)
    mu = rng.uniform(-0.95, 0.95, (6, 40, 2)).astype(np.float32)
    maps = model_utils.get_gaussian_maps(_dev(mu, cuda_dev), list(hw)).cpu().numpy()
    ref = o.get_gaussian_maps(mu.astype(np.float64), list(hw))
    assert maps.shape == ref.shape
    assert np.abs(maps - ref).max() <= MAP_TOL
    # odd K (scalar path)
    mu3 = rng.uniform(-0.95, 0.95, (2, 3, 2)).astype(np.float32)
    maps3 = model_utils.get_gaussian_maps(_dev(mu3, cuda_dev), list(hw)).cpu().numpy()
    assert np.abs(maps3 - o.get_gaussian_maps(mu3.astype(np.float64), list(hw))).max() <= MAP_TOL


def test_get_coord_api(cuda_dev):
    from kp_b200 import model_utils
    rng = np.random.default_rng(11)
    x = rng.normal(0, 4.0, (2, 128, 128, 40)).astype(np.float32)
    xd = _dev(x, cuda_dev)
    gy, py = model_utils.get_coord(xd, 2, 128)
    gx, px = model_utils.get_coord(xd, 1, 128)
    ry, rpy = o.get_coord(x.astype(np.float64), 2, 128)
    rx, rpx = o.get_coord(x.astype(np.float64), 1, 128)
    assert np.abs(gy.cpu().numpy() - ry).max() <= MU_TOL
    assert np.abs(gx.cpu().numpy() - rx).max() <= MU_TOL
    assert np.abs(px.cpu().numpy() - rpx).max() <= 1e-5
    assert np.abs(py.cpu().numpy() - rpy).max() <= 1e-5


def test_colorize(cuda_dev):
    from kp_b200 import model_utils
    rng = np.random.default_rng(13)
    mu = rng.uniform(-0.9, 0.9, (3, 40, 2)).astype(np.float32)
    colors = rng.uniform(0, 1, (40, 3)).astype(np.float32)
    maps = o.get_gaussian_maps(mu.astype(np.float64), [128, 128])
    ref = o.colorize_point_maps(maps, colors.astype(np.float64))
    fused = model_utils.gaussian_maps_colorized(_dev(mu, cuda_dev), colors.tolist(), [128, 128]).cpu().numpy()
    assert np.abs(fused - ref).max() <= MAP_TOL
    two_step = model_utils.colorize_point_maps(_dev(maps.astype(np.float32), cuda_dev), colors.tolist()).cpu().numpy()
    assert np.abs(two_step - ref).max() <= MAP_TOL


@pytest.mark.parametrize("shape,hw", [((2, 128, 128, 40), (32, 32)), ((2, 6, 5, 3), (4, 7)), ((1, 48, 128, 40), (16, 8))])
def test_backward_against_oracle(cuda_dev, shape, hw):
    from kp_b200 import k1
    rng = np.random.default_rng(17)
    B, H, W, K = shape
    x = rng.normal(0, 3.0, shape).astype(np.float32)
    cot_maps = rng.normal(size=(B, hw[0], hw[1], K)).astype(np.float32)
    cot_mu = rng.normal(size=(B, K, 2)).astype(np.float32)
    x64 = x.astype(np.float64)
    mu, px, py, _ = o.softargmax_render_fwd(x64, list(hw))
    ref = o.softargmax_render_bwd(cot_maps.astype(np.float64), cot_mu.astype(np.float64), mu, px, py, H, W)
    xd = _dev(x, cuda_dev).requires_grad_(True)
    mu_d, maps_d = k1.SoftArgmaxRender.apply(xd, hw[0], hw[1], 14.3)
    (maps_d * _dev(cot_maps, cuda_dev)).sum().add((mu_d * _dev(cot_mu, cuda_dev)).sum()).backward()
    got = xd.grad.cpu().numpy()
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-4 * scale + 1e-9
    # maps-only and mu-only gradient entry points
    only_maps = k1.softargmax_render_bwd(_dev(cot_maps, cuda_dev), None, mu_d.detach(),
                                         *[t for t in k1.softargmax_render_fwd(xd.detach(), None)[1:3]], H, W)
    ref_m = o.softargmax_render_bwd(cot_maps.astype(np.float64), None, mu, px, py, H, W)
    assert np.abs(only_maps.cpu().numpy() - ref_m).max() <= 2e-4 * np.abs(ref_m).max() + 1e-9
    mu2 = k1.SoftArgmax.apply(xd)
    g2 = torch.autograd.grad((mu2 * _dev(cot_mu, cuda_dev)).sum(), xd)[0].cpu().numpy()
    ref_u = o.softargmax_render_bwd(None, cot_mu.astype(np.float64), mu, px, py, H, W)
    assert np.abs(g2 - ref_u).max() <= 2e-4 * np.abs(ref_u).max() + 1e-9


def test_render_backward(cuda_dev):
    from kp_b200 import model_utils
    rng = np.random.default_rng(19)
    mu = rng.uniform(-0.9, 0.9, (4, 40, 2)).astype(np.float32)
    cot = rng.normal(size=(4, 32, 32, 40)).astype(np.float32)
    ref = o.gaussian_maps_bwd(cot.astype(np.float64), mu.astype(np.float64))
    md = _dev(mu, cuda_dev).requires_grad_(True)
    (model_utils.get_gaussian_maps(md, [32, 32]) * _dev(cot, cuda_dev)).sum().backward()
    assert np.abs(md.grad.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max()


def test_full_size_properties(cuda_dev):
    """BASELINE config 2 size (1024 frames): size-independent properties instead of a slow oracle run."""
    from kp_b200 import k1
    g = torch.Generator(device=cuda_dev).manual_seed(0)
    B = 1024
    x = torch.randn((B, 128, 128, 40), device=cuda_dev, generator=g) * 5.0
    mu, px, py, maps = k1.softargmax_render_fwd(x, (32, 32))
    assert torch.isfinite(mu).all() and mu.abs().max() <= 1.0
    # probabilities sum to one; mu is the expectation of the grid under them
    assert (px.sum(1) - 1).abs().max() < 1e-5 and (py.sum(1) - 1).abs().max() < 1e-5
    lin = torch.linspace(-1, 1, 128, device=cuda_dev).view(1, 128, 1)
    assert ((px * lin).sum(1) - mu[:, :, 0]).abs().max() < 1e-5
    assert ((py * lin).sum(1) - mu[:, :, 1]).abs().max() < 1e-5
    # shift invariance: adding a per-keypoint constant leaves mu unchanged
    mu2 = k1.softargmax_render_fwd(x + 3.0, None, want_prob=False)[0]
    assert (mu2 - mu).abs().max() < 5e-5
    # flipping the image flips the coordinates
    mu3 = k1.softargmax_render_fwd(torch.flip(x, dims=[1, 2]).contiguous(), None, want_prob=False)[0]
    assert (mu3 + mu).abs().max() < 5e-5
    # frames are independent: a sub-batch gives identical results (bit-exact)
    mu4, _, _, maps4 = k1.softargmax_render_fwd(x[100:108].contiguous(), (32, 32))
    assert torch.equal(mu4, mu[100:108]) and torch.equal(maps4, maps[100:108])
    # maps equal the standalone renderer on the same mu, bit for bit
    assert torch.equal(k1.render_fwd(mu, (32, 32)), maps)
    # full-size check against the oracle on a strided sample of frames
    idx = list(range(0, B, 97))
    mu64, _, _ = o.soft_argmax(x[idx].double().cpu().numpy())
    assert np.abs(mu[idx].cpu().numpy() - mu64).max() <= MU_TOL
