"""CPU: hand-computed known answers for the TF-1.12 leaf semantics restated in oracle/tf_ops.py.

oracle/tf_shim.py (which executes the reference's own sources for the golden fixtures) and oracle/networks.py both
call oracle/tf_ops.py, so the fixture tests cannot catch a wrong leaf rule.  These cases are derived by hand from the
TF-1.12 definitions, independent of the code under test.
"""
import math

import numpy as np
import torch

from oracle import k1_numpy, tf_ops as T


def test_same_padding_rule():
    # out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total//2; after = total - before
    assert T.same_pad(128, 3, 1) == (1, 1)
    assert T.same_pad(128, 3, 2) == (0, 1)          # the asymmetric case of encoder conv_3/5/7
    assert T.same_pad(128, 7, 1) == (3, 3)
    assert T.same_pad(130, 4, 2) == (1, 1)          # img_discr conv_0: explicit pad 1 (128 -> 130), then SAME
    assert T.same_pad(67, 4, 2) == (1, 2)           # conv_1: 65 + 2 = 67 (odd) -> total 3
    assert T.same_pad(5, 3, 2) == (1, 1)
    assert T.same_pad(6, 3, 1) == (1, 1)            # D_logit after explicit pad 1: 4 -> 6


def test_img_discr_spatial_chain():
    """128 -> 65 -> 34 -> 18 -> 10 -> 6 -> 4, then D_logit 4 -> 6."""
    size, sizes = 128, []
    for _ in range(6):
        x = torch.zeros(1, size, size, 1, dtype=torch.float64)
        size = T.conv2d(x, torch.zeros(4, 4, 1, 1, dtype=torch.float64), None, 2, 1).shape[1]
        sizes.append(size)
    assert sizes == [65, 34, 18, 10, 6, 4]
    x = torch.zeros(1, 4, 4, 1, dtype=torch.float64)
    assert T.conv2d(x, torch.zeros(3, 3, 1, 1, dtype=torch.float64), None, 1, 1).shape[1] == 6


def test_conv2d_asymmetric_same_padding_known_answer():
    """4x4 ones * 3x3 ones, stride 2: SAME pads (0,1), so windows start at rows/cols 0 and 2 and the second one hangs
    over the bottom/right edge: [[9, 6], [6, 4]].  (Symmetric-before padding (1,0) would give [[4, 6], [6, 9]].)"""
    x = torch.ones(1, 4, 4, 1, dtype=torch.float64)
    w = torch.ones(3, 3, 1, 1, dtype=torch.float64)
    y = T.conv2d(x, w, None, 2, 0)[0, :, :, 0]
    assert torch.equal(y, torch.tensor([[9.0, 6.0], [6.0, 4.0]], dtype=torch.float64))
    # a position-dependent input pins the window origin as well: x[h,w] = 10*h + w, 1x1 centre tap of a 3x3 kernel
    x = (10 * torch.arange(4).view(4, 1) + torch.arange(4).view(1, 4)).double().view(1, 4, 4, 1)
    w = torch.zeros(3, 3, 1, 1, dtype=torch.float64)
    w[1, 1] = 1.0
    y = T.conv2d(x, w, None, 2, 0)[0, :, :, 0]
    assert torch.equal(y, torch.tensor([[11.0, 13.0], [31.0, 33.0]], dtype=torch.float64))   # centres at (1,1),(1,3),(3,1),(3,3)
    # stride 1: centre tap is the identity
    assert torch.equal(T.conv2d(x, w, None, 1, 0), x)
    # bias, HWIO layout: out[..., o] = sum_i x[..., i] * w[0,0,i,o] + b[o]
    x2 = torch.tensor([1.0, 2.0], dtype=torch.float64).view(1, 1, 1, 2)
    w2 = torch.tensor([[1.0, 10.0, 100.0], [2.0, 20.0, 200.0]], dtype=torch.float64).view(1, 1, 2, 3)
    b2 = torch.tensor([0.5, 0.25, 0.125], dtype=torch.float64)
    assert torch.equal(T.conv2d(x2, w2, b2, 1, 0).flatten(), torch.tensor([5.5, 50.25, 500.125], dtype=torch.float64))


def test_explicit_pad_then_same():
    """layers.conv(pad=1) = tf.pad 1 on every side THEN a SAME convolution (img_discr): 2x2 ones, 4x4 ones kernel,
    stride 2 -> padded 4x4, out = ceil(4/2) = 2, SAME total = 2 -> (1,1): windows rows [-1,3) and [1,5) of the padded
    image; each window covers the 2x2 block of ones exactly once in every direction."""
    x = torch.ones(1, 2, 2, 1, dtype=torch.float64)
    w = torch.ones(4, 4, 1, 1, dtype=torch.float64)
    y = T.conv2d(x, w, None, 2, 1)[0, :, :, 0]
    assert torch.equal(y, torch.full((2, 2), 4.0, dtype=torch.float64))


def test_legacy_bilinear_resize_known_answer():
    """tf.image.resize_images in TF 1.12: bilinear, align_corners=False, no half-pixel centres: src = dst * in/out.
    [0,1,2,3] -> [0,.5,1,1.5,2,2.5,3,3] (the last sample clamps: src 3.5 -> lerp(x[3], x[3]))."""
    x = torch.arange(4, dtype=torch.float64).view(1, 1, 4, 1)
    y = T.resize_bilinear_legacy(x, 1, 8).flatten()
    assert torch.equal(y, torch.tensor([0, .5, 1, 1.5, 2, 2.5, 3, 3], dtype=torch.float64))
    x = torch.arange(4, dtype=torch.float64).view(1, 4, 1, 1)
    y = T.resize_bilinear_legacy(x, 8, 1).flatten()
    assert torch.equal(y, torch.tensor([0, .5, 1, 1.5, 2, 2.5, 3, 3], dtype=torch.float64))
    # 2-D: out[2i+1, 2j+1] is the mean of the 2x2 neighbourhood, even positions copy
    x = torch.tensor([[0.0, 4.0], [8.0, 16.0]], dtype=torch.float64).view(1, 2, 2, 1)
    y = T.resize_bilinear_legacy(x, 4, 4)[0, :, :, 0]
    ref = torch.tensor([[0, 2, 4, 4], [4, 7, 10, 10], [8, 12, 16, 16], [8, 12, 16, 16]], dtype=torch.float64)
    assert torch.equal(y, ref)


def test_fused_batch_norm_known_answer():
    """contrib batch_norm, train: normalise with the BIASED variance, moving_var takes the UNBIASED one, decay 0.999."""
    x = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64).view(4, 1, 1, 1)
    g, b = torch.tensor([2.0], dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64)
    mm, mv = torch.tensor([10.0], dtype=torch.float64), torch.tensor([1.0], dtype=torch.float64)
    y, nmm, nmv = T.batch_norm(x, g, b, mm, mv, True)
    ref = (np.array([1, 2, 3, 4.0]) - 2.5) / math.sqrt(1.25 + 1e-5) * 2.0 + 0.5
    assert np.allclose(y.flatten().numpy(), ref, rtol=0, atol=1e-14)
    assert abs(float(nmm) - (10.0 * 0.999 + 2.5 * 0.001)) < 1e-14
    assert abs(float(nmv) - (1.0 * 0.999 + (5.0 / 3.0) * 0.001)) < 1e-14       # unbiased: 1.25 * 4/3
    yi, _, _ = T.batch_norm(x, g, b, mm, mv, False)
    assert np.allclose(yi.flatten().numpy(), (np.array([1, 2, 3, 4.0]) - 10.0) / math.sqrt(1.0 + 1e-5) * 2.0 + 0.5, atol=1e-14)


def test_max_pool_leaky_bce_known_answers():
    x = torch.tensor([[1.0, 5.0, 2.0, 0.0], [3.0, 4.0, 9.0, 1.0], [0.0, 0.0, -1.0, -2.0], [7.0, 0.0, -3.0, -4.0]], dtype=torch.float64)
    y = T.max_pool_2x2(x.view(1, 4, 4, 1))[0, :, :, 0]
    assert torch.equal(y, torch.tensor([[5.0, 9.0], [7.0, -1.0]], dtype=torch.float64))
    assert torch.equal(T.leaky_relu(torch.tensor([-2.0, 0.0, 3.0]), 0.01), torch.tensor([-0.02, 0.0, 3.0]))
    # max(x,0) - x z + log1p(exp(-|x|))
    v = T.sigmoid_cross_entropy_with_logits(torch.tensor([0.0, 2.0, -2.0, 2.0], dtype=torch.float64),
                                            torch.tensor([1.0, 0.0, 0.0, 1.0], dtype=torch.float64))
    ref = [math.log(2.0), 2.0 + math.log1p(math.exp(-2.0)), math.log1p(math.exp(-2.0)), math.log1p(math.exp(-2.0))]
    assert np.allclose(v.numpy(), ref, atol=1e-15)


def test_adam_and_learning_rate_known_answers():
    """TF Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).  First step with g = 1, b1 = .5, b2 = .999:
    m = .5, v = .001, lr_t = lr*sqrt(.001)/.5, so p moves by lr*(1 - eps/sqrt(.001)) ~ lr."""
    p, m, v = T.adam_tf(torch.tensor([1.0], dtype=torch.float64), torch.tensor([1.0], dtype=torch.float64),
                        torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64), 1, 1e-4)
    assert abs(float(m) - 0.5) < 1e-15 and abs(float(v) - 0.001) < 1e-15
    step = 1e-4 * math.sqrt(0.001) / 0.5 * 0.5 / (math.sqrt(0.001) + 1e-8)
    assert abs(float(p) - (1.0 - step)) < 1e-15 and abs(step - 1e-4) < 1e-10
    # second step, g = -1: m = .5*.5 - .5 = -.25, v = .001*.999 + .001
    p2, m2, v2 = T.adam_tf(p, torch.tensor([-1.0], dtype=torch.float64), m, v, 2, 1e-4)
    assert abs(float(m2) + 0.25) < 1e-15 and abs(float(v2) - 0.001999) < 1e-15
    lr_t = 1e-4 * math.sqrt(1 - 0.999 ** 2) / (1 - 0.25)
    assert abs(float(p2) - (float(p) + lr_t * 0.25 / (math.sqrt(0.001999) + 1e-8))) < 1e-15
    # exponential_decay, staircase=False: 1e-4 * 0.95^(10000/20000)
    assert abs(T.exponential_decay(1e-4, 10000, 20000, 0.95) - 1e-4 * math.sqrt(0.95)) < 1e-18


def test_xavier_uniform_limit():
    rng = np.random.default_rng(0)
    w = T.xavier_uniform(rng, (3, 3, 16, 32), torch.float64)
    limit = math.sqrt(6.0 / (9 * 16 + 9 * 32))
    assert float(w.abs().max()) <= limit and float(w.abs().max()) > 0.98 * limit
    assert abs(float(w.var()) - limit ** 2 / 3) < 0.1 * limit ** 2 / 3             # uniform variance


def test_keypoint_known_answers():
    """utils/model.py:49-70: a constant map gives (0,0); a spike gives its grid coordinate -1 + 2 i/(N-1); the Gaussian
    map at distance d is exp(-(14.3 d)^2)."""
    H = W = 16
    logits = np.zeros((1, H, W, 1))
    mu = k1_numpy.soft_argmax(logits)[0]
    assert np.abs(mu).max() < 1e-15
    logits[0, 3, 11, 0] = 1e4        # (row 3, column 11); mean over the other axis keeps 1e4/16 = 625 >> 0
    mu = k1_numpy.soft_argmax(logits)[0]
    assert abs(mu[0, 0, 0] - (-1 + 2 * 11 / 15)) < 1e-12 and abs(mu[0, 0, 1] - (-1 + 2 * 3 / 15)) < 1e-12   # (x, y)
    maps = k1_numpy.get_gaussian_maps(np.array([[[0.0, 0.0]]]), [3, 3])
    assert abs(maps[0, 1, 1, 0] - 1.0) < 1e-15
    assert abs(maps[0, 0, 1, 0] - math.exp(-(14.3 ** 2))) < 1e-300 + 1e-15 * math.exp(-(14.3 ** 2))
    maps = k1_numpy.get_gaussian_maps(np.array([[[0.05, -0.02]]]), [2, 2])     # grid {-1, 1}: all far away
    assert maps.max() < 1e-70
    maps = k1_numpy.get_gaussian_maps(np.array([[[0.9, -0.95]]]), [2, 2])
    assert abs(maps[0, 0, 1, 0] - math.exp(-(14.3 ** 2) * (0.05 ** 2 + 0.1 ** 2))) < 1e-15   # y = -1 row, x = +1 column
