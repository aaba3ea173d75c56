"""CPU: the host-side lowering of TF convolutions (fwd, dgrad, stride 2, virtual concat) onto the tap-GEMM
primitive, executed by the numpy emulator of the device primitive and compared with the conv oracle."""
import numpy as np
import pytest
import torch

import kp_b200  # noqa: F401  (import alias)
from kp_b200 import tapconv as tc
from oracle import tapconv_emu as emu
from oracle import tf_ops as T

CASES = [
    # (N, H, W, [C sources], k, stride, pad, cout)
    (2, 8, 8, [16], 3, 1, 0, 24),
    (1, 8, 10, [32], 3, 2, 0, 16),      # SAME stride 2 on even input: pad (0,1)
    (1, 9, 7, [16], 3, 2, 0, 16),       # odd input: pad (1,1)
    (2, 6, 6, [16, 32], 3, 1, 0, 8),    # virtual concat, mixed channel counts -> CB 16
    (1, 16, 16, [64, 64], 3, 1, 0, 32),
    (1, 12, 12, [16], 7, 1, 0, 32),     # encoder conv_1 shape class (49 taps)
    (1, 8, 8, [16], 1, 1, 0, 40),       # 1x1 head, Cout not a multiple of 16
    (1, 13, 13, [16], 4, 2, 1, 32),     # img_discr: explicit pad 1 + SAME, odd size (65 -> 34 class)
    (1, 10, 10, [64], 4, 2, 1, 16),     # img_discr even size (34 -> 18 class): pad (2,2)
    (1, 4, 4, [32], 3, 1, 1, 16),       # D_logit: pad 1 + SAME -> 6x6 output
    (1, 8, 8, [40], 3, 1, 0, 16),       # channels not a multiple of the block (zero-filled tail)
]


@pytest.mark.parametrize("case", CASES)
def test_forward_lowering(case):
    N, H, W, Cs, k, s, pad, cout = case
    rng = np.random.default_rng(hash(case[:3]) % 1000)
    xs = [rng.normal(size=(N, H, W, C)) for C in Cs]
    w = rng.normal(size=(k, k, sum(Cs), cout))
    b = rng.normal(size=(cout,))
    ref = T.conv2d(torch.from_numpy(np.concatenate(xs, axis=-1)), torch.from_numpy(w), torch.from_numpy(b), s, pad).numpy()
    plan, (n, ho, wo) = tc.plan_conv_fwd([x.shape for x in xs], k, s, pad, cout)
    assert (n, ho, wo) == ref.shape[:3]
    wp = tc.pack_weights_np(plan, w)
    assert wp.shape == (plan.rows_pad, plan.Ktot)
    out = np.full((n, ho, wo, cout), np.nan)
    emu.run_plan(plan, xs, wp, np.pad(b, (0, plan.rows_pad - cout)), out)
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)
    d = plan.desc()
    assert d.n_taps == k * k and d.Ktot == plan.Ktot and d.Cout_pad % 16 == 0


def test_forward_into_channel_slice():
    rng = np.random.default_rng(5)
    x = rng.normal(size=(1, 6, 6, 16))
    w = rng.normal(size=(3, 3, 16, 8))
    ref = T.conv2d(torch.from_numpy(x), torch.from_numpy(w), None, 1, 0).numpy()
    plan, _ = tc.plan_conv_fwd([x.shape], 3, 1, 0, 8, out_channels_total=24, out_channel_off=8)
    out = np.zeros((1, 6, 6, 24))
    emu.run_plan(plan, [x], tc.pack_weights_np(plan, w), None, out)
    np.testing.assert_allclose(out[..., 8:16], ref, atol=1e-10)
    assert np.all(out[..., :8] == 0) and np.all(out[..., 16:] == 0)


@pytest.mark.parametrize("case", CASES)
def test_dgrad_lowering(case):
    N, H, W, Cs, k, s, pad, cout_real = case
    cout = 16 * (-(-cout_real // 16))  # dY channels are a GEMM-K operand here: multiples of 8 only; keep simple
    cout = cout_real if cout_real % 8 == 0 else cout
    rng = np.random.default_rng(1 + hash(case[:3]) % 1000)
    cin = sum(Cs)
    x = torch.from_numpy(rng.normal(size=(N, H, W, cin))).requires_grad_(True)
    w = rng.normal(size=(k, k, cin, cout))
    y = T.conv2d(x, torch.from_numpy(w), None, s, pad)
    dy = rng.normal(size=tuple(y.shape))
    y.backward(torch.from_numpy(dy))
    ref = x.grad.numpy()
    # gradient w.r.t. each concat source separately (channel slices of the kernel)
    c0 = 0
    for C in Cs:
        plans = tc.plan_conv_dgrad((N, H, W, C), k, s, pad, cout, cin_slice=(c0, c0 + C, cin))
        assert len(plans) == (1 if s == 1 else 4)
        dx = np.full((N, H, W, C), np.nan)
        for p in plans:
            emu.run_plan(p, [dy], tc.pack_weights_np(p, w), None, dx)
        np.testing.assert_allclose(dx, ref[..., c0:c0 + C], rtol=1e-10, atol=1e-10)
        c0 += C


def test_same_pad_rules():
    assert tc.same_pad(128, 3, 2) == (0, 1)
    assert tc.same_pad(128, 3, 1) == (1, 1)
    assert tc.same_pad(128, 7, 1) == (3, 3)
    assert tc.same_pad(130, 4, 2) == (1, 1)   # img_discr conv_0: 128 + 2*1 -> 65, effective pad (2,2)
    assert tc.same_pad(67, 4, 2) == (1, 2)    # img_discr conv_1: 65 + 2 -> 34, effective pad (2,3)
    sizes = [128]
    for _ in range(6):
        sizes.append(-(-(sizes[-1] + 2) // 2))
    assert sizes == [128, 65, 34, 18, 10, 6, 4]


@pytest.mark.parametrize("case", CASES)
def test_wgrad_lowering(case):
    N, H, W, Cs, k, s, pad, cout_real = case
    cout = cout_real if cout_real % 8 == 0 else 16 * (-(-cout_real // 16))
    rng = np.random.default_rng(2 + hash(case[:3]) % 1000)
    cin = sum(Cs)
    x = rng.normal(size=(N, H, W, cin))
    w = torch.from_numpy(rng.normal(size=(k, k, cin, cout))).requires_grad_(True)
    y = T.conv2d(torch.from_numpy(x), w, None, s, pad)
    dy = rng.normal(size=tuple(y.shape))
    y.backward(torch.from_numpy(dy))
    ref = w.grad.numpy()
    dw = np.zeros_like(ref)
    c0 = 0
    for C in Cs:
        plan = tc.plan_conv_wgrad((N, H, W, C), k, s, pad, cout, cin_slice=(c0, c0 + C, cin))
        emu.run_wgrad_plan(plan, np.ascontiguousarray(x[..., c0:c0 + C]), dy, dw)
        d = plan.desc()
        assert d.n_taps == k * k and d.Cin == C
        c0 += C
    np.testing.assert_allclose(dw, ref, rtol=1e-9, atol=1e-9)
