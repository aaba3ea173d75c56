"""KP_TAPCONV_2CTA=1: the CTA-pair kernel against the single-CTA kernel and the fp64 oracle (experimental path)."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kp_b200  # noqa
from kp_b200 import conv, tapconv as tc
from oracle import tf_ops as T

dev = torch.device("cuda:0")
BF = torch.bfloat16
for (N, H, W, Cs, k, cout, stats) in [(16, 32, 32, [256], 3, 256, False), (16, 32, 32, [128], 3, 128, True), (9, 32, 32, [64, 64], 3, 512, False),
                                      (32, 16, 16, [512], 3, 512, False)]:
    rng = np.random.default_rng(N + cout)
    cin = sum(Cs)
    xs = [torch.from_numpy(rng.normal(size=(N, H, W, C)).astype(np.float32)).to(BF) for C in Cs]
    w = torch.from_numpy((rng.normal(size=(k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)).to(BF).float()
    b = torch.from_numpy(rng.normal(0, 0.2, cout).astype(np.float32))
    plan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape) for x in xs], k, 1, 0, cout)
    srcs = [x.to(dev) for x in xs]
    wp = conv.pack_weights(plan, w.to(dev))
    bias = conv.pad_vec(b.to(dev), plan.rows_pad)

    def run(two):
        os.environ["KP_TAPCONV_2CTA"] = "1" if two else "0"
        os.environ["KP_TAPCONV_HALO"] = "0"
        out = torch.full((n, ho, wo, cout), float("nan"), device=dev, dtype=BF)
        st = (torch.zeros(plan.rows_pad, device=dev), torch.zeros(plan.rows_pad, device=dev)) if stats else None
        conv.run_plan(plan, srcs, wp, None if stats else bias, out, act=tc.ACT_NONE if stats else tc.ACT_RELU, stats=st)
        torch.cuda.synchronize()
        return out.float().cpu(), st
    y2, s2 = run(True)
    y1, s1 = run(False)
    ref = T.conv2d(torch.cat([x.double() for x in xs], dim=-1), w.double(), None if stats else b.double(), 1, 0)
    if not stats:
        ref = torch.relu(ref)
    scale = ref.abs().max().item()
    print((N, H, W, Cs, cout), "2cta vs oracle %.2e  1cta vs oracle %.2e  2cta vs 1cta %.2e  nan %d" % (
        (y2.double() - ref).abs().max().item() / scale, (y1.double() - ref).abs().max().item() / scale,
        (y2 - y1).abs().max().item() / scale, int(torch.isnan(y2).sum())),
        "" if not stats else "stats diff %.2e" % ((s2[0] - s1[0]).abs().max().item() / s1[0].abs().max().item()), flush=True)
