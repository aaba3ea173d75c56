"""One launch sequence of a single stride-1 layer through the halo kernel (for ncu captures): LAYER=N,H,W,Cin,Cout[,stats]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kp_b200  # noqa: F401
from kp_b200 import conv, tapconv as tc

dev = torch.device("cuda:0")
spec = [int(v) for v in os.environ.get("LAYER", "32,128,128,64,64,0").split(",")]
N, H, W, C, cout = spec[:5]
stats = len(spec) > 5 and spec[5] != 0
x = torch.randn((N, H, W, C), device=dev).to(torch.bfloat16)
w = torch.randn((3, 3, C, cout), device=dev) / (9 * C) ** 0.5
plan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape)], 3, 1, 0, cout)
wp = conv.pack_weights(plan, w)
out = torch.empty((n, ho, wo, cout), device=dev, dtype=torch.bfloat16)
st = (torch.zeros(plan.rows_pad, device=dev), torch.zeros(plan.rows_pad, device=dev)) if stats else None
for i in range(3):
    conv.run_plan(plan, [x], wp, None, out, stats=st, act=conv.ACT_LEAKY, alpha=0.2)
torch.cuda.synchronize()
