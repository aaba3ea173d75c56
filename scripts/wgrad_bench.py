"""Per-layer timing of the weight-gradient kernels at BASELINE batch: per-tap kernel (csrc/conv_wgrad.cu) against the halo-tile
kernel (csrc/conv_wgrad2.cu).  Launches captured in a CUDA graph, CUDA events around the replay."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

LAYERS = [
    ("tr 128->64 @128", 32, 128, 128, 128, 64, 3),
    ("tr/vgg 64->64 @128", 32, 128, 128, 64, 64, 3),
    ("tr 128->128 @64", 32, 64, 64, 128, 128, 3),
    ("tr 256->128 @64", 32, 64, 64, 256, 128, 3),
    ("tr 256->256 @32", 32, 32, 32, 256, 256, 3),
    ("pose 32->16 @128 (N64)", 64, 128, 128, 32, 16, 3),
    ("pose 32->32 @128 (N64)", 64, 128, 128, 32, 32, 3),
    ("pose 16->16 @128 (N64)", 64, 128, 128, 16, 16, 3),
    ("enc 7x1 32->32 @128 (N64)", 64, 128, 128, 32, 32, (7, 1)),
    ("pose 64->32 @64 (N64)", 64, 64, 64, 64, 32, 3),
    ("pose 32->32 @64 (N64)", 64, 64, 64, 32, 32, 3),
    ("enc 64->64 @64 (N64)", 64, 64, 64, 64, 64, 3),
    ("pose 64->64 @32 (N64)", 64, 32, 32, 64, 64, 3),
    ("pose 128->128 @32 (N64)", 64, 32, 32, 128, 128, 3),
    ("pose 128->128 @16 (N64)", 64, 16, 16, 128, 128, 3),
    ("enc 256->256 @16 (N64)", 64, 16, 16, 256, 256, 3),
]


def main():
    import __graft_entry__ as g
    g.build()
    from kp_b200 import conv, tapconv as tc
    dev = torch.device("cuda:0")
    reps = int(os.environ.get("REPS", "10"))
    print("%-28s %10s %10s %8s   %s" % ("layer", "tap us", "halo us", "speedup", "TFLOP/s tap -> halo"))
    only = os.environ.get("ONLY")
    for tag, N, H, W, cin, cout, k in LAYERS:
        if only and only not in tag:
            continue
        kh, kw = (k, k) if isinstance(k, int) else k
        x = torch.randn((N, H, W, cin), device=dev).to(torch.bfloat16)
        dy = torch.randn((N, H, W, cout), device=dev).to(torch.bfloat16)
        plan = tc.plan_conv_wgrad((N, H, W, cin), k, 1, 0, cout)
        dw = torch.zeros((kh, kw, cin, cout), device=dev)
        flops = 2.0 * N * H * W * kh * kw * cin * cout
        res, outs = [], []
        for mode in ("0", "2"):
            os.environ["KP_WGRAD_HALO"] = mode
            for _ in range(2):
                conv.run_wgrad(plan, x, dy, dw)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(reps):
                    conv.run_wgrad(plan, x, dy, dw)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / reps * 1e3)
            dw.zero_()
            conv.run_wgrad(plan, x, dy, dw)
            outs.append(dw.clone())
            dw.zero_()
        diff = float((outs[0] - outs[1]).abs().max() / (outs[0].abs().max() + 1e-30))
        print("%-28s %10.1f %10.1f %8.2f   %6.0f -> %6.0f   maxdiff %.1e" % (tag, res[0], res[1], res[0] / res[1], flops / res[0] * 1e-6,
                                                                          flops / res[1] * 1e-6, diff), flush=True)
    os.environ.pop("KP_WGRAD_HALO", None)


if __name__ == "__main__":
    main()
