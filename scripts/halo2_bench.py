"""Per-layer timing of the stride-1 convolution kernels at BASELINE batch: TMA-tap / cp.async-halo (round 1 default) against
the TMA-staged halo kernel (csrc/conv_halo2.cu).  CUDA events around `reps` back-to-back launches after warm-up."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

LAYERS = [
    # (tag, N, H, W, [C], k, cout, stats)
    ("pose 16->16 @128", 32, 128, 128, [16], 3, 16, True),
    ("pose 32+32->16 @128", 32, 128, 128, [32, 32], 3, 16, True),
    ("enc 32->32 @128", 32, 128, 128, [32], 3, 32, True),
    ("enc 7x1 21(32)->32 @128", 32, 128, 128, [32], (7, 1), 32, True),
    ("vgg 3x1 9(16)->64 @128", 32, 128, 128, [16], (3, 1), 64, False),
    ("dgrad 16->32 @128", 32, 128, 128, [16], 3, 32, False),
    ("pose 32->32 @64", 32, 64, 64, [32], 3, 32, True),
    ("pose 64+64->32 @64", 32, 64, 64, [64, 64], 3, 32, True),
    ("enc 64->64 @64", 32, 64, 64, [64], 3, 64, True),
    ("vgg/tr 64->64 @128", 32, 128, 128, [64], 3, 64, False),
    ("tr 128->64 @128", 32, 128, 128, [128], 3, 64, True),
    ("dgrad 64->128 @128", 32, 128, 128, [64], 3, 128, False),
    ("vgg 64->128 @64", 32, 64, 64, [64], 3, 128, False),
    ("tr 128->128 @64", 32, 64, 64, [128], 3, 128, True),
    ("tr 256->128 @64", 32, 64, 64, [256], 3, 128, True),
    ("dgrad 128->256 @64", 32, 64, 64, [128], 3, 256, False),
    ("tr 64->4 heads @128", 32, 128, 128, [64], 3, 4, False),
    ("pose 64->64 @32", 32, 32, 32, [64], 3, 64, True),
    ("enc 128->128 @32", 32, 32, 32, [128], 3, 128, True),
    ("tr 256->256 @32", 32, 32, 32, [256], 3, 256, True),
    ("vgg 256->256 @32", 32, 32, 32, [256], 3, 256, False),
    ("pose 128->128 @16", 32, 16, 16, [128], 3, 128, True),
    ("vgg 512->512 @16", 32, 16, 16, [512], 3, 512, False),
]


def main():
    import __graft_entry__ as g
    g.build()
    from kp_b200 import conv, tapconv as tc
    dev = torch.device("cuda:0")
    reps = int(os.environ.get("REPS", "20"))
    only = os.environ.get("ONLY")
    print("%-28s %10s %10s %8s   %s" % ("layer", "base us", "halo2 us", "speedup", "TFLOP/s base -> halo2"))
    layers = LAYERS
    if os.environ.get("BATCH_SWEEP"):
        layers = [("%s N=%d" % (tag, nn), nn, H, W, Cs, k, cout, stats) for tag, N, H, W, Cs, k, cout, stats in LAYERS
                  for nn in (4, 8, 16, 32, 64) if (only and only in tag)]
    for tag, N, H, W, Cs, k, cout, stats in layers:
        if only and only not in tag:
            continue
        kh, kw = (k, k) if isinstance(k, int) else k
        cin = sum(Cs)
        xs = [torch.randn((N, H, W, C), device=dev).to(torch.bfloat16) for C in Cs]
        w = torch.randn((kh, kw, cin, cout), device=dev) / (kh * kw * cin) ** 0.5
        plan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape) for x in xs], k, 1, 0, cout)
        wp = conv.pack_weights(plan, w)
        bias = torch.zeros(plan.rows_pad, device=dev)
        out = torch.empty((n, ho, wo, cout), device=dev, dtype=torch.bfloat16)
        st = (torch.zeros(plan.rows_pad, device=dev), torch.zeros(plan.rows_pad, device=dev)) if stats else None
        flops = 2.0 * N * H * W * kh * kw * cin * cout
        res = []
        outs = []
        for mode in ("0", "2"):
            os.environ["KP_TAPCONV_HALO2"] = mode
            try:
                for _ in range(3):
                    conv.run_plan(plan, xs, wp, bias, out, act=tc.ACT_NONE if stats else tc.ACT_RELU, stats=st)
                torch.cuda.synchronize()
                # the launches are captured into a CUDA graph: the Python-side descriptor building (~20 us per call) would
                # otherwise bound every layer faster than that
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for _ in range(reps):
                        conv.run_plan(plan, xs, wp, bias, out, act=tc.ACT_NONE if stats else tc.ACT_RELU, stats=st)
                gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) / reps * 1e3)
                outs.append(out.float().clone())
            except Exception as e:   # noqa
                res.append(float("nan"))
                outs.append(None)
                print("   %s mode %s failed: %s" % (tag, mode, str(e)[:200]))
        diff = float((outs[0] - outs[1]).abs().max() / (outs[0].abs().max() + 1e-30)) if outs[0] is not None and outs[1] is not None else float("nan")
        print("%-28s %10.1f %10.1f %8.2f   %6.0f -> %6.0f   maxdiff %.1e" % (tag, res[0], res[1], res[0] / res[1], flops / res[0] * 1e-6,
                                                                          flops / res[1] * 1e-6, diff), flush=True)
    os.environ.pop("KP_TAPCONV_HALO2", None)


if __name__ == "__main__":
    main()
