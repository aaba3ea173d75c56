"""Isolated bandwidth of the batch-norm streaming kernels (CUDA graph of `reps` launches, CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    import __graft_entry__ as g
    g.build()
    from kp_b200 import ops
    dev = torch.device("cuda:0")
    print("%-26s %10s %10s %10s   (GB/s: apply 2 transfers, reduce 2, bwd-apply 3)" % ("shape", "apply", "bwd", "copy"))
    for N, H, W, C in [(32, 128, 128, 64), (64, 128, 128, 32), (64, 128, 128, 16), (32, 64, 64, 128), (32, 32, 32, 256), (64, 16, 16, 128)]:
        x = torch.randn((N, H, W, C), device=dev).to(torch.bfloat16)
        dout = torch.randn((N, H, W, C), device=dev).to(torch.bfloat16)
        ssum = torch.zeros(C, device=dev); ssq = torch.full((C,), float(N * H * W), device=dev)
        gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
        out, scale, shift, mean, rstd = ops.bn_stats_apply(x, ssum, ssq, None, gamma, beta, N * H * W, None, None)
        nbytes = x.numel() * 2
        t_apply = timed(lambda: ops.bn_stats_apply(x, ssum, ssq, None, gamma, beta, N * H * W, None, None))
        z = torch.zeros(2 * C, device=dev)
        t_bwd = timed(lambda: ops.bn_act_bwd(dout, x, scale, shift, mean, rstd, zeroed=z))
        y = torch.empty_like(x)
        t_copy = timed(lambda: y.copy_(x))
        print("%-26s %7.1f us %7.1f us %7.1f us   apply %5.0f  bwd(5 transfers) %5.0f  copy %5.0f" %
              ("N%d %dx%d C%d" % (N, H, W, C), t_apply, t_bwd, t_copy, 2 * nbytes / t_apply * 1e-3, 5 * nbytes / t_bwd * 1e-3,
               2 * nbytes / t_copy * 1e-3), flush=True)


if __name__ == "__main__":
    main()
