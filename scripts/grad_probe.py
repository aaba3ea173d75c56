"""GPU probe: dgrad (tap-GEMM with transposed weights) and wgrad (MN-major tcgen05) against fp64 autograd of the
conv oracle, one case per subprocess; then a throughput microbench of representative layers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (N, H, W, [C], k, stride, pad, cout)
    "g_cb64_256": (2, 32, 32, [256], 3, 1, 0, 256),
    "g_cb64_128to64": (2, 32, 32, [128], 3, 1, 0, 64),
    "g_cb32": (2, 64, 64, [32], 3, 1, 0, 32),
    "g_cb16": (1, 128, 128, [16], 3, 1, 0, 16),
    "g_concat": (2, 32, 32, [128, 128], 3, 1, 0, 64),
    "g_head_1x1": (2, 128, 128, [16], 1, 1, 0, 40),
    "g_enc7x7": (1, 128, 128, [16], 7, 1, 0, 32),
    "g_s2_cb32": (2, 128, 128, [32], 3, 2, 0, 64),
    "g_s2_cb64": (2, 64, 64, [64], 3, 2, 0, 128),
    "g_discr_odd": (2, 65, 65, [64], 4, 2, 1, 128),
    "g_discr_img": (2, 128, 128, [16], 4, 2, 1, 64),
    "g_vgg512": (4, 8, 8, [512], 3, 1, 0, 512),
    "g_cout8": (1, 128, 128, [64], 3, 1, 0, 8),
}


def run_case(name):
    import numpy as np
    import torch
    import kp_b200  # noqa: F401
    from kp_b200 import conv, tapconv as tc
    from oracle import tf_ops as T
    N, H, W, Cs, k, s, pad, cout = CASES[name]
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(abs(hash(name)) % 10000)
    cin = sum(Cs)
    xs = [torch.from_numpy(rng.normal(size=(N, H, W, C)).astype(np.float32)).to(dev).to(torch.bfloat16) for C in Cs]
    w = torch.from_numpy((rng.normal(size=(k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)).to(dev)
    fplan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape) for x in xs], k, s, pad, cout)
    cpad = tc.round_up(cout, 8)
    dy = torch.zeros((n, ho, wo, cpad), device=dev, dtype=torch.bfloat16)
    dy[..., :cout] = torch.from_numpy(rng.normal(size=(n, ho, wo, cout)).astype(np.float32)).to(dev).to(torch.bfloat16)
    # oracle (fp64 on bf16-rounded operands)
    x64 = torch.cat([x.double().cpu() for x in xs], dim=-1).requires_grad_(True)
    w64 = w.to(torch.bfloat16).double().cpu().requires_grad_(True)
    y = T.conv2d(x64, w64, None, s, pad)
    y.backward(dy[..., :cout].double().cpu())
    res = {"case": name}
    # dgrad per source; dY is stored with cpad channels, the kernel reads the first `cout` (extra are zero)
    wpad = torch.nn.functional.pad(w, (0, cpad - cout))
    c0 = 0
    errs = []
    for x in xs:
        C = x.shape[3]
        plans = tc.plan_conv_dgrad(tuple(x.shape), k, s, pad, cpad, cin_slice=(c0, c0 + C, cin))
        dx = torch.full(tuple(x.shape), float("nan"), device=dev, dtype=torch.bfloat16)
        for p in plans:
            conv.run_plan(p, [dy], conv.pack_weights(p, wpad), None, dx)
        torch.cuda.synchronize()
        ref = x64.grad[..., c0:c0 + C]
        errs.append(((dx.double().cpu() - ref).abs().max() / ref.abs().max()).item())
        res["dgrad_nan"] = int(torch.isnan(dx.float()).sum().item())
        c0 += C
    res["dgrad_err"] = max(errs)
    # wgrad
    dw = torch.zeros((k, k, cin, cpad), device=dev, dtype=torch.float32)
    c0 = 0
    for x in xs:
        C = x.shape[3]
        wp = tc.plan_conv_wgrad(tuple(x.shape), k, s, pad, cpad, cin_slice=(c0, c0 + C, cin))
        conv.run_wgrad(wp, x, dy, dw)
        c0 += C
    torch.cuda.synchronize()
    ref = w64.grad
    got = dw[..., :cout].double().cpu()
    res["wgrad_err"] = ((got - ref).abs().max() / ref.abs().max()).item()
    res["wgrad_pad_abs"] = dw[..., cout:].abs().max().item() if cpad > cout else 0.0
    print(json.dumps(res), flush=True)


PERF = {
    # name: (N, H, W, [C], k, stride, cout)  at training batch 32 (VGG sees 64 images)
    "trans_32x32_256": (32, 32, 32, [256], 3, 1, 256),
    "trans_64x64_128": (32, 64, 64, [128], 3, 1, 128),
    "trans_128_64": (32, 128, 128, [64], 3, 1, 64),
    "vgg_128_64": (64, 128, 128, [64], 3, 1, 64),
    "vgg_64_128": (64, 64, 64, [128], 3, 1, 128),
    "vgg_32_256": (64, 32, 32, [256], 3, 1, 256),
    "vgg_16_512": (64, 16, 16, [512], 3, 1, 512),
    "det_128_32": (32, 128, 128, [32], 3, 1, 32),
    "det_128_16": (32, 128, 128, [16], 3, 1, 16),
    "enc_s2_64": (32, 128, 128, [32], 3, 2, 64),
    "head_128_16_40": (32, 128, 128, [16], 1, 1, 40),
    "det_128_48_16": (32, 128, 128, [16, 32], 3, 1, 16),
    "vggfirst_128_16_64": (64, 128, 128, [16], 3, 1, 64),
    # W-packed equivalents (P pixels folded into channels): what a 16->16 / 32->32 / 48->16 / head layer would cost
    "packed4_16_16": (32, 128, 32, [64], 3, 1, 64),
    "packed2_32_32": (32, 128, 64, [64], 3, 1, 64),
    "packed4_48_16": (32, 128, 32, [64, 128], 3, 1, 64),
    "packed4_head": (32, 128, 32, [64], 1, 1, 160),
    "packed2_32_32_64sq": (32, 64, 32, [64], 3, 1, 64),
    "det_64_32": (32, 64, 64, [32], 3, 1, 32),
    "trans5_128_128_64": (32, 128, 128, [128], 3, 1, 64),
    "heads_128_64_8": (32, 128, 128, [64], 3, 1, 8),
    "enc1_128_32_32_7x1": (32, 128, 128, [32], (7, 1), 1, 32),
    "dec_64_96_32": (32, 64, 64, [32, 64], 3, 1, 32),
    "dec_64_32_32": (32, 64, 64, [32], 3, 1, 32),
}


def run_perf():
    import torch
    import kp_b200  # noqa: F401
    from kp_b200 import conv, tapconv as tc
    dev = torch.device("cuda:0")
    only = os.environ.get("PERF_ONLY")
    reps_env = int(os.environ.get("PERF_REPS", "20"))
    for name, (N, H, W, Cs, k, s, cout) in PERF.items():
        if only and name not in only.split(","):
            continue
        xs = [torch.randn((N, H, W, C), device=dev).to(torch.bfloat16) for C in Cs]
        cin = sum(Cs)
        kh, kw = (k, k) if isinstance(k, int) else k
        w = torch.randn((kh, kw, cin, cout), device=dev) / (kh * kw * cin) ** 0.5
        plan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape) for x in xs], k, s, 0, cout)
        wp = conv.pack_weights(plan, w)
        out = torch.empty((n, ho, wo, cout), device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(plan.rows_pad, device=dev)
        dy = torch.randn((n, ho, wo, cout), device=dev).to(torch.bfloat16)
        wplan = tc.plan_conv_wgrad(tuple(xs[0].shape), k, s, 0, cout, cin_slice=(0, Cs[0], cin) if len(Cs) > 1 else None)
        dw = torch.zeros((kh, kw, cin, cout), device=dev)

        def timeit(fn, reps=reps_env):
            for _ in range(3 if reps > 1 else 1):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        t_f = timeit(lambda: conv.run_plan(plan, xs, wp, bias, out, act=tc.ACT_RELU))
        extra = {}
        if os.environ.get("PERF_VARIANTS"):
            ssum = torch.zeros(plan.rows_pad, device=dev)
            ssq = torch.zeros(plan.rows_pad, device=dev)
            out32 = torch.empty((n, ho, wo, cout), device=dev, dtype=torch.float32)
            extra["fwd_stats_ms"] = timeit(lambda: conv.run_plan(plan, xs, wp, None, out, stats=(ssum, ssq)))
            extra["fwd_nobias_noact_ms"] = timeit(lambda: conv.run_plan(plan, xs, wp, None, out))
            extra["fwd_f32_ms"] = timeit(lambda: conv.run_plan(plan, xs, wp, bias, out32))
        t_w = timeit(lambda: conv.run_wgrad(wplan, xs[0], dy, dw))
        flop = 2.0 * n * ho * wo * kh * kw * cin * cout
        byt = 2.0 * (sum(x.numel() for x in xs) + out.numel())
        print(json.dumps({"perf": name, "fwd_ms": t_f, "fwd_tflops": flop / t_f / 1e9, "fwd_gbs_min": byt / t_f / 1e6,
                          "wgrad_ms": t_w, "wgrad_tflops": flop / t_w / 1e9, **extra}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "perf":
        run_perf()
    elif len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        import __graft_entry__ as g
        g.build()
        for name in list(CASES) + ["perf"]:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                                   timeout=240)
                if r.returncode != 0:
                    print(json.dumps({"case": name, "rc": r.returncode, "stderr": r.stderr[-800:],
                                      "stdout": r.stdout[-300:]}), flush=True)
                else:
                    print(r.stdout.strip(), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"case": name, "timeout": True}), flush=True)
