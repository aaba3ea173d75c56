"""GPU probe: runs tap-GEMM conv cases one per subprocess (a trap in one case does not hide the others)
and prints one JSON line per case with the max relative error against the fp64 conv oracle."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (N, H, W, [C], k, stride, pad, cout, act, out_f32, stats)
    "cb64_bn64_3x3": (2, 32, 32, [64], 3, 1, 0, 64, "relu", False, False),
    "cb32_bn32_3x3": (2, 64, 64, [32], 3, 1, 0, 32, "relu", False, False),
    "cb16_bn16_3x3": (1, 128, 128, [16], 3, 1, 0, 16, "none", False, False),
    "cb64_bn256": (2, 32, 32, [256], 3, 1, 0, 256, "relu", False, True),
    "cb64_bn128_concat": (2, 32, 32, [128, 128], 3, 1, 0, 64, "relu", False, False),
    "cb32_concat_bn16": (1, 128, 128, [32, 32], 3, 1, 0, 16, "relu", False, True),
    "head_1x1_f32": (2, 128, 128, [16], 1, 1, 0, 40, "none", True, False),
    "enc_7x7": (2, 128, 128, [16], 7, 1, 0, 32, "none", False, True),
    "s2_3x3_cb32": (2, 128, 128, [32], 3, 2, 0, 64, "none", False, False),
    "s2_3x3_cb64": (2, 64, 64, [64], 3, 2, 0, 128, "relu", False, True),
    "discr_4x4_s2_odd": (2, 65, 65, [64], 4, 2, 1, 128, "leaky", False, False),
    "discr_logit": (2, 4, 4, [2048], 3, 1, 1, 1, "none", True, False),
    "vgg_512_8x8": (4, 8, 8, [512], 3, 1, 0, 512, "relu", False, False),
    "heads_sigmoid_last": (1, 128, 128, [64], 3, 1, 0, 4, "sigmoid_last", True, False),
    "c40_tail_block": (2, 32, 32, [128, 40, 40], 3, 1, 0, 256, "none", False, False),
}


def run_case(name):
    import numpy as np
    import torch
    import kp_b200  # noqa: F401
    from kp_b200 import conv, tapconv as tc
    from oracle import tf_ops as T
    N, H, W, Cs, k, s, pad, cout, act, out_f32, stats = CASES[name]
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(abs(hash(name)) % 10000)
    xs = [torch.from_numpy(rng.normal(size=(N, H, W, C)).astype(np.float32)).to(dev).to(torch.bfloat16) for C in Cs]
    cin = sum(Cs)
    w = torch.from_numpy((rng.normal(size=(k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)).to(dev)
    b = torch.from_numpy(rng.normal(size=(cout,)).astype(np.float32)).to(dev)
    plan, (n, ho, wo) = tc.plan_conv_fwd([tuple(x.shape) for x in xs], k, s, pad, cout)
    wp = conv.pack_weights(plan, w)
    out = torch.full((n, ho, wo, cout), float("nan"), device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    st = None
    if stats:
        st = (torch.zeros(plan.rows_pad, device=dev), torch.zeros(plan.rows_pad, device=dev))
    actmap = {"none": tc.ACT_NONE, "relu": tc.ACT_RELU, "leaky": tc.ACT_LEAKY, "sigmoid_last": tc.ACT_SIGMOID_LAST}
    conv.run_plan(plan, xs, wp, conv.pad_vec(b, plan.rows_pad), out, act=actmap[act], alpha=0.01, stats=st)
    torch.cuda.synchronize()
    # oracle on the bf16-rounded operands, fp64
    x64 = torch.cat([x.double().cpu() for x in xs], dim=-1)
    w64 = w.to(torch.bfloat16).double().cpu()
    pre = T.conv2d(x64, w64, None, s, pad)
    ref = pre + b.double().cpu()
    if act == "relu":
        ref = torch.relu(ref)
    elif act == "leaky":
        ref = T.leaky_relu(ref, 0.01)
    elif act == "sigmoid_last":
        ref = torch.cat([ref[..., :-1], torch.sigmoid(ref[..., -1:])], dim=-1)
    got = out.double().cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    res = {"case": name, "max_err_over_scale": err, "nan": int(torch.isnan(got).sum().item()), "shape": list(got.shape)}
    if stats:
        s1 = pre.sum(dim=(0, 1, 2))
        s2 = (pre * pre).sum(dim=(0, 1, 2))
        res["stats_sum_err"] = ((st[0][:cout].double().cpu() - s1).abs().max() / (s1.abs().max() + 1e-9)).item()
        res["stats_sq_err"] = ((st[1][:cout].double().cpu() - s2).abs().max() / s2.abs().max()).item()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        import __graft_entry__ as g
        g.build()
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                                   timeout=180)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                if r.returncode != 0:
                    print(json.dumps({"case": name, "rc": r.returncode, "stderr": r.stderr[-600:], "stdout": tail}), flush=True)
                else:
                    print(tail, flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"case": name, "timeout": True}), flush=True)
