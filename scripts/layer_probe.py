"""GPU probe: train-mode forward compared with the oracle layer by layer (finds where an error enters)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from net_probe import CONFIG, rel  # noqa: E402


def main():
    import __graft_entry__ as g
    g.build()
    from kp_b200 import models
    from oracle import networks as ON
    from oracle import tf_ops as T
    dev = torch.device("cuda:0")
    B = 4
    rng = np.random.default_rng(0)
    im = torch.from_numpy(rng.uniform(-1, 1, (B, 128, 128, 3)).astype(np.float32))
    fut = torch.from_numpy(rng.uniform(-1, 1, (B, 128, 128, 3)).astype(np.float32))
    P = ON.init_params(0, dtype=torch.float32, bias_scale=0.02)
    ON.randomize_bn(P, 1)
    model = models.DetectorTranslatorModel(CONFIG, is_training=True, device=dev)
    model.ctx.load_state_dict(P)
    model.build({"image": im.to(dev), "future_image": fut.to(dev)})
    ctx = model.ctx
    ctx.debug = {}
    ctx.tape, ctx.update_moving = None, False
    model._define_forward_pass(im.to(dev), fut.to(dev), for_G_run=True)
    torch.cuda.synchronize()
    octx = ON.Ctx({k: v.clone() for k, v in P.items()})
    ON.forward_pass(octx, im, fut, 40, True)
    # note: pose_encoder runs twice; both sides keep the LAST call (future_im)
    for name, (out, y_pre, scale, shift, mean, rstd, up) in ctx.debug.items():
        ref = octx.taps.get(name)
        if ref is None:
            continue
        if up:
            ref = T.resize_bilinear_legacy(ref, 2 * ref.shape[1], 2 * ref.shape[2])
        r = rel(out.float(), ref)
        # batch statistics of the oracle's conv output for this layer
        print(json.dumps({"layer": name, "up": bool(up), "rel_l2": round(r["rel_l2"], 5), "max_abs": round(r["max_abs"], 4),
                          "mean|rstd|": round(float(rstd.abs().mean()), 3), "max_rstd": round(float(rstd.max()), 2),
                          "max|mean|*rstd": round(float((mean.abs() * rstd).max()), 2)}), flush=True)


if __name__ == "__main__":
    main()
