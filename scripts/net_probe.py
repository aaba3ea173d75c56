"""GPU probe of the full stage-1 path against the torch-CPU oracle: forward (train / inference BN), losses,
G-step and D-step gradients, then a timed training step."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

CONFIG = {"paths": {"data_dir": "", "vggnet": None, "log_dir": "/tmp/kp_logs"},
          "training": {"batch_size": 4, "lr": {"start_val": 1e-4, "step": 20000, "decay": 0.95}},
          "model": {"n_pts": 40, "n_action": 9, "cell_info": [1024, 1024], "vae_dim": 64}}


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return {"max_abs": float((a - b).abs().max()), "rel_l2": float((a - b).norm() / (b.norm() + 1e-30)),
            "cos": float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))}


def main():
    import __graft_entry__ as g
    g.build()
    import kp_b200
    from kp_b200 import models, networks, engine as E
    from oracle import networks as ON
    dev = torch.device("cuda:0")
    B = int(os.environ.get("PROBE_B", "4"))
    rng = np.random.default_rng(0)
    im = torch.from_numpy(rng.uniform(-1, 1, (B, 128, 128, 3)).astype(np.float32))
    fut = torch.from_numpy(rng.uniform(-1, 1, (B, 128, 128, 3)).astype(np.float32))
    P = ON.init_params(0, dtype=torch.float32, bias_scale=0.02)
    ON.randomize_bn(P, 1)

    model = models.DetectorTranslatorModel(CONFIG, is_training=True, device=dev)
    model.ctx.load_state_dict(P)
    model.build({"image": im.to(dev), "future_image": fut.to(dev)})
    ctx = model.ctx

    # ---------------- forward, train-mode BN ----------------
    ctx.tape, ctx.update_moving = None, False
    final = model._define_forward_pass(im.to(dev), fut.to(dev), for_G_run=True)
    torch.cuda.synchronize()
    octx = ON.Ctx({k: v.clone() for k, v in P.items()})
    t0 = time.time()
    ref = ON.forward_pass(octx, im, fut, 40, True)
    t_oracle = time.time() - t0
    out = {"fwd_train": {"final": rel(final, ref["final_output"]), "mu_cur": rel(model.current_keypoints, ref["current_pt"]),
                         "mu_fut": rel(model.future_keypoints, ref["future_pt"]), "mask": rel(model.mask, ref["mask"]),
                         "crude": rel(model.crude_output, ref["crude_output"]), "oracle_fwd_s": t_oracle}}
    print(json.dumps(out), flush=True)

    # ---------------- losses (test_step) ----------------
    lD, lG, _, _ = model.test_step()
    oD = ON.loss_D(octx, ref["final_output"], fut)
    oG = ON.loss_G(octx, ref["final_output"], fut)
    print(json.dumps({"losses": {"D": lD, "D_ref": float(oD[0]), "G": lG, "G_ref": float(oG[0]),
                                 "G_recon": model.loss_G_recon, "G_recon_ref": float(oG[1]),
                                 "G_adv": model.loss_G_adv, "G_adv_ref": float(oG[2])}}), flush=True)

    # ---------------- inference-mode forward ----------------
    model.is_training = False
    final_i = model._define_forward_pass(im.to(dev), fut.to(dev), for_G_run=True)
    octx2 = ON.Ctx({k: v.clone() for k, v in P.items()})
    ref_i = ON.forward_pass(octx2, im, fut, 40, False)
    print(json.dumps({"fwd_infer": {"final": rel(final_i, ref_i["final_output"]),
                                    "mu_cur": rel(model.current_keypoints, ref_i["current_pt"])}}), flush=True)
    model.is_training = True

    # ---------------- G-step gradients ----------------
    Pg = {k: v.clone().requires_grad_(not k.startswith("vgg") and "moving" not in k) for k, v in P.items()}
    octx3 = ON.Ctx(Pg)
    refg = ON.forward_pass(octx3, im, fut, 40, True)
    lg = ON.loss_G(octx3, refg["final_output"], fut)[0]
    lg.backward()
    ctx.tape, ctx.update_moving, ctx.train_G, ctx.train_D = E.Tape(), False, True, False
    ctx.G.grad.zero_()
    fake = model._define_forward_pass(im.to(dev), fut.to(dev), for_G_run=True)
    model._compute_loss_G(fake, fut.to(dev), backward=True)
    ctx.tape.backward()
    ctx.tape, ctx.train_G = None, False
    torch.cuda.synchronize()
    names = ["translator/conv_6_0/conv2d/kernel", "translator/conv_6_1/conv2d/bias", "translator/conv_5_1/conv2d/kernel",
             "translator/b_norm_5_1/gamma", "translator/conv_3_0/conv2d/kernel", "translator/conv_1_0/conv2d/kernel",
             "translator/b_norm_1_0/beta", "image_encoder/encoder/conv_6/conv2d/kernel", "image_encoder/encoder/conv_1/conv2d/kernel",
             "pose_encoder/conv_0/conv2d/kernel", "pose_encoder/conv_0/conv2d/bias", "pose_encoder/conv_7_1/conv2d/kernel",
             "pose_encoder/conv_5_0/conv2d/kernel", "pose_encoder/conv_3_0/conv2d/kernel", "pose_encoder/conv_1_0/conv2d/kernel",
             "pose_encoder/encoder/conv_8/conv2d/kernel", "pose_encoder/encoder/conv_5/conv2d/kernel",
             "pose_encoder/encoder/conv_1/conv2d/kernel", "pose_encoder/encoder/b_norm_3/gamma"]
    gres = {}
    for n in names:
        if Pg[n].grad is None:
            gres[n] = "no ref grad"
            continue
        r = rel(ctx.G.g(n), Pg[n].grad)
        gres[n] = {"rel_l2": round(r["rel_l2"], 4), "cos": round(r["cos"], 5)}
    print(json.dumps({"G_grads": gres}), flush=True)

    # ---------------- D-step gradients ----------------
    Pd = {k: v.clone().requires_grad_("img_discr" in k) for k, v in P.items()}
    octx4 = ON.Ctx(Pd)
    with torch.no_grad():
        refd = ON.forward_pass(octx4, im, fut, 40, True)
    ld = ON.loss_D(octx4, refd["final_output"], fut)[0]
    ld.backward()
    ctx.tape, ctx.train_D = None, False
    fake = model._define_forward_pass(im.to(dev), fut.to(dev), for_G_run=False)
    ctx.tape, ctx.train_D = E.Tape(), True
    ctx.D.grad.zero_()
    model._compute_loss_D(fake, fut.to(dev), backward=True)
    ctx.tape.backward()
    ctx.tape, ctx.train_D = None, False
    torch.cuda.synchronize()
    dres = {}
    for n in ctx.D.names():
        r = rel(ctx.D.g(n), Pd[n].grad)
        dres[n] = {"rel_l2": round(r["rel_l2"], 4), "cos": round(r["cos"], 5)}
    print(json.dumps({"D_grads": dres}), flush=True)

    # ---------------- timed training steps at B=32 ----------------
    Bt = 32
    model2 = models.DetectorTranslatorModel(CONFIG, is_training=True, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    batches = [{"image": torch.rand((Bt, 128, 128, 3), device=dev, generator=gen) * 2 - 1,
                "future_image": torch.rand((Bt, 128, 128, 3), device=dev, generator=gen) * 2 - 1} for _ in range(4)]
    it = {"i": 0}

    def feed():
        it["i"] += 1
        return batches[it["i"] % len(batches)]
    model2.build(feed)
    n0 = kp_b200._lib.load().kp_launch_count()
    for _ in range(3):
        model2.train_step()
    torch.cuda.synchronize()
    n1 = kp_b200._lib.load().kp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    steps = 10
    for _ in range(steps):
        model2.train_step()
    e1.record()
    torch.cuda.synchronize()
    wall = time.time() - t0
    ms = e0.elapsed_time(e1) / steps
    model2.train_step(should_write_log=True)
    print(json.dumps({"train_step_B32": {"ms_per_step": ms, "wall_ms_per_step": wall / steps * 1e3,
                                         "examples_per_s": Bt / (ms * 1e-3), "kernel_launches_per_step": (n1 - n0) / 3,
                                         "loss_D": model2.loss_D, "loss_G": model2.loss_G}}), flush=True)


if __name__ == "__main__":
    main()
