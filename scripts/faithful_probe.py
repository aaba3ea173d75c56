"""GPU probe: CUDA path vs the bf16-faithful oracle, layer by layer (forward), then the loss-side gradient on an
identical generated frame.  Diagnostic companion of tests/test_whole_step_gpu.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return "rel_l2 %.3e  max_abs %.3e  cos %.6f" % (float((a - b).norm() / (b.norm() + 1e-300)), float((a - b).abs().max()),
                                                   float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)))


def main():
    import __graft_entry__ as g
    g.build()
    import test_whole_step_gpu as T
    from kp_b200 import models, networks, engine as E, ops
    from oracle import networks as ON, precision
    dev = torch.device("cuda:0")
    B = 2
    P = T._params()
    im, fut = T._noise(B)
    model = models.DetectorTranslatorModel(T.CFG, is_training=True, device=dev)
    model.ctx.load_state_dict(P)
    model.build({"image": im.to(dev), "future_image": fut.to(dev)})
    ctx = model.ctx
    # ---------------- forward ----------------
    ctx.begin_run()
    ctx.tape, ctx.update_moving = None, False
    ctx.debug = {}
    final = model._define_forward_pass(im.to(dev), fut.to(dev), for_G_run=True)
    torch.cuda.synchronize()
    dbg = ctx.debug
    ctx.debug = None
    Pd = {k: v.double().clone() for k, v in P.items()}
    octx = ON.Ctx(Pd, q=precision.Bf16Faithful())
    with torch.no_grad():
        ref = ON.forward_pass(octx, im.double(), fut.double(), 40, True)
    for scope, (out, y_pre, scale, shift, mean, rstd, ups) in dbg.items():
        pre = octx.taps.get(scope + ":pre")
        print("%-40s y_pre %s" % (scope, rel(y_pre.float(), pre)))
        if not ups:
            print("%-40s out   %s" % ("", rel(out.float(), octx.taps[scope])))
    print("final  ", rel(final, ref["final_output"]))
    print("mask   ", rel(model.mask, ref["mask"]))
    print("crude  ", rel(model.crude_output, ref["crude_output"]))
    print("mu_cur ", rel(model.current_keypoints, ref["current_pt"]))
    print("mu_fut ", rel(model.future_keypoints, ref["future_pt"]))

    # ---------------- loss side on an identical generated frame ----------------
    fake_ref = ref["final_output"].float()
    fake_leaf = fake_ref.double().clone().requires_grad_(True)
    octx2 = ON.Ctx(Pd, q=precision.Bf16Faithful())
    lG = ON.loss_G(octx2, fake_leaf, fut.double())
    lG[0].backward()
    ctx.begin_run()
    tape = E.Tape()
    ctx.tape, ctx.train_G, ctx.train_D = tape, True, False
    fake_dev = fake_ref.to(dev).contiguous()
    loss = model._compute_loss_G(fake_dev, fut.to(dev), backward=True)
    tape.run_closures()
    d_fake = tape.grad(fake_dev)
    torch.cuda.synchronize()
    print("loss_G cuda", loss.tolist(), "oracle", float(lG[1]), float(lG[2]))
    print("d_final (VGG + D) ", rel(d_fake, fake_leaf.grad))
    ctx.tape, ctx.train_G = None, False
    # VGG features alone
    with torch.no_grad():
        feats_ref = ON.vgg19(octx2, (fake_ref.double() + 1) / 2 * 255)
    feats = networks.vgg.features_from_prepared(networks.vgg.prepare(fake_dev, ops.VGG_PREP), need_input_grad=False)
    for i, (a, b) in enumerate(zip(feats, feats_ref)):
        print("vgg feature %d " % i, rel(a.float(), b))
    # recon-only and adv-only gradients
    for part in (1, 2):
        leaf = fake_ref.double().clone().requires_grad_(True)
        o = ON.Ctx(Pd, q=precision.Bf16Faithful())
        ON.loss_G(o, leaf, fut.double())[part].backward()
        print("oracle d_final part %d norm %.4e" % (part, float(leaf.grad.norm())))


if __name__ == "__main__":
    main()
