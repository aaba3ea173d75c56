"""Run-to-run spread of one train step (fp32 atomics in the BN statistics / weight gradients reorder) against the difference
between the single-stream and the side-stream schedule: cosines of the flat gradient buffers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import kp_b200  # noqa: F401
from kp_b200 import models

B = int(os.environ.get("B", "4"))
CFG = {"paths": {"data_dir": "", "vggnet": None, "log_dir": "/tmp/kp_b200_test"},
       "training": {"batch_size": B, "lr": {"start_val": 1e-4, "step": 20000, "decay": 0.95}},
       "model": {"n_pts": 40, "n_action": 9, "cell_info": [1024, 1024], "vae_dim": 64}}
dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
data = [{k: torch.from_numpy(rng.uniform(-1, 1, (B, 128, 128, 3)).astype(np.float32)).to(dev) for k in ("image", "future_image")}
        for _ in range(2)]


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def one_step(serial):
    for k in ("KP_WGRAD_STREAM", "KP_BRANCH_STREAM", "KP_OVERLAP_D_UPDATE"):
        if serial:
            os.environ[k] = "0"
        else:
            os.environ.pop(k, None)
    m = models.DetectorTranslatorModel(CFG, is_training=True, device=dev, seed=3)
    cur = {"i": -1}

    def feed():
        cur["i"] += 1
        return data[cur["i"] % len(data)]
    m.build(feed)
    m.train_step()
    torch.cuda.synchronize()
    lD, lG = m._last_losses
    return torch.cat([lD, lG]).cpu(), m.ctx.G.grad.clone(), m.ctx.D.grad.clone()


runs = {"s1": one_step(True), "s2": one_step(True), "o1": one_step(False), "o2": one_step(False)}
for a, b in (("s1", "s2"), ("o1", "o2"), ("s1", "o1"), ("s2", "o2"), ("s1", "o2")):
    print(a, b, "loss", runs[a][0].tolist(), runs[b][0].tolist(), "cos G %.4f D %.4f" % (cos(runs[a][1], runs[b][1]), cos(runs[a][2], runs[b][2])))
