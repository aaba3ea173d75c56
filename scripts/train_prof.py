"""Runs a few stage-1 training steps at B=32 (for ncu launch lists / captures).  STEPS env var = number of steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import torch

from net_probe import CONFIG  # noqa: E402


def main():
    import __graft_entry__ as g
    g.build()
    from kp_b200 import models
    dev = torch.device("cuda:0")
    Bt = int(os.environ.get("BATCH", "32"))
    model = models.DetectorTranslatorModel(CONFIG, is_training=True, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    batch = {"image": torch.rand((Bt, 128, 128, 3), device=dev, generator=gen) * 2 - 1,
             "future_image": torch.rand((Bt, 128, 128, 3), device=dev, generator=gen) * 2 - 1}
    model.build(batch)
    for _ in range(int(os.environ.get("STEPS", "2"))):
        model.train_step()
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
