"""Runs kp_augment_frames a few times at bench size (for ncu: `-k regex:augment_kernel`)."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kp_b200  # noqa: E402,F401
from kp_b200 import augment as A  # noqa: E402

dev = torch.device("cuda:0")
n, w, h = 2048, 320, 240
src = torch.randint(0, 256, (n * w * h * 3,), device=dev, dtype=torch.uint8)
rnd = random.Random(5)
table = A.PlanTable(n)
for i in range(n):
    fid = rnd.randint(0, 9)
    fac = {6: rnd.randint(0, 50), 7: rnd.randint(7, 20), 8: rnd.randint(0, 50), 9: rnd.randint(7, 30)}.get(fid, 0) * 0.1
    table.set(i, i * w * h * 3, w, h, 170, 128, rnd.randint(0, 42), 0, rnd.randrange(-10, 11), rnd.randint(0, 1), fid, fac)
plans = table.host.to(dev)
out = torch.empty((n, 128, 128, 3), device=dev)
for _ in range(4):
    A.augment_frames(src, plans, n, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum()))
