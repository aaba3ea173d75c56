"""Kernel timeline of ONE replay of the captured train-step graph (torch.profiler / CUPTI): per-kernel in-situ durations,
idle gaps between kernels, totals per kernel family.  Output: gpurun_out/graph_trace.json (summary) + .csv (all kernels)."""
import collections
import json
import os
import re
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
    B = int(os.environ.get("BATCH", "32"))
    model = models.DetectorTranslatorModel(CONFIG, is_training=True, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    batch = {"image": torch.rand((B, 128, 128, 3), device=dev, generator=gen) * 2 - 1,
             "future_image": torch.rand((B, 128, 128, 3), device=dev, generator=gen) * 2 - 1}
    model.build(batch)
    model.enable_cuda_graph(B)
    for _ in range(5):
        model.train_step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            model.train_step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    # take the last replay: kernels after the last adam of the previous step
    names = [k[2] for k in ks]
    adam = [i for i, n in enumerate(names) if "adam_tf" in n]
    per_step = len(adam) // 3 if adam else 0
    start = adam[-per_step - 1] + 1 if per_step and len(adam) > per_step else 0
    step = ks[start:adam[-1] + 1] if adam else ks
    t0, t1 = step[0][0], step[-1][1]
    busy = sum(e - s for s, e, _ in step)
    gaps = [max(0.0, step[i + 1][0] - step[i][1]) for i in range(len(step) - 1)]
    fam = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in step:
        key = re.sub(r"\(.*", "", n)
        key = re.sub(r"^void ", "", key)[:60]
        fam[key][0] += 1
        fam[key][1] += e - s
    out = {"kernels": len(step), "span_us": t1 - t0, "busy_us": busy, "gap_us": sum(gaps),
           "gap_mean_us": sum(gaps) / max(1, len(gaps)), "gaps_over_5us": sum(1 for x in gaps if x > 5),
           "families": sorted(([k, v[0], round(v[1], 1)] for k, v in fam.items()), key=lambda r: -r[2])[:30]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "graph_trace.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "graph_trace.csv"), "w") as fh:
        fh.write("start_us,dur_us,gap_before_us,name\n")
        prev = None
        for s, e, n in step:
            fh.write("%.2f,%.2f,%.2f,%s\n" % (s - t0, e - s, 0.0 if prev is None else s - prev, n.replace(",", ";")[:120]))
            prev = e
    print(json.dumps(out)[:3000])


if __name__ == "__main__":
    main()
