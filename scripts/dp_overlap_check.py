"""torchrun --nproc-per-node 2 scripts/dp_overlap_check.py : the overlapped discriminator update (side stream) must give the
same training trajectory as the serial one (up to the run-to-run noise of the fp32 atomics), eagerly and in the graph."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import torch
import torch.distributed as dist

from net_probe import CONFIG  # noqa: E402


def main():
    rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import kp_b200  # noqa: F401
    from kp_b200 import models
    dev = torch.device("cuda", lr)
    B = 8
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    batches = [{"image": torch.rand((B, 128, 128, 3), device=dev, generator=gen) * 2 - 1,
                "future_image": torch.rand((B, 128, 128, 3), device=dev, generator=gen) * 2 - 1} for _ in range(4)]

    def run(overlap, graph):
        m = models.DetectorTranslatorModel(CONFIG, is_training=True, device=dev, seed=5)
        m.overlap_d_update = overlap
        cur = {"i": -1}

        def feed():
            cur["i"] += 1
            return batches[cur["i"] % 4]
        m.build(feed)
        if graph:
            m.enable_cuda_graph(B)
        for _ in range(3):
            m.train_step()
        torch.cuda.synchronize()
        lD, lG = m._last_losses
        return m.ctx.D.data.clone(), m.ctx.G.data.clone(), float(lD.sum()), float(lG.sum())
    ref = run(False, False)
    ref2 = run(False, False)
    ov = run(True, False)
    ovg = run(True, True)
    def diff(a, b):
        return max((a[0] - b[0]).abs().max().item(), (a[1] - b[1]).abs().max().item())
    noise = diff(ref, ref2)
    d1, d2 = diff(ref, ov), diff(ref, ovg)
    # replicas must stay identical: compare parameters across ranks
    g = ov[1].clone()
    dist.broadcast(g, 0)
    rep = (g - ov[1]).abs().max().item()
    if rank == 0:
        print("noise %.3e  overlap-eager %.3e  overlap-graph %.3e  replica-skew %.3e  losses ref (%.5f, %.5f) ov (%.5f, %.5f) ovg (%.5f, %.5f)"
              % (noise, d1, d2, rep, ref[2], ref[3], ov[2], ov[3], ovg[2], ovg[3]), flush=True)
        ok = d1 <= max(10 * noise, 2e-4) and d2 <= max(10 * noise, 2e-4) and rep == 0.0
        print("OK" if ok else "MISMATCH", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
