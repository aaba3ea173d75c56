"""Worker of tests/test_dp_gpu.py (launched with torch.distributed.run, one process per GPU; not a test module).

Data-parallel gradient equivalence: with per-replica batch-norm statistics, the all-reduced (summed)
generator / discriminator gradients of N replicas must equal the sum of the gradients ONE process computes when it runs the
N shards one after the other - and the replicas' parameters must stay bit-identical after training steps."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist


def main():
    rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import kp_b200  # noqa: F401
    from kp_b200 import models
    import test_whole_step_gpu as W
    dev = torch.device("cuda", lr)
    B = 2
    shards = [W._noise(B, seed=10 + r) for r in range(world)]
    P = W._params()
    out = {}
    model = models.DetectorTranslatorModel(W.CFG, is_training=True, device=dev)
    assert model.world == world
    model.ctx.load_state_dict(P)
    model.build({"image": shards[rank][0].to(dev), "future_image": shards[rank][1].to(dev)})
    # Run the REAL D run and G run (bucketed, overlapped collectives), recording every slice handed to the all-reduce just
    # before it is reduced: the slices must tile each flat gradient buffer exactly once, and the reduced buffer must be the
    # sum over the ranks of the recorded local gradients.  (A comparison with a second, single-process pass over the same
    # shards is not meaningful on this graph: two executions of the same step differ by tens of per cent in their
    # gradients because fp32-atomic ordering noise is amplified chaotically, tests/test_whole_step_gpu.py docstring.)
    captured = []
    real_allreduce = model._allreduce

    def recording_allreduce(buf):
        captured.append((buf.data_ptr(), buf.numel(), buf.detach().clone()))
        real_allreduce(buf)
    model._allreduce = recording_allreduce
    model.overlap_g_allreduce = True
    for which, run, grp in (("D", model._run_D, model.ctx.D), ("G", model._run_G, model.ctx.G)):
        captured.clear()
        run(shards[rank][0].to(dev), shards[rank][1].to(dev))
        model._join_D()                      # the overlapped discriminator update of the D run lands here
        torch.cuda.synchronize()
        base, total = grp.grad.data_ptr(), grp.grad.numel()
        local = torch.zeros_like(grp.grad)
        cover = torch.zeros(total, dtype=torch.int32, device=dev)
        for ptr, n, t in captured:
            if not (base <= ptr < base + 4 * total):
                continue                      # a collective on the other optimiser's buffer (overlapped D update)
            lo = (ptr - base) // 4
            local[lo:lo + n] = t
            cover[lo:lo + n] += 1
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        expect = torch.stack(gathered).sum(dim=0)
        out[which] = {"slices": len([1 for ptr, n, t in captured if base <= ptr < base + 4 * total]),
                      "covered_exactly_once": bool((cover == 1).all().item()),
                      "max_abs_diff": float((grp.grad - expect).abs().max()), "norm": float(expect.norm()),
                      "rel_l2": float((grp.grad - expect).norm() / expect.norm())}
    model._allreduce = real_allreduce
    # two real training steps (bucketed, overlapped all-reduces): replicas must stay identical
    cur = {"i": 0}

    def feed():
        cur["i"] += 1
        s = shards[(rank + cur["i"]) % world]
        return {"image": s[0].to(dev), "future_image": s[1].to(dev)}
    model.build(feed)
    for _ in range(2):
        model.train_step()
    torch.cuda.synchronize()
    skew = 0.0
    for grp in (model.ctx.G, model.ctx.D):
        g = grp.data.clone()
        dist.broadcast(g, 0)
        skew = max(skew, float((g - grp.data).abs().max()))
    t = torch.tensor([skew], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out["replica_skew_after_2_steps"] = float(t.item())
        out["world"] = world
        print("DPRESULT " + json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
