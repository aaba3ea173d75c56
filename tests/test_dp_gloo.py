"""CPU, world_size 2 over gloo: the host-side data-parallel logic (flat-buffer gradient all-reduce with the 1/world
scaling folded into Adam, parameter broadcast, frame sharding) gives the same update as one process that sees
both micro-batches."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import kp_b200  # noqa: F401
from kp_b200 import dp, engine as E
from oracle import tf_ops as T


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_group():
    g = E.ParamGroup(torch.device("cpu"))
    g.add("a/conv2d/kernel", (3, 3, 4, 8))
    g.add("a/conv2d/bias", (8,))
    g.add("b/gamma", (5,))          # odd size: exercises the 16-byte alignment padding between tensors
    g.add("c/conv2d/kernel", (1, 1, 8, 2))
    g.finalize()
    return g


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = _make_group()
        rng = np.random.default_rng(7)               # identical initial parameters on every rank ...
        g.data.copy_(torch.from_numpy(rng.normal(size=g.total).astype(np.float32)))
        if rank == 1:
            g.data.add_(1.0)                          # ... except that rank 1 drifted: broadcast must repair it

        class Ctx:
            G = g
            D = S = V = E.ParamGroup(torch.device("cpu"), False)

            def params_changed(self):
                pass
        for grp in (Ctx.D,):
            grp.finalize()
        dp.broadcast_parameters(Ctx(), src=0)
        grng = np.random.default_rng(100 + rank)      # rank-specific micro-batch gradient
        g.grad.copy_(torch.from_numpy(grng.normal(size=g.total).astype(np.float32)))
        dp.allreduce_sum_(g.grad)
        # the Adam kernel's grad_scale = 1/world (here emulated with the oracle update on the flat buffer)
        p, m, v = T.adam_tf(g.data.double(), g.grad.double() / world, g.m.double(), g.v.double(), 1, 1e-3)
        out[rank] = (p.numpy(), g.grad.numpy().copy(), dp.shard_range(10, rank, world))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    p0, g0, s0 = out[0]
    p1, g1, s1 = out[1]
    np.testing.assert_array_equal(p0, p1)             # replicas stay bit-identical
    np.testing.assert_array_equal(g0, g1)
    # single-process reference: both micro-batch gradients summed, averaged, one Adam step
    g = _make_group()
    rng = np.random.default_rng(7)
    data = torch.from_numpy(rng.normal(size=g.total).astype(np.float32))
    grads = [torch.from_numpy(np.random.default_rng(100 + r).normal(size=g.total).astype(np.float32)) for r in range(world)]
    np.testing.assert_allclose(g0, (grads[0] + grads[1]).numpy(), rtol=1e-6)
    p, _, _ = T.adam_tf(data.double(), (grads[0] + grads[1]).double() / world, torch.zeros(g.total).double(),
                        torch.zeros(g.total).double(), 1, 1e-3)
    np.testing.assert_allclose(p0, p.numpy(), rtol=1e-12)
    assert s0 == (0, 5) and s1 == (5, 10)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_param_group_views_are_aligned_and_disjoint():
    g = _make_group()
    seen = 0
    for name, shape, off in g.specs:
        assert off % 4 == 0 and off >= seen
        seen = off + int(np.prod(shape))
        g.p(name).fill_(1.0)
    assert g.total % 4 == 0
    assert int(g.data.sum().item()) == sum(int(np.prod(s)) for _, s, _ in g.specs)
