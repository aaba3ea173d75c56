"""GPU: the side streams of the train step (D run beside the G run, weight gradients, image_encoder / D(fake) branches)
change the schedule, not the result: one step with all of them against one step on a single stream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = {"paths": {"data_dir": "", "vggnet": None, "log_dir": "/tmp/kp_b200_test"},
       "training": {"batch_size": 4, "lr": {"start_val": 1e-4, "step": 20000, "decay": 0.95}},
       "model": {"n_pts": 40, "n_action": 9, "cell_info": [1024, 1024], "vae_dim": 64}}


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.parametrize("graph", [False, True])
def test_side_streams_do_not_change_the_step(cuda_dev, monkeypatch, graph):
    from kp_b200 import models
    rng = np.random.default_rng(5)
    data = [{k: torch.from_numpy(rng.uniform(-1, 1, (4, 128, 128, 3)).astype(np.float32)).to(cuda_dev) for k in ("image", "future_image")}
            for _ in range(2)]

    def one_step(serial):
        for k in ("KP_WGRAD_STREAM", "KP_BRANCH_STREAM", "KP_OVERLAP_D_UPDATE"):
            if serial:
                monkeypatch.setenv(k, "0")
            else:
                monkeypatch.delenv(k, raising=False)
        m = models.DetectorTranslatorModel(CFG, is_training=True, device=cuda_dev, seed=3)
        assert (m.ctx.wgrad_stream is None) == serial and (m.ctx.branch_stream is None) == serial and m.overlap_d_update != serial
        cur = {"i": -1}

        def feed():
            cur["i"] += 1
            return data[cur["i"] % len(data)]
        m.build(feed)
        if graph:
            m.enable_cuda_graph(4)
        m.train_step()
        torch.cuda.synchronize()
        lD, lG = m._last_losses
        return (torch.cat([lD, lG]).cpu(), m.ctx.G.grad.clone(), m.ctx.D.grad.clone(), m.ctx.S.data.clone(), m.ctx.G.data.clone(),
                m.ctx.D.data.clone())
    a = one_step(serial=True)
    a2 = one_step(serial=True)
    b = one_step(serial=False)
    # losses: the forward passes differ only by the order of the fp32 atomics inside the BN statistics
    assert torch.allclose(a[0], b[0], rtol=2e-3, atol=1e-5), (a[0], b[0])
    # Gradients.  On this random-init net at batch 4 a step is chaotic: the atomics' order alone moves the generator
    # gradient to cos ~0.87 between two runs of the SAME schedule (scripts/overlap_probe.py).  The side-stream schedule must
    # sit inside that run-to-run spread (a race - a gradient read before it is complete, a buffer reused too early -
    # would not); the per-layer pinning of the same schedule against the oracle is tests/test_whole_step_gpu.py.
    for gi, floor in ((1, 0.80), (2, 0.97)):
        spread, diff = _cos(a[gi], a2[gi]), _cos(a[gi], b[gi])
        assert spread > floor, spread
        assert diff > spread - 0.05, (gi, spread, diff)
        assert abs(float(a[gi].norm() / b[gi].norm()) - 1) < 0.1
    # BN moving statistics and the updated parameters (Adam's first step moves every weight by <= lr)
    assert torch.allclose(a[3], b[3], rtol=2e-3, atol=2e-4)
    lr = CFG["training"]["lr"]["start_val"]
    assert float((a[4] - b[4]).abs().max()) <= 2.001 * lr and float((a[5] - b[5]).abs().max()) <= 2.001 * lr
