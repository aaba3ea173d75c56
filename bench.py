#!/usr/bin/env python
"""bench.py — stage-1 hot-path benchmark (contract: task statement; details in DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|k1] [--impl ours|reference]

Prints ONE JSON line (rank 0).
Workloads
  train  BASELINE.json configs[2] — the configuration the metric "stage-1 frames/sec at 1/2/4/8 B200" is quoted on:
         one reference `train_step` (D run on one batch + G run on another, fwd+bwd incl. VGG19 perceptual loss,
         two Adam updates) at batch 32 per GPU, random-init weights incl. VGG19, synthetic frames 128x128,
         40 keypoints, frame-batch data parallel with one NCCL all-reduce per optimiser.  A step consumes
         4*32 frames per GPU (2 batches x (image, future_image)).
  k1     BASELINE.json configs[1] — fused soft-argmax + Gaussian render micro-bench, 1024 frames per GPU.
  pseudo BASELINE.json configs[3] — make_pseudo_labels: KeypointModel (detector only, inference BN) over the rank's
         shard of the frame list, no collective; a step = --pseudo-frames frames per GPU.
  render BASELINE.json configs[4] — evaluate-style rendering: FinalModel turns 64 first frames + 64 synthetic
         32-step keypoint trajectories per GPU into 64 x 32 frames (translator, inference BN).
  fwd8   BASELINE.json configs[0] — stage-1 detector+translator FORWARD on 8 frame pairs (train- and inference-mode BN), with the
         CPU oracle of the same call timed beside it on the host cores.
The default run measures `train` and, at N=1, appends the other configurations as sub-objects `"k1"`, `"fwd8"`, `"pseudo"`,
`"render"` (short step counts, each with its own `roofline` and `e2e`), the per-layer table of the HBM-bound convolutions
(`roofline.hbm_bound_layers`) and the CPU baseline (oracle train step at batch 32 on all host threads).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K1_BYTES_PER_FRAME = 128 * 128 * 40 * 4 + 32 * 32 * 40 * 4 + 40 * 2 * 4   # 2 785 600
TRAIN_GFLOP_PER_EXAMPLE = 152.369                                          # SURVEY.md §8d config 3 (useful conv work)
TRAIN_WORKLOAD = ("BASELINE configs[2]: stage-1 detector_translator train_step (D run + G run on different batches, fwd+bwd incl. "
                  "VGG19 perceptual loss + img_discr, 2x Adam), random-init incl. VGG19")
K1_WORKLOAD = "BASELINE configs[1]: fused soft-argmax + Gaussian render, [128,128,40] fp32 logits -> mu [40,2] + maps [32,32,40]"
CONFIG = {"paths": {"data_dir": "", "vggnet": None, "log_dir": "/tmp/kp_b200_logs"},
          "training": {"batch_size": 32, "lr": {"start_val": 1e-4, "step": 20000, "decay": 0.95}},
          "model": {"n_pts": 40, "n_action": 9, "cell_info": [1024, 1024], "vae_dim": 64}}


def train_config(world, batch=32):
    """Static description of the train workload - identical in the CUDA arm and in the reference arm (the driver compares
    the two `config` objects); everything run-specific lives under the top-level key "run"."""
    return {"workload": TRAIN_WORKLOAD, "batch_per_gpu": batch, "frames_per_step_per_gpu": 4 * batch, "image_hw": [128, 128],
            "n_pts": 40, "parallelism": "dp%d" % world,
            "l2": "per-step working set (activations + 51 M parameters + Adam slots, several GB) >> 126 MB L2; input batches rotate"}


def _host_threads():
    """All host cores for the CPU arms, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    os.environ.pop("OMP_NUM_THREADS", None)
    os.environ.pop("MKL_NUM_THREADS", None)
    torch.set_num_threads(n)
    return n


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (NVML, 20 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop, self._thr = threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join(timeout=2)
        med = int(statistics.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# CPU arms: the oracle (torch-CPU / numpy restatement of the reference) on the host cores
# --------------------------------------------------------------------------------------------------
def _cpu_k1_chunk(args):
    import numpy as np
    from oracle import k1_numpy as o
    seed, n = args
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, 128, 128, 40), dtype=np.float32) * np.float32(5.0))
    t0 = time.perf_counter()
    mu, px, py, maps = o.softargmax_render_fwd(x, [32, 32])
    return time.perf_counter() - t0, float(mu.sum())


def cpu_k1(frames_per_worker, workers, reps):
    """frames/s of the numpy oracle of utils/model.py over `workers` processes."""
    import multiprocessing as mp
    best = None
    if workers <= 1:
        for r in range(reps):
            dt, _ = _cpu_k1_chunk((r, frames_per_worker))
            best = dt if best is None else min(best, dt)
        return frames_per_worker / best
    with mp.get_context("fork").Pool(workers) as pool:
        for r in range(reps):
            res = pool.map(_cpu_k1_chunk, [(r * workers + i, frames_per_worker) for i in range(workers)])
            dt = max(d for d, _ in res)
            best = dt if best is None else min(best, dt)
    return frames_per_worker * workers / best


class CpuTrainer:
    """The reference train_step restated on torch CPU (oracle/networks.py): D run + G run with autograd + TF Adam."""

    def __init__(self, batch, seed=0):
        import numpy as np
        import torch
        from oracle import networks as ON
        self.torch, self.ON, self.B = torch, ON, batch
        self.P = ON.init_params(seed, dtype=torch.float32)
        rng = np.random.default_rng(seed)
        self.batches = [tuple(torch.from_numpy(rng.uniform(-1, 1, (batch, 128, 128, 3)).astype(np.float32)) for _ in range(2))
                        for _ in range(2)]
        self.state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in self.P.items()
                      if not k.startswith("vgg") and "moving" not in k}
        self.t = 0

    def _apply(self, names, lr):
        from oracle import tf_ops as T
        torch = self.torch
        with torch.no_grad():
            for n in names:
                p = self.P[n]
                if p.grad is None:
                    continue
                m, v = self.state[n]
                newp, m, v = T.adam_tf(p.detach(), p.grad, m, v, self.t, lr)
                self.state[n] = (m, v)
                self.P[n] = newp

    def step(self):
        torch, ON = self.torch, self.ON
        self.t += 1
        d_names = [k for k in self.state if "img_discr" in k]
        g_names = [k for k in self.state if "img_discr" not in k]
        # D run
        im, fut = self.batches[0]
        for k in self.P:
            self.P[k] = self.P[k].detach().requires_grad_(k in d_names)
        ctx = ON.Ctx(self.P)
        with torch.no_grad():
            out = ON.forward_pass(ctx, im, fut, 40, True)
        lD = ON.loss_D(ctx, out["final_output"], fut)[0]
        lD.backward()
        self._apply(d_names, 1e-4)
        # G run
        im, fut = self.batches[1]
        for k in self.P:
            self.P[k] = self.P[k].detach().requires_grad_(k in g_names)
        ctx = ON.Ctx(self.P)
        out = ON.forward_pass(ctx, im, fut, 40, True)
        lG = ON.loss_G(ctx, out["final_output"], fut)[0]
        lG.backward()
        self._apply(g_names, 1e-4)
        with torch.no_grad():
            for name, val in ctx.updates:
                self.P[name] = val.detach()
        return float(lD.detach()), float(lG.detach())


def cpu_train(batch, steps, warmup=1, budget_s=None):
    """frames/s (4*batch frames per step) of the torch-CPU oracle train step on all host threads, median over the timed
    steps; stops early when `budget_s` seconds of timed work are spent (at least 2 steps when steps >= 2)."""
    threads = _host_threads()
    tr = CpuTrainer(batch)
    for _ in range(warmup):
        tr.step()
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        tr.step()
        ts.append(time.perf_counter() - t0)
        if budget_s is not None and sum(ts) >= budget_s and len(ts) >= min(2, steps):
            break
    return 4 * batch / statistics.median(ts), threads, statistics.median(ts), len(ts)


def cpu_fwd8(pairs=8, reps=5):
    """BASELINE configs[0] on the host: oracle forward (fp32, train- and inference-mode BN) of `pairs` frame pairs; median of
    `reps` after one warm-up.  Returns {mode: seconds}, threads."""
    import numpy as np
    import torch
    from oracle import networks as ON
    threads = _host_threads()
    P = ON.init_params(0, dtype=torch.float32)
    rng = np.random.default_rng(0)
    im, fut = [torch.from_numpy(rng.uniform(-1, 1, (pairs, 128, 128, 3)).astype(np.float32)) for _ in range(2)]
    out = {}
    with torch.no_grad():
        for mode, train in (("train_bn", True), ("inference_bn", False)):
            ts = []
            for r in range(reps + 1):
                t0 = time.perf_counter()
                ON.forward_pass(ON.Ctx(P), im, fut, 40, train)
                ts.append(time.perf_counter() - t0)
            out[mode] = statistics.median(ts[1:])
    return out, threads


def run_reference(args):
    """--impl reference: the reference's CPU path.  TensorFlow 1.12 cannot be installed here, so this is the oracle
    port of the same step (torch CPU, all host threads) at the SAME configuration (batch 32); the number of timed steps is
    bounded by a wall-clock budget so that the run ends within a few minutes.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.workload == "k1":
        workers = max(1, min(os.cpu_count() or 1, 32))
        vals = [cpu_k1(8, workers, 1) for _ in range(max(args.warmup, 1) + min(args.steps, 10))][max(args.warmup, 1):]
        v, cores = statistics.median(vals), workers
        sample = "%d steps x %d frames (%d procs), numpy fp32 oracle of utils/model.py" % (len(vals), 8 * workers, workers)
        metric, steps, ms = "stage-1 frames/sec (fused soft-argmax + Gaussian render)", len(vals), 1e3 * 8 * workers / v
        config = {"workload": K1_WORKLOAD}
    else:
        b = args.batch
        v, cores, sec, steps = cpu_train(b, max(1, args.steps), warmup=1, budget_s=args.cpu_budget)
        sample = ("%d timed train_steps (D run + G run, fwd+bwd incl. VGG19, TF Adam) at batch %d after 1 warm-up step, %.1f s per "
                  "step: torch-CPU fp32 oracle of the reference graph on %d threads (host has %d cores); timed steps bounded by a "
                  "%d s budget" % (steps, b, sec, cores, os.cpu_count() or 0, args.cpu_budget))
        metric, ms = "stage-1 frames/sec (train_step: D run + G run)", sec * 1e3
        config = train_config(args.gpus, b)
    emit({
        "impl": "reference", "metric": metric, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


# --------------------------------------------------------------------------------------------------
# GPU arms
# --------------------------------------------------------------------------------------------------
_RESULT_FD = None
CONV_KERNELS = ("tapconv_kernel", "tapconv2_kernel", "haloconv_kernel", "halo2_kernel", "wgrad_kernel", "wgrad2_kernel")
RIDGE_FLOP_PER_BYTE = 247.0     # bf16 ridge of B200: below it a layer is HBM-bound


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL prints its version banner (and any NCCL_DEBUG output) on stdout: send it to a file so that stdout
        # stays ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/kp_b200_nccl.%h.%p.log")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
        local_rank = 0
    return world, rank, local_rank


def _barrier(world):
    import torch
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(x, world, dev):
    import torch
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def conv_profile(replay, tags, peaks, n_rep=2, detail_path=None):
    """In-situ durations of the convolution kernels of a CAPTURED step: CUPTI activity records (torch.profiler) of `n_rep`
    replays, zipped in launch order with the (kind, algorithmic FLOPs, algorithmic bytes, tag) entries that conv.TAGS
    collected during the capture.  CUDA events cannot bracket a kernel inside a graph replay, and per-launch brackets of an
    eager step include the host-side launch cost.  Returns the summary that goes into `roofline`."""
    import torch
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_rep):
            replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    all_us = sum(e.time_range.end - e.time_range.start for e in evs) / n_rep
    conv = [e for e in evs if any(k in e.name for k in CONV_KERNELS)]
    fam = {}
    for e in conv:
        k = next(k for k in CONV_KERNELS if k in e.name)
        a = fam.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    out = {"all_kernels_ms": all_us * 1e-3, "kernels_per_step": len(evs) // n_rep,
           "families": {k: {"launches": v[0] // n_rep, "ms": v[1] / n_rep * 1e-3} for k, v in fam.items()}}
    conv_us = sum(v[1] for v in fam.values()) / n_rep
    flops = sum(t[1] for t in tags)
    out.update({"conv_launches": len(tags), "conv_flops": flops, "conv_ms": conv_us * 1e-3,
                "conv_tflops": flops / (conv_us * 1e-6) / 1e12 if conv_us > 0 else None})
    if len(conv) != n_rep * len(tags):
        out["layer_table"] = "unavailable: %d conv kernels in %d replays vs %d tagged launches" % (len(conv), n_rep, len(tags))
        return out
    # per-layer aggregation (same signature = same row), averaged over the replays
    rows = {}
    for i, e in enumerate(conv):
        kind, fl, by, tag = tags[i % len(tags)]
        r = rows.setdefault((kind, tag), {"kind": kind, "layer": tag, "launches": 0, "us": 0.0, "flops": 0.0, "bytes": 0.0})
        r["launches"] += 1
        r["us"] += e.time_range.end - e.time_range.start
        r["flops"] += fl
        r["bytes"] += by
    table = []
    for r in rows.values():
        for k in ("launches", "us", "flops", "bytes"):
            r[k] = r[k] / n_rep
        r["launches"] = int(round(r["launches"]))
        r["tflops"] = r["flops"] / (r["us"] * 1e-6) / 1e12
        r["gbs"] = r["bytes"] / (r["us"] * 1e-6) / 1e9
        r["flop_per_byte"] = r["flops"] / max(r["bytes"], 1.0)
        table.append(r)
    table.sort(key=lambda r: -r["us"])
    hbm = [r for r in table if r["flop_per_byte"] < RIDGE_FLOP_PER_BYTE]
    ten = [r for r in table if r["flop_per_byte"] >= RIDGE_FLOP_PER_BYTE]

    def agg(rs):
        us, fl, by = sum(r["us"] for r in rs), sum(r["flops"] for r in rs), sum(r["bytes"] for r in rs)
        return {"launches": sum(r["launches"] for r in rs), "ms": us * 1e-3, "tflops": fl / (us * 1e-6) / 1e12 if us else None,
                "gbs": by / (us * 1e-6) / 1e9 if us else None}
    a_h, a_t = agg(hbm), agg(ten)
    out["tensor_bound_layers"] = dict(a_t, frac_of_tensor_peak=(a_t["tflops"] or 0) / peaks["bf16_tflops_sustained"],
                                      note="conv launches with algorithmic intensity >= %.0f FLOP/B" % RIDGE_FLOP_PER_BYTE)
    out["hbm_bound_layers"] = dict(a_h, frac_of_hbm_peak=(a_h["gbs"] or 0) / peaks["hbm_gbs"],
                                   note="conv launches with algorithmic intensity < %.0f FLOP/B (algorithmic bytes = inputs + "
                                        "outputs + packed weights, each once)" % RIDGE_FLOP_PER_BYTE,
                                   top=[{"kind": r["kind"], "layer": r["layer"], "launches": r["launches"], "us": round(r["us"], 1),
                                         "MB": round(r["bytes"] / 1e6, 1), "gbs": round(r["gbs"], 1),
                                         "frac_of_hbm_peak": round(r["gbs"] / peaks["hbm_gbs"], 3),
                                         "tflops": round(r["tflops"], 1)} for r in hbm[:12]])
    if detail_path:
        try:
            os.makedirs(os.path.dirname(detail_path), exist_ok=True)
            with open(detail_path, "w") as fh:
                json.dump({"layers": table, "families": out["families"]}, fh, indent=1)
            out["layer_table"] = os.path.relpath(detail_path, ROOT)
        except OSError:
            pass
    return out


def measure_inference(kind, args, world, rank, dev, lib, peaks, steps, warmup):
    """pseudo / render / fwd8 workloads: sharded over ranks with no collective.  The step is captured
    into a CUDA graph (as the train step is); `value` = K replays between CUDA events, `e2e` = pinned host inputs -> H2D ->
    step -> results D2H every step, conv roofline from CUPTI durations inside the replayed graph."""
    import torch
    from kp_b200 import models, conv as cv
    cfg = json.loads(json.dumps(CONFIG))
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    extra = {}
    if kind == "pseudo":
        F = args.pseudo_frames
        model = models.KeypointModel(cfg, device=dev)
        static_in = [torch.rand((F, 128, 128, 3), device=dev, generator=gen) * 2 - 1]
        pool = [torch.rand((F, 128, 128, 3), device=dev, generator=gen) * 2 - 1 for _ in range(2)]   # 2 x 805 MB at F=4096
        step = lambda: model.detect(static_in[0])
        units, h2d, d2h = F, F * 128 * 128 * 3 * 4, F * 40 * 2 * 4
        metric = "stage-1 frames/sec (make_pseudo_labels: detector-only pass)"
        wl = ("BASELINE configs[3]: KeypointModel.detect over %d synthetic frames per GPU per step (the rank's shard of the frame "
              "list, no collective), inference-mode BN folded into the convolutions, random-init" % F)
        flop_unit = 3.648e9
    elif kind == "render":
        V, T = args.render_videos, 32
        model = models.FinalModel(cfg, device=dev)
        static_in = [torch.rand((V, 128, 128, 3), device=dev, generator=gen) * 2 - 1,
                     torch.rand((V, T, 40, 2), device=dev, generator=gen) * 1.6 - 0.8]
        pool = [(torch.rand((V, 128, 128, 3), device=dev, generator=gen) * 2 - 1,
                 torch.rand((V, T, 40, 2), device=dev, generator=gen) * 1.6 - 0.8) for _ in range(2)]

        def step():
            model.build({"image": static_in[0], "pred_seq": static_in[1]})
            return model.run(visualize=False)["pred_im_seq"]
        units, h2d, d2h = V * T, V * 128 * 128 * 3 * 4 + V * T * 40 * 2 * 4, V * T * 128 * 128 * 3 * 4
        metric = "stage-1 frames/sec (evaluate-style rendering: translator over keypoint trajectories)"
        wl = ("BASELINE configs[4]: FinalModel.run on %d videos per GPU per step: image_encoder + pose_encoder on the first frame, "
              "%d-step synthetic trajectories -> Gaussian maps -> translator -> mask compose; %d frames per step, random-init"
              % (V, T, V * T))
        flop_unit = 14.345e9
    else:   # fwd8
        Bp = args.fwd_pairs
        cfg["training"]["batch_size"] = Bp
        model = models.DetectorTranslatorModel(cfg, is_training=True, device=dev, seed=0)
        static_in = [torch.rand((Bp, 128, 128, 3), device=dev, generator=gen) * 2 - 1 for _ in range(2)]
        pool = [tuple(torch.rand((Bp, 128, 128, 3), device=dev, generator=gen) * 2 - 1 for _ in range(2)) for _ in range(2)]
        model.build({"image": static_in[0], "future_image": static_in[1]})

        def step():
            model.ctx.begin_run()
            model.ctx.tape, model.ctx.update_moving = None, False
            return model._define_forward_pass(static_in[0], static_in[1], for_G_run=True)
        units, h2d, d2h = 2 * Bp, 2 * Bp * 128 * 128 * 3 * 4, Bp * 128 * 128 * 3 * 4
        metric = "stage-1 frames/sec (detector + translator forward, %d frame pairs)" % Bp
        wl = ("BASELINE configs[0]: DetectorTranslatorModel._define_forward_pass on %d synthetic frame pairs (image_encoder, "
              "pose_encoder x2, Gaussian maps, translator, mask compose), batch-statistics BN as train.py runs it, random-init" % Bp)
        flop_unit = 23.0026e9 / 2
    # eager warm-up (plan / packed-weight caches), then capture
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    cv.TAGS = []
    n0 = lib.kp_launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        result = step()
    launches_per_step = int(lib.kp_launch_count() - n0)
    tags, cv.TAGS = cv.TAGS, None

    def set_inputs(src):
        src = src if isinstance(src, (tuple, list)) else (src,)
        for dst, t in zip(static_in, src):
            dst.copy_(t, non_blocking=True)
    for i in range(warmup):
        set_inputs(pool[i % 2])
        graph.replay()
    _barrier(world)
    sampler = ClockSampler(dev.index or 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    ev0.record()
    for i in range(steps):
        set_inputs(pool[i % 2])          # device-resident inputs rotate (L2 >> is exceeded by every set)
        graph.replay()
    ev1.record()
    _barrier(world)
    clocks = sampler.stop()
    ms = _max_over_ranks(ev0.elapsed_time(ev1), world, dev) / steps
    value = world * units / (ms * 1e-3)
    # end to end through host buffers
    host = [tuple(t.cpu().pin_memory() for t in (pp if isinstance(pp, (tuple, list)) else (pp,))) for pp in pool]
    out_host = torch.empty(tuple(result.shape), dtype=result.dtype).pin_memory()
    e2e_steps = max(3, min(steps, 10))
    set_inputs(host[0]); graph.replay(); torch.cuda.synchronize()
    _barrier(world)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        set_inputs(host[i % 2])
        graph.replay()
        out_host.copy_(result, non_blocking=True)
        torch.cuda.synchronize()
    _barrier(world)
    e2e_value = world * units * e2e_steps / _max_over_ranks(time.perf_counter() - t0, world, dev)
    prof = conv_profile(graph.replay, tags, peaks, n_rep=2,
                        detail_path=os.path.join(ROOT, "gpurun_out", "layers_%s.json" % kind) if rank == 0 else None)
    peak = peaks["bf16_tflops_sustained"]
    achieved = prof.get("conv_tflops") or 0.0
    line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": wl, "frames_per_step_per_gpu": units, "parallelism": "dp%d (sharded, no collective)" % world,
                       "cuda_graph": True, "l2": "inputs of a step >> 126 MB L2 (fwd8: 3 MB inputs, activations 0.5 GB); two input sets alternate"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peaks["source"] + " (sustained)",
                         "kernel": "all %d convolution launches of one step (kp::halo2_kernel / kp::tapconv_kernel): algorithmic "
                                   "FLOPs / summed in-situ CUPTI durations of the replayed graph" % len(tags),
                         "algorithmic_flops_per_step": prof.get("conv_flops"), "conv_ms": prof.get("conv_ms"),
                         "whole_step_tflops": flop_unit * units / (ms * 1e-3) / 1e12,
                         "whole_step_frac": flop_unit * units / (ms * 1e-3) / 1e12 / peak,
                         "tensor_bound_layers": prof.get("tensor_bound_layers"), "hbm_bound_layers": prof.get("hbm_bound_layers"),
                         "families": prof.get("families"), "all_kernels_ms": prof.get("all_kernels_ms")},
            "cpu_baseline": None,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "note": "pinned host inputs -> H2D -> captured step -> results D2H, every step"},
            "gpu_launches": launches_per_step * steps, "gpu_launches_per_step": launches_per_step, "clocks": clocks}
    if kind == "fwd8" and rank == 0 and world == 1:
        try:
            secs, threads = cpu_fwd8(args.fwd_pairs)
            line["cpu_baseline"] = {"value": units / secs["train_bn"], "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": "oracle forward (fp32, torch CPU) of the same %d pairs: median of 5 after 1 warm-up: "
                                              "%.3f s train-mode BN, %.3f s inference-mode BN; %d threads, host has %d cores"
                                              % (args.fwd_pairs, secs["train_bn"], secs["inference_bn"], threads, os.cpu_count() or 0)}
        except Exception as e:   # reporting only
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    del graph
    return line


def bench_k1(args, world, rank, dev, lib):
    import torch
    from kp_b200 import k1
    B = args.frames
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    logits = torch.randn((B, 128, 128, 40), device=dev, generator=gen) * 5.0   # 2.68 GB >> 126 MB L2
    for _ in range(max(args.warmup, 3)):
        k1.softargmax_render_fwd(logits, (32, 32), want_prob=False)
    _barrier(world)
    steps = args.k1_steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.kp_launch_count()
    ev0.record()
    for _ in range(steps):
        k1.softargmax_render_fwd(logits, (32, 32), want_prob=False)
    ev1.record()
    _barrier(world)
    ms = _max_over_ranks(ev0.elapsed_time(ev1), world, dev) / steps
    # end to end: pinned host logits -> H2D -> kernel -> mu + maps D2H
    e2e = None
    if world == 1:
        hb = min(B, 256)
        host = torch.randn((hb, 128, 128, 40)).pin_memory()
        mu_h, maps_h = torch.empty((hb, 40, 2)).pin_memory(), torch.empty((hb, 32, 32, 40)).pin_memory()
        t0 = None
        for i in range(4):
            if i == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            mu, _, _, maps = k1.softargmax_render_fwd(host.to(dev, non_blocking=True), (32, 32), want_prob=False)
            mu_h.copy_(mu, non_blocking=True); maps_h.copy_(maps, non_blocking=True)
            torch.cuda.synchronize()
        e2e = {"value": hb * 3 / (time.perf_counter() - t0), "unit": "frames/s", "h2d_bytes_per_step": hb * 128 * 128 * 40 * 4,
               "d2h_bytes_per_step": hb * (40 * 2 + 32 * 32 * 40) * 4, "steps": 3,
               "note": "%d frames per step from pinned host memory; bound by the host link, not by the kernel" % hb}
    del logits
    peaks = _peaks()
    achieved = K1_BYTES_PER_FRAME * B / (ms * 1e-3) / 1e9
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
        tsrc = "profiles/k1_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch)"
    return {"workload": K1_WORKLOAD, "frames_per_gpu": B, "l2": "input 2.68 GB per launch >> 126 MB L2",
            "frames_per_s": world * B / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
            "gpu_launches": int(lib.kp_launch_count() - n0), "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": tsrc,
                         "peak_source": peaks["source"],
                         "kernel": "kp::k1_fwd_fast<5,8>", "algorithmic_bytes_per_launch": K1_BYTES_PER_FRAME * B}}


INPUT_BYTES_PER_FRAME = 128 * 128 * 3 * (1 + 4)     # bytes gathered (one per output byte) + float32 output written


def _synthetic_jpeg_dataset(root, videos=8, frames=24, size=(320, 240), seed=0):
    """<root>/frames/%04d/%06d.jpg + train_set.txt, the layout data/image_pair_dataloader.py reads."""
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    w, h = size
    yy, xx = np.mgrid[0:h, 0:w]
    names = []
    for v in range(videos):
        d = os.path.join(root, "frames", "%04d" % (v + 1))
        os.makedirs(d, exist_ok=True)
        names.append("frames/%04d %d" % (v + 1, v % 3))
        a, ph = rng.uniform(0.02, 0.09, (3, 2)), rng.uniform(0, 6.28, 3)
        for t in range(frames):
            img = np.stack([127 + 110 * np.sin(a[c, 0] * xx + ph[c] + 0.1 * t) * np.cos(a[c, 1] * yy - ph[c]) for c in range(3)], -1)
            img = np.clip(img + rng.normal(0, 10, img.shape), 0, 255).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(d, "%06d.jpg" % (t + 1)), quality=85)
    with open(os.path.join(root, "train_set.txt"), "w") as fh:
        fh.write("\n".join(names))


def _pillow_pair(root, names, rnd):
    """cpu_baseline leg only: one training pair through the reference's Pillow call sequence (data/image_pair_dataloader.py:
    72-165, utils/data.py:8-35; resize resample NEAREST = the pinned Pillow 6.2.0 default), on the installed Pillow."""
    import numpy as np
    from PIL import Image, ImageEnhance, ImageFilter
    folder = os.path.join(root, names[rnd.randrange(len(names))].split()[0])
    n = len(os.listdir(folder))
    step, i0 = rnd.randint(8, 11), rnd.randint(0, n - 1)
    ims = [Image.open(os.path.join(folder, "%06d.jpg" % (i + 1))) for i in (i0, (i0 + step) % n)]
    w, h = ims[0].size
    ang = rnd.randrange(-10, 11)
    ims = [im.rotate(ang) for im in ims]
    ratio = min(w, h) / 128.0
    ims = [im.resize([int(w / ratio), int(h / ratio)], Image.NEAREST) for im in ims]
    c = rnd.randint(0, int(max(w, h) / ratio - 128))
    box = (c, 0, c + 128, 128) if w > h else (0, c, 128, c + 128)
    ims = [im.crop(box) for im in ims]
    if rnd.randint(0, 1):
        ims = [im.transpose(Image.FLIP_LEFT_RIGHT) for im in ims]
    r = rnd.randint(0, 9)
    F = [ImageFilter.DETAIL, ImageFilter.EDGE_ENHANCE, ImageFilter.SMOOTH, ImageFilter.SMOOTH_MORE,
         ImageFilter.EDGE_ENHANCE_MORE, ImageFilter.BLUR]
    if r < 6:
        ims = [im.filter(F[r]) for im in ims]
    else:
        E, (lo, hi) = [(ImageEnhance.Sharpness, (0, 50)), (ImageEnhance.Brightness, (7, 20)), (ImageEnhance.Color, (0, 50)),
                       (ImageEnhance.Contrast, (7, 30))][r - 6]
        v = rnd.randint(lo, hi) * 0.1
        ims = [E(im).enhance(v) for im in ims]
    return [(np.asarray(im) / 255.0).astype(np.float32) * 2.0 - 1.0 for im in ims]


def _train_from_files(args, dev, root):
    """The whole chain a user of train.py runs: JPEG files -> ImagePairDataLoader -> DetectorTranslatorModel.train_step (captured
    graph), losses read back every step."""
    import random
    import numpy as np
    import torch
    from kp_b200 import data, models
    B = args.batch
    cfg = json.loads(json.dumps(CONFIG))
    cfg["training"]["batch_size"] = B
    np.random.seed(1); random.seed(1)
    ld = data.ImagePairDataLoader(root, "train", random_order=True, randomness=True)
    workers = int(os.environ.get("KP_INPUT_WORKERS", str(_host_threads())))   # capped by the loader at cores - 2
    # the loader's batch-building thread and this thread share the GIL: a short switch interval keeps the step's launches
    # from waiting the default 5 ms behind it
    old_interval = sys.getswitchinterval()
    sys.setswitchinterval(float(os.environ.get("KP_SWITCH_INTERVAL", "0.0005")))
    ds = ld.get_dataset(batch_size=B, repeat=True, shuffle=True, num_preprocess_threads=workers, prefetch=True, device=dev)
    it = iter(ds)
    model = models.DetectorTranslatorModel(cfg, is_training=True, device=dev, seed=0)
    model.build(lambda: next(it))
    model.enable_cuda_graph(B)
    for _ in range(8):
        model.train_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.input_train_steps):
        model.train_step(should_write_log=True)       # reads the losses back
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    ds.close()
    model._graph = None
    sys.setswitchinterval(old_interval)
    return {"value": 4 * B * args.input_train_steps / sec, "unit": "frames/s", "ms_per_step": sec / args.input_train_steps * 1e3,
            "steps": args.input_train_steps,
            "note": "train_step at batch %d fed by ImagePairDataLoader.get_dataset (JPEG decode in %d worker processes, augmentation "
                    "kernel, shuffle buffer) - two batches per step, losses D2H every step" % (B, ds.n_workers)}


def _with_deadline(seconds, fn, *a):
    """Run a sub-benchmark that talks to worker processes under an alarm, so that a stuck pipe cannot take the headline line
    with it (main thread only; blocking queue / pipe reads are interrupted by the signal)."""
    import signal

    def on_alarm(signum, frame):
        raise TimeoutError("sub-benchmark exceeded %d s" % seconds)

    old = signal.signal(signal.SIGALRM, on_alarm)
    signal.alarm(seconds)
    try:
        return fn(*a)
    finally:
        signal.alarm(0)
        signal.signal(signal.SIGALRM, old)


def bench_input(args, dev, lib, peaks):
    """Input pipeline: kp_augment_frames on frames resident in HBM (roofline: HBM), the loader end to
    end from JPEG files (PIL decode on host threads -> pinned staging -> H2D -> one launch per batch), and the reference's
    Pillow chain on one host thread beside it."""
    import random
    import tempfile
    import numpy as np
    import torch
    from kp_b200 import augment as A
    from kp_b200 import data
    n, w, h = args.input_frames, 320, 240                       # 2048 x 230 400 B = 472 MB of decoded frames >> 126 MB L2
    gen = torch.Generator(device=dev).manual_seed(5)
    src = torch.randint(0, 256, (n * w * h * 3,), device=dev, dtype=torch.uint8, generator=gen)
    rnd = random.Random(5)
    table = A.PlanTable(n)
    for i in range(n):
        fid = rnd.randint(0, 9)
        fac = {6: rnd.randint(0, 50), 7: rnd.randint(7, 20), 8: rnd.randint(0, 50), 9: rnd.randint(7, 30)}.get(fid, 0) * 0.1
        table.set(i, i * w * h * 3, w, h, 170, 128, rnd.randint(0, 42), 0, rnd.randrange(-10, 11), rnd.randint(0, 1), fid, fac)
    plans = table.host.to(dev)
    out = torch.empty((n, 128, 128, 3), device=dev)
    for _ in range(3):
        A.augment_frames(src, plans, n, out=out)
    steps = 20
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.kp_launch_count()
    ev0.record()
    for _ in range(steps):
        A.augment_frames(src, plans, n, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    launches = int(lib.kp_launch_count() - n0)
    # the keypoint loader's frames (data/keypoint_dataloader.py:71): resize + centre crop only
    for i in range(n):
        table.set(i, i * w * h * 3, w, h, 170, 128, 21, 0)
    plans.copy_(table.host)
    for _ in range(3):
        A.augment_frames(src, plans, n, out=out)
    ev0.record()
    for _ in range(steps):
        A.augment_frames(src, plans, n, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms_plain = ev0.elapsed_time(ev1) / steps
    del src, out
    achieved = INPUT_BYTES_PER_FRAME * n / (ms * 1e-3) / 1e9
    traffic, tsrc = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "input_traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("frames_per_launch") == n:      # the capture is of this launch size
            traffic, tsrc = tj["dram_bytes_per_launch"], tj["source"]
    except (OSError, KeyError, ValueError):
        pass
    sub = {"workload": "input pipeline of the stage-1 loaders (data/image_pair_dataloader.py:72-165): rotate + resize + crop + flip "
                       "+ random filter + normalise of %d decoded 320x240 frames per launch, random plans as the reference draws them" % n,
           "frames_per_s": n / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "gpu_launches": launches, "dtype": "u8",
           "l2": "decoded frames %d MB + output %d MB per launch >> 126 MB L2" % (n * w * h * 3 >> 20, n * 196608 >> 20),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                        "traffic": traffic, "traffic_source": tsrc, "peak_source": peaks["source"], "kernel": "kp::augment_kernel",
                        "algorithmic_bytes_per_launch": INPUT_BYTES_PER_FRAME * n},
           "resize_crop_only": {"note": "the keypoint loader's plan (no rotation, no filter), same frames", "ms_per_step": ms_plain,
                                "frames_per_s": n / (ms_plain * 1e-3), "gbs": INPUT_BYTES_PER_FRAME * n / (ms_plain * 1e-3) / 1e9,
                                "frac_of_hbm_peak": INPUT_BYTES_PER_FRAME * n / (ms_plain * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
    with tempfile.TemporaryDirectory() as root:
        _synthetic_jpeg_dataset(root)
        np.random.seed(0); random.seed(0)
        ld = data.ImagePairDataLoader(root, "train", random_order=True, randomness=True)
        ds = ld.get_dataset(batch_size=32, repeat=True, num_preprocess_threads=_host_threads(), prefetch=True, device=dev)
        t0, k, acc = None, 0, torch.zeros((), device=dev)
        for batch in ds:
            acc += batch["image"][0, 0, 0, 0] + batch["future_image"][0, 0, 0, 0]    # consume on the current stream
            k += 1
            if k == 5:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            if k == 5 + args.input_batches:
                break
        float(acc)      # D2H of a value that depends on every batch
        sec = time.perf_counter() - t0
        ds.close()
        jpeg = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(root) for f in fs if f.endswith(".jpg"))
        sub["e2e"] = {"value": 64 * args.input_batches / sec, "unit": "frames/s", "h2d_bytes_per_step": 64 * (w * h * 3 + A.PLAN_BYTES),
                      "d2h_bytes_per_step": 4, "steps": args.input_batches,
                      "note": "ImagePairDataLoader.get_dataset(32): JPEG files (%d KB each) -> PIL decode in %d worker processes -> shared pinned "
                              "staging -> H2D -> one kp_augment_frames launch per batch of 64 frames; bound by the host JPEG decode"
                              % (jpeg // (8 * 24) >> 10, ds.n_workers)}
        if args.input_train_steps > 0:
            sub["train_from_files"] = _train_from_files(args, dev, root)
        names = open(os.path.join(root, "train_set.txt")).read().splitlines()
        rnd = random.Random(0)
        t0, pairs = time.perf_counter(), 0
        while time.perf_counter() - t0 < 3.0:
            _pillow_pair(root, names, rnd)
            pairs += 1
        sub["cpu_baseline"] = {"value": 2 * pairs / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": "%d pairs in 3 s through the reference's Pillow call sequence (decode included) on one host "
                                         "thread, as its generator runs (tf.data parallelises only map_fn)" % pairs}
    return sub


def run_ours(args):
    import torch
    import __graft_entry__ as g
    world, rank, local_rank = _dist_setup()
    if rank == 0:
        g.build()
    _barrier(world)
    import kp_b200
    from kp_b200 import models, conv as cv
    lib = kp_b200._lib.load()
    dev = torch.device("cuda", local_rank)
    peaks = _peaks()

    if args.workload == "k1":
        r = bench_k1(args, world, rank, dev, lib)
        if rank == 0:
            emit({"metric": "stage-1 frames/sec (fused soft-argmax + Gaussian render)", "value": r["frames_per_s"],
                              "unit": "frames/s", "n_gpus": world, "steps": r["steps"], "warmup": args.warmup,
                              "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                              "dtype": "f32", "data": "synthetic", "config": {"workload": r["workload"], "frames_per_gpu": r["frames_per_gpu"], "l2": r["l2"]},
                              "roofline": r["roofline"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"]})
        if world > 1:
            torch.distributed.barrier()
            os._exit(0)
        return

    if args.workload == "input":
        if rank == 0:
            sub = bench_input(args, dev, lib, peaks)
            sub.update({"metric": "input-pipeline frames/sec (kp_augment_frames)", "value": sub["frames_per_s"], "unit": "frames/s",
                        "n_gpus": 1, "higher_is_better": True, "data": "synthetic", "config": {"workload": sub["workload"]}})
            emit(sub)
        return
    if args.workload in ("pseudo", "render", "fwd8"):
        line = measure_inference(args.workload, args, world, rank, dev, lib, peaks, args.steps, args.warmup)
        if rank == 0:
            emit(line)
        if world > 1:
            torch.distributed.barrier()
            os._exit(0)
        return

    # ------------------------------ train workload ------------------------------
    B = args.batch
    cfg = json.loads(json.dumps(CONFIG))
    cfg["training"]["batch_size"] = B
    model = models.DetectorTranslatorModel(cfg, is_training=True, device=dev, seed=0)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    pool = [{"image": torch.rand((B, 128, 128, 3), device=dev, generator=gen) * 2 - 1,
             "future_image": torch.rand((B, 128, 128, 3), device=dev, generator=gen) * 2 - 1} for _ in range(6)]
    cursor = {"i": 0}

    def feed():
        cursor["i"] += 1
        return pool[cursor["i"] % len(pool)]
    model.build(feed)
    launches_per_step = None
    tags = None
    if not args.no_graph:
        cv.TAGS = []
        model.enable_cuda_graph(B)      # two eager warm-up steps + the capture: the last third of the tags is the captured step
        tags, cv.TAGS = cv.TAGS, None
        tags = tags[len(tags) - len(tags) // 3:]
        launches_per_step = model.graph_launches                         # library kernels recorded in the captured step
    for _ in range(args.warmup):
        model.train_step()
    _barrier(world)
    sampler = ClockSampler(dev.index or 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.kp_launch_count()
    sampler.start()
    _barrier(world)
    ev0.record()
    for _ in range(args.steps):
        model.train_step()
    ev1.record()
    _barrier(world)
    clocks = sampler.stop()
    if launches_per_step is None:
        launches_per_step = int(lib.kp_launch_count() - n0) // args.steps
    ms_per_step = _max_over_ranks(ev0.elapsed_time(ev1), world, dev) / args.steps
    frames_per_step = 4 * B                                              # 2 batches x (image, future_image)
    value = world * frames_per_step / (ms_per_step * 1e-3)
    lD, lG = model._last_losses
    losses = [float(lD.sum().item()), float(lG.sum().item())]

    # ---- end to end: pinned host frames -> H2D -> train_step -> loss D2H, every step ----
    host = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in pool[:4]]
    hcur = {"i": 0}

    def feed_host():
        hcur["i"] += 1
        return host[hcur["i"] % len(host)]
    # public input-pipeline helper: the H2D copies of the next two batches run on a side stream under the current step
    from kp_b200.utils import DevicePrefetcher
    model.build(DevicePrefetcher(feed_host, dev, depth=2))
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        model.train_step()
    _barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.train_step()
        lD, lG = model._last_losses
        _ = torch.cat([lD, lG]).cpu()                                    # the step's result crosses back every step
    _barrier(world)
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world, dev)
    e2e_value = world * frames_per_step * e2e_steps / e2e_s
    model.build(feed)

    # ---- convolution kernels inside the captured step (CUPTI), every rank replays (the graph holds the all-reduces) ----
    kern = None
    if tags is not None and not args.no_kernel_profile:
        try:
            # The timed step overlaps independent kernels on side streams (weight gradients, the image_encoder / D(fake)
            # branches, the discriminator update): kernels that share the SMs stretch each other, so a kernel's in-situ
            # duration is no longer its own.  The per-kernel numbers come from a re-capture of the SAME step with the side
            # streams switched off (identical kernels and launch order, serialised).
            model.ctx.wgrad_stream = None
            model.ctx.branch_stream = None
            model.overlap_d_update = False
            model.overlap_g_allreduce = False
            cv.TAGS = []
            model.enable_cuda_graph(B)
            tags, cv.TAGS = cv.TAGS, None
            tags = tags[len(tags) - len(tags) // 3:]
            for _ in range(2):
                model.train_step()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(5):
                model.train_step()
            s1.record()
            torch.cuda.synchronize()
            kern = conv_profile(model.train_step, tags, peaks, n_rep=2,
                                detail_path=os.path.join(ROOT, "gpurun_out", "layers_train.json") if rank == 0 else None)
            kern["mode"] = ("serialised re-capture of the timed step (side streams off: same kernels, same order), so that each "
                            "kernel's duration is its own; the timed step overlaps them")
            kern["serialised_ms_per_step"] = s0.elapsed_time(s1) / 5
        except Exception as e:   # reporting only
            kern = {"error": repr(e)}

    subs = {}
    if world == 1 and not args.no_subs:
        # the other BASELINE configurations as sub-objects (short step counts; each frees its model before the next)
        model_graph, model._graph = model._graph, None
        del model_graph
        model = None
        torch.cuda.empty_cache()
        if not args.no_k1:
            subs["k1"] = bench_k1(args, world, rank, dev, lib)
        try:
            subs["input"] = _with_deadline(180, bench_input, args, dev, lib, peaks)
        except Exception as e:
            subs["input"] = {"error": repr(e)}
        torch.cuda.empty_cache()
        for kind, st in (("fwd8", 20), ("pseudo", 5), ("render", 5)):
            try:
                subs[kind] = measure_inference(kind, args, world, rank, dev, lib, peaks, st, 3)
            except Exception as e:   # a failing sub-benchmark must not take the headline line with it
                subs[kind] = {"error": repr(e)}
            torch.cuda.empty_cache()
    elif not args.no_k1:
        subs["k1"] = bench_k1(args, world, rank, dev, lib)

    if rank == 0:
        step_tflops = TRAIN_GFLOP_PER_EXAMPLE * B / (ms_per_step * 1e-3) / 1e3
        peak = peaks["bf16_tflops_sustained"]
        achieved = (kern or {}).get("conv_tflops") or step_tflops
        # DRAM bytes of all conv launches of one step, from the committed ncu launch list of the eager step
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "train_traffic.json")) as fh:
                tj = json.load(fh)
            traffic, traffic_src = tj["conv_dram_bytes_per_step"], tj["source"] + " - sum over the conv launches of one step"
        except (OSError, KeyError, ValueError):
            pass
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                v, cores, sec, n = cpu_train(args.cpu_batch, 1, warmup=0)
                cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                       "sample": "1 train_step at batch %d (%.1f s, the same configuration): torch-CPU fp32 oracle of the reference "
                                 "graph, %d threads, host has %d cores" % (args.cpu_batch, sec, cores, os.cpu_count() or 0)}
            except Exception as e:   # reporting only
                cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        line = {
            "metric": "stage-1 frames/sec (train_step: D run + G run)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": train_config(world, B),
            "run": {"examples_per_s": world * B / (ms_per_step * 1e-3), "cuda_graph": not args.no_graph, "losses_last_step": losses},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_step": sum(t[2] for t in tags) if tags else None,
                         "peak_source": peaks["source"] + " (sustained: kernels timed inside a long step)",
                         "kernel": "kp::halo2_kernel / kp::tapconv_kernel / kp::wgrad_kernel: algorithmic FLOPs of all conv launches of "
                                   "one step / their summed durations inside the step (CUPTI activity records of replays of the "
                                   "captured step with the side streams off - see kernels.mode -, zipped with the per-launch tags "
                                   "recorded at capture)",
                         "whole_step_tflops": step_tflops, "whole_step_frac": step_tflops / peak,
                         "algorithmic_flops_per_step": TRAIN_GFLOP_PER_EXAMPLE * B * 1e9, "kernels": kern},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": 4 * B * 128 * 128 * 3 * 4,
                    "d2h_bytes_per_step": 16, "steps": e2e_steps,
                    "note": "pinned host frames -> H2D (kp_b200.utils.DevicePrefetcher: copies of the next two batches overlap the "
                            "running step) -> DetectorTranslatorModel.train_step -> losses D2H, every step"},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
        }
        line.update(subs)
        emit(line)
    if world > 1:
        # Tear down in a safe order: drop the captured graph (it holds NCCL kernels) before the communicator goes,
        # and leave through os._exit so that no destructor can block on a peer that is already gone.
        torch.distributed.barrier()
        torch.cuda.synchronize()
        if model is not None:
            model._graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="train", choices=["train", "k1", "pseudo", "render", "fwd8", "input"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="train: pairs per GPU per run")
    ap.add_argument("--frames", type=int, default=1024, help="k1: frames per GPU per launch")
    ap.add_argument("--k1-steps", type=int, default=100)
    ap.add_argument("--pseudo-frames", type=int, default=4096, help="pseudo: frames per GPU per step")
    ap.add_argument("--render-videos", type=int, default=64, help="render: videos per GPU per step (32 frames each)")
    ap.add_argument("--fwd-pairs", type=int, default=8, help="fwd8: frame pairs per call")
    ap.add_argument("--input-frames", type=int, default=2048, help="input: decoded frames per launch")
    ap.add_argument("--input-batches", type=int, default=20, help="input: loader batches (32 pairs) timed end to end")
    ap.add_argument("--input-train-steps", type=int, default=40, help="input: train steps fed from JPEG files (0 = skip)")
    ap.add_argument("--cpu-batch", type=int, default=32, help="cpu_baseline of the CUDA arm: oracle train step at this batch")
    ap.add_argument("--cpu-budget", type=int, default=100, help="--impl reference: seconds of timed CPU steps")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-k1", action="store_true")
    ap.add_argument("--no-subs", action="store_true", help="skip the fwd8 / pseudo / render sub-objects")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline of the CUDA arm")
    ap.add_argument("--no-kernel-profile", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL's version banner, nvcc, pytest plugins)
    # are sent to stderr for the whole run, the result goes to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
