"""Torch-facing wrappers of the tap-GEMM convolution primitive (kp_tapconv_bf16).

PyTorch is plumbing: it owns the buffers and the stream and does the (pure data-movement) weight
re-layout; all arithmetic of the convolution runs in csrc/conv_tc.cu on the tcgen05 tensor cores.
"""
import ctypes

import torch

from . import _lib
from . import tapconv as tc
from .tapconv import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SIGMOID_LAST  # noqa: F401


def _stream():
    return torch.cuda.current_stream().cuda_stream


# bench.py sets PROFILE = [] to time every conv launch with CUDA events: entries (kind, algorithmic_flops, ev0, ev1).
# Algorithmic = real channels only: plans of layers whose sources carry zero-padded channels have flop_scale < 1.
PROFILE = None
# bench.py sets TAGS = [] to record (kind, algorithmic_flops, tag) of every conv launch WITHOUT events (legal inside a CUDA-graph
# capture): the n-th entry describes the n-th convolution kernel of the captured step, whose in-situ duration comes from CUPTI.
TAGS = None


class _Timed:
    def __init__(self, kind, flops, tag="", nbytes=0.0):
        self.kind, self.flops, self.tag, self.nbytes = kind, flops, tag, nbytes

    def __enter__(self):
        if TAGS is not None:
            TAGS.append((self.kind, self.flops, self.nbytes, self.tag))
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.kind, self.flops, self.e0, self.e1, self.tag))


def pack_weights(plan, w, row_scale=None):
    """HWIO float kernel [k,k,Cin,Cout] -> packed [rows_pad, Ktot] bf16 (K-major) following plan.pack, in ONE
    kernel launch (kp_pack_weights).  `row_scale` (fwd mode, [Cout]) multiplies each output channel's weights
    (batch-norm folding).  Same recipe as tapconv.pack_weights_np, which the CPU tests pin against the conv oracle."""
    if not (w.is_cuda and w.dtype == torch.float32):
        raise ValueError("pack_weights: w must be a float32 CUDA tensor")
    w = w.contiguous()
    d = tc.pack_desc(plan, tuple(w.shape))
    out = torch.empty((plan.rows_pad, plan.Ktot), device=w.device, dtype=torch.bfloat16)
    rs = None
    if row_scale is not None:
        if plan.pack["mode"] != "fwd":
            raise ValueError("row_scale only applies to forward packing")
        rs = row_scale.to(torch.float32).contiguous()
    with torch.cuda.device(w.device):
        _lib.call("kp_pack_weights", w.data_ptr(), ctypes.byref(d), None if rs is None else rs.data_ptr(), out.data_ptr(),
                  _stream())
    return out


class PackTable:
    """Device job table for kp_pack_weights_batch: re-packs many (plan, fp32 HWIO kernel view, packed bf16 output)
    triples with ONE launch.  The table is uploaded once per job list (never inside a CUDA-graph capture)."""

    def __init__(self, jobs, device):
        lib = _lib.load()
        arr = (tc.PackJob * len(jobs))()
        total = 0
        for i, (plan, w, out) in enumerate(jobs):
            if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
                raise ValueError("PackTable: kernels must be contiguous float32 CUDA tensors")
            d = tc.pack_desc(plan, tuple(w.shape))
            nb = lib.kp_pack_job_blocks(ctypes.byref(d))
            if nb <= 0:
                raise ValueError("PackTable: bad pack descriptor for job %d" % i)
            arr[i].w, arr[i].dst, arr[i].d = w.data_ptr(), out.data_ptr(), d
            arr[i].block_begin, arr[i].n_blocks = total, nb
            total += nb
        self.n_jobs, self.total_blocks = len(jobs), total
        self.keep = [(w, out) for _, w, out in jobs]          # the table holds raw pointers
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = host.to(device)

    def run(self):
        with torch.cuda.device(self.table.device):
            _lib.call("kp_pack_weights_batch", self.table.data_ptr(), self.n_jobs, self.total_blocks, _stream())


def pack_weights_torch(plan, w, row_scale=None, dtype=torch.bfloat16):
    """Reference re-layout with torch ops (tests cross-check the kernel above against it)."""
    k1, k2, cin, cout = w.shape
    CB = plan.CB
    w3 = w.reshape(k1 * k2, cin, cout)
    if list(plan.pack["taps"]) != list(range(k1 * k2)):     # stride-2 dgrad parity classes use a subset of the taps
        cache = plan.__dict__.setdefault("_taps_dev", {})      # device index tensor, built once (not inside a graph capture)
        if w.device not in cache:
            cache[w.device] = torch.as_tensor(plan.pack["taps"], device=w.device, dtype=torch.long)
        w3 = w3.index_select(0, cache[w.device])
    if plan.pack["mode"] == "fwd":
        if row_scale is not None:
            w3 = w3 * row_scale.view(1, 1, cout)
        parts = []
        for c_start, c_count in plan.pack["segs"]:
            seg = w3[:, c_start:c_start + c_count, :]
            padc = tc.round_up(c_count, CB) - c_count
            if padc:
                seg = torch.nn.functional.pad(seg, (0, 0, 0, padc))
            parts.append(seg)
        wk = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        m = wk.permute(2, 0, 1).reshape(cout, -1)
    else:
        c0, c1 = plan.pack["row_slice"]
        seg = w3[:, c0:c1, :]
        padc = tc.round_up(cout, CB) - cout
        if padc:
            seg = torch.nn.functional.pad(seg, (0, padc))
        m = seg.permute(1, 0, 2).reshape(c1 - c0, -1)
    if m.shape[0] != plan.rows_pad:
        m = torch.nn.functional.pad(m, (0, 0, 0, plan.rows_pad - m.shape[0]))
    assert m.shape == (plan.rows_pad, plan.Ktot), (tuple(m.shape), plan.rows_pad, plan.Ktot)
    return m.to(dtype).contiguous()


def pad_vec(v, n):
    """f32 vector padded with zeros to length n (bias / statistics buffers are [Cout_pad])."""
    v = v.to(torch.float32)
    if v.shape[0] == n:
        return v.contiguous()
    return torch.nn.functional.pad(v, (0, n - v.shape[0])).contiguous()


def run_plan(plan, srcs, wpacked, bias, out, act=ACT_NONE, alpha=0.0, stats=None, cout_written=None, tile=None, bn=0,
             accumulate=False, stat_groups=1):
    """Launch one tap-GEMM.  srcs: bf16 NHWC CUDA tensors; out: bf16 or f32 tensor written through the plan's
    output view; bias: f32 [rows_pad] or None; stats: (sum, sumsq) f32 [rows_pad] accumulators or None."""
    for s in srcs:
        if not (s.is_cuda and s.dtype == torch.bfloat16 and s.is_contiguous()):
            raise ValueError("tap-GEMM sources must be contiguous bf16 CUDA tensors")
    if not (wpacked.is_cuda and wpacked.dtype == torch.bfloat16 and wpacked.is_contiguous()
            and tuple(wpacked.shape) == (plan.rows_pad, plan.Ktot)):
        raise ValueError("packed weights must be bf16 [rows_pad=%d, Ktot=%d]" % (plan.rows_pad, plan.Ktot))
    if out.dtype not in (torch.bfloat16, torch.float32) or not out.is_cuda:
        raise ValueError("output must be a bf16 or f32 CUDA tensor")
    if bias is not None and (bias.dtype != torch.float32 or bias.shape[0] != plan.rows_pad):
        raise ValueError("bias must be f32 [rows_pad]")
    d = plan.desc(act=act, alpha=alpha, out_f32=(out.dtype == torch.float32), cout_written=cout_written, tile=tile, bn=bn,
                  accumulate=accumulate, stat_groups=stat_groups)
    ptrs = (ctypes.c_void_p * tc.KP_MAX_MAPS)()
    for i, s in enumerate(srcs):
        ptrs[i] = s.data_ptr()
    ssum = ssq = None
    if stats is not None:
        ssum, ssq = stats
        if ssum.dtype != torch.float32 or ssum.numel() != stat_groups * plan.rows_pad or ssq.numel() != stat_groups * plan.rows_pad:
            raise ValueError("stats buffers must be f32 [stat_groups * rows_pad]")
        if plan.N % stat_groups != 0:
            raise ValueError("batch %d does not split into %d statistics segments" % (plan.N, stat_groups))
    first = plan.taps[0][2]
    k_real = len(plan.taps) * sum(plan.views[first + s]["C"] for s in range(plan.n_src))
    flops = 2.0 * plan.N * plan.Ho * plan.Wo * k_real * plan.rows * getattr(plan, "flop_scale", 1.0)
    tag = "N%d %dx%d K%d(taps %d) -> %d CB%d%s" % (plan.N, plan.Ho, plan.Wo, plan.Ktot, len(plan.taps), plan.rows, plan.CB,
                                                  " +stats" if stats is not None else "") if (PROFILE is not None or TAGS is not None) else ""
    nbytes = 0.0
    if TAGS is not None:
        # algorithmic bytes: every distinct source once, the output view once, the packed weights once
        seen = set()
        for s_ in srcs:
            if s_.data_ptr() not in seen:
                seen.add(s_.data_ptr())
                nbytes += s_.numel() * 2.0
        nbytes += float(plan.N * plan.Ho * plan.Wo * plan.rows) * (4.0 if out.dtype == torch.float32 else 2.0) * (2.0 if accumulate else 1.0)
        nbytes += wpacked.numel() * 2.0
    with torch.cuda.device(out.device), _Timed("dgrad" if plan.pack["mode"] == "dgrad" else "fwd", flops, tag, nbytes):
        _lib.call("kp_tapconv_bf16", ctypes.byref(d), ptrs, wpacked.data_ptr(),
                  None if bias is None else bias.data_ptr(), out.data_ptr(),
                  None if ssum is None else ssum.data_ptr(), None if ssq is None else ssq.data_ptr(), _stream())
    return out


def run_wgrad(plan, x, dy, dw, splits=0):
    """Accumulate (+=) the weight gradient of one source into the f32 HWIO tensor `dw` (zero it first)."""
    for t in (x, dy):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
            raise ValueError("wgrad operands must be contiguous bf16 CUDA tensors")
    if not (dw.is_cuda and dw.dtype == torch.float32 and dw.is_contiguous()):
        raise ValueError("dW must be a contiguous f32 CUDA tensor (HWIO)")
    d = plan.desc(splits)
    flops = 2.0 * plan.N * plan.Ho * plan.Wo * len(plan.taps) * plan.Cin * plan.Cout * getattr(plan, "flop_scale", 1.0)
    tag = "N%d %dx%d Cin%d Cout%d taps %d CB%d" % (plan.N, plan.Ho, plan.Wo, plan.Cin, plan.Cout, len(plan.taps), plan.CB)
    nbytes = x.numel() * 2.0 + dy.numel() * 2.0 + 2.0 * 4.0 * len(plan.taps) * plan.Cin * plan.Cout
    with torch.cuda.device(dw.device), _Timed("wgrad", flops, tag, nbytes):
        _lib.call("kp_tapconv_wgrad_bf16", ctypes.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _stream())
    return dw
