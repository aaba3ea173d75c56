"""Host-side lowering of TF-style convolutions onto the tap-GEMM primitive (kp_tapconv_bf16).

Pure geometry — no torch, no CUDA — so that it can be unit-tested on a CPU box against the oracle
(tests/test_tapconv_lowering.py runs these plans through a numpy emulator of the device primitive).

Reference semantics being lowered: layers.conv (/root/reference/models/networks/layers.py:4-10) =
tf.pad(pad) + tf.layers.conv2d(padding='same', strides=s); its data-gradient; channel concats feeding a
convolution (models/networks/__init__.py:44) as several sources per tap.
"""
import ctypes

import numpy as np

KP_MAX_MAPS = 4
KP_MAX_TAPS = 64

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_SIGMOID_LAST = 0, 1, 2, 3, 4


class TapView(ctypes.Structure):
    _fields_ = [("src", ctypes.c_int), ("C", ctypes.c_int), ("Wd", ctypes.c_int), ("Hd", ctypes.c_int),
                ("off", ctypes.c_longlong), ("sw", ctypes.c_longlong), ("sh", ctypes.c_longlong),
                ("sn", ctypes.c_longlong)]


class TapConvDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("n_maps", ctypes.c_int), ("map", TapView * KP_MAX_MAPS),
                ("n_taps", ctypes.c_int), ("n_src", ctypes.c_int),
                ("dh", ctypes.c_byte * KP_MAX_TAPS), ("dw", ctypes.c_byte * KP_MAX_TAPS),
                ("map_first", ctypes.c_byte * KP_MAX_TAPS),
                ("CB", ctypes.c_int), ("Cout_pad", ctypes.c_int), ("Ktot", ctypes.c_int),
                ("Ho", ctypes.c_int), ("Wo", ctypes.c_int),
                ("out_off", ctypes.c_longlong), ("out_sw", ctypes.c_longlong), ("out_sh", ctypes.c_longlong),
                ("out_sn", ctypes.c_longlong),
                ("Cout", ctypes.c_int), ("out_f32", ctypes.c_int), ("act", ctypes.c_int), ("alpha", ctypes.c_float),
                ("accumulate", ctypes.c_int),
                ("TW", ctypes.c_int), ("TH", ctypes.c_int), ("TN", ctypes.c_int), ("BN", ctypes.c_int),
                ("stat_groups", ctypes.c_int)]


def same_pad(in_size, k, s):
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def round_up(v, m):
    return -(-v // m) * m


def _khw(k):
    """Kernel size as (kh, kw); an int means a square kernel."""
    return (int(k[0]), int(k[1])) if isinstance(k, (tuple, list)) else (int(k), int(k))


def pick_cb(channels):
    for cb in (64, 32):
        if all(c % cb == 0 for c in channels):
            return cb
    return 16


class TapPlan:
    """One launch of the device primitive: views, taps, weight gather index, output view."""

    def __init__(self):
        self.N = 0
        self.views = []        # dicts: src, C, Wd, Hd, off, sw, sh, sn
        self.n_src = 1
        self.taps = []         # (dh, dw, map_first)
        self.CB = 16
        self.rows = 0          # real rows of the packed matrix (output channels of this GEMM)
        self.rows_pad = 0
        self.Ktot = 0
        self.Ho = self.Wo = 0  # tile-space output extent
        self.out_off = 0
        self.out_sw = self.out_sh = self.out_sn = 0
        self.pack = None       # weight packing recipe (mode, taps, segs, row_slice)

    def blocks_per_tap(self):
        first = self.taps[0][2]
        return sum(-(-self.views[first + s]["C"] // self.CB) for s in range(self.n_src))

    def desc(self, act=ACT_NONE, alpha=0.0, out_f32=False, cout_written=None, tile=None, bn=0, accumulate=False, stat_groups=1):
        d = TapConvDesc()
        d.N = self.N
        d.n_maps = len(self.views)
        for i, v in enumerate(self.views):
            d.map[i].src, d.map[i].C, d.map[i].Wd, d.map[i].Hd = v["src"], v["C"], v["Wd"], v["Hd"]
            d.map[i].off, d.map[i].sw, d.map[i].sh, d.map[i].sn = v["off"], v["sw"], v["sh"], v["sn"]
        d.n_taps = len(self.taps)
        d.n_src = self.n_src
        for i, (dh, dw, mf) in enumerate(self.taps):
            d.dh[i], d.dw[i], d.map_first[i] = dh, dw, mf
        d.CB, d.Cout_pad, d.Ktot = self.CB, self.rows_pad, self.Ktot
        d.Ho, d.Wo = self.Ho, self.Wo
        d.out_off, d.out_sw, d.out_sh, d.out_sn = self.out_off, self.out_sw, self.out_sh, self.out_sn
        d.Cout = self.rows if cout_written is None else cout_written
        d.out_f32 = 1 if out_f32 else 0
        d.act, d.alpha = act, alpha
        d.accumulate = int(accumulate) if accumulate in (0, 1, 2) else (1 if accumulate else 0)
        if tile is not None:
            d.TW, d.TH, d.TN = tile
        d.BN = bn
        d.stat_groups = stat_groups
        return d


def _finish(plan, rows, mode, tap_flat, segs, row_slice=None):
    """Record the weight-packing recipe (see pack_weights_np / ops.pack_weights)."""
    CB = plan.CB
    plan.rows = rows
    plan.rows_pad = round_up(rows, 16)
    per_tap = sum(round_up(c, CB) for _, c in segs)
    plan.Ktot = per_tap * len(plan.taps)
    plan.pack = dict(mode=mode, taps=list(tap_flat), segs=list(segs), row_slice=row_slice)
    return plan


def pack_weights_np(plan, w):
    """numpy twin of the device-side weight packing: HWIO kernel -> [rows_pad, Ktot] (K-major)."""
    k1, k2, cin, cout = w.shape
    w3 = w.reshape(k1 * k2, cin, cout)[plan.pack["taps"]]            # [T, cin, cout]
    CB = plan.CB
    if plan.pack["mode"] == "fwd":
        parts = []
        for c_start, c_count in plan.pack["segs"]:
            seg = w3[:, c_start:c_start + c_count, :]
            padc = round_up(c_count, CB) - c_count
            parts.append(np.pad(seg, ((0, 0), (0, padc), (0, 0))))
        wk = np.concatenate(parts, axis=1)                              # [T, Kper, cout]
        m = np.transpose(wk, (2, 0, 1)).reshape(cout, -1)
    else:
        c0, c1 = plan.pack["row_slice"]
        seg = w3[:, c0:c1, :]                                           # [T, rows, cout]
        padc = round_up(cout, CB) - cout
        seg = np.pad(seg, ((0, 0), (0, 0), (0, padc)))
        m = np.transpose(seg, (1, 0, 2)).reshape(c1 - c0, -1)
    out = np.zeros((plan.rows_pad, plan.Ktot), dtype=w.dtype)
    out[:m.shape[0]] = m
    return out


def plan_conv_fwd(src_shapes, k, stride, pad, cout, out_channels_total=None, out_channel_off=0):
    """Forward convolution of the channel-concat of `src_shapes` [(N,H,W,C), ...] with an HWIO kernel
    [k,k,sum(C),cout].  Output tensor is NHWC with `out_channels_total` channels (default cout); this conv
    writes channels [out_channel_off, out_channel_off+cout).  Returns (plan, (N,Ho,Wo))."""
    N, H, W, _ = src_shapes[0]
    for s in src_shapes:
        assert s[:3] == (N, H, W), "concat sources must share N,H,W"
    Cs = [s[3] for s in src_shapes]
    cin = sum(Cs)
    k_h, k_w = _khw(k)
    Hp, Wp = H + 2 * pad, W + 2 * pad
    Ho, Wo = -(-Hp // stride), -(-Wp // stride)
    pt = pad + same_pad(Hp, k_h, stride)[0]
    pl = pad + same_pad(Wp, k_w, stride)[0]
    p = TapPlan()
    p.N = N
    p.CB = pick_cb(Cs)
    p.n_src = len(Cs)
    tap_k = []
    if stride == 1:
        for i, C in enumerate(Cs):
            p.views.append(dict(src=i, C=C, Wd=W, Hd=H, off=0, sw=C, sh=W * C, sn=H * W * C))
        for kh in range(k_h):
            for kw in range(k_w):
                p.taps.append((kh - pt, kw - pl, 0))
                tap_k.append((kh, kw))
    elif stride == 2:
        assert len(Cs) == 1, "stride-2 convolutions take a single source on this path"
        C = Cs[0]
        used = {}
        for kh in range(k_h):
            for kw in range(k_w):
                eh, ew = kh - pt, kw - pl
                ph, pw = eh % 2, ew % 2
                key = (ph, pw)
                if key not in used:
                    used[key] = len(p.views)
                    p.views.append(dict(src=0, C=C, Wd=-(-(W - pw) // 2), Hd=-(-(H - ph) // 2), off=(ph * W + pw) * C,
                                        sw=2 * C, sh=2 * W * C, sn=H * W * C))
                p.taps.append(((eh - ph) // 2, (ew - pw) // 2, used[key]))
                tap_k.append((kh, kw))
    else:
        raise ValueError("stride must be 1 or 2")
    assert len(p.taps) <= KP_MAX_TAPS and len(p.views) <= KP_MAX_MAPS
    ct = cout if out_channels_total is None else out_channels_total
    p.Ho, p.Wo = Ho, Wo
    p.out_off, p.out_sw, p.out_sh, p.out_sn = out_channel_off, ct, Wo * ct, Ho * Wo * ct

    segs, base = [], 0
    for C in Cs:
        segs.append((base, C))
        base += C
    _finish(p, cout, "fwd", [kh * k_w + kw for kh, kw in tap_k], segs)
    return p, (N, Ho, Wo)


def plan_conv_dgrad(x_shape, k, stride, pad, cout, cin_slice=None):
    """Data gradient of plan_conv_fwd: dY [N,Ho,Wo,cout] -> dX [N,H,W,Cin] (Cin = x_shape[3]).

    `cin_slice=(c0, c1, cin_total)`: produce only input channels [c0,c1) of a kernel with cin_total input
    channels (gradient w.r.t. ONE source of a virtual concat).  Returns a list of plans (1 for stride 1,
    up to 4 parity classes for stride 2); each writes a strided view of dX.
    """
    N, H, W, Cx = x_shape
    c0, c1, cin_total = (0, Cx, Cx) if cin_slice is None else cin_slice
    assert c1 - c0 == Cx
    k_h, k_w = _khw(k)
    Hp, Wp = H + 2 * pad, W + 2 * pad
    Ho, Wo = -(-Hp // stride), -(-Wp // stride)
    pt = pad + same_pad(Hp, k_h, stride)[0]
    pl = pad + same_pad(Wp, k_w, stride)[0]
    plans = []
    classes = [(0, 0)] if stride == 1 else [(0, 0), (0, 1), (1, 0), (1, 1)]
    for ah, aw in classes:
        p = TapPlan()
        p.N = N
        p.CB = pick_cb([cout])
        p.n_src = 1
        p.views.append(dict(src=0, C=cout, Wd=Wo, Hd=Ho, off=0, sw=cout, sh=Wo * cout, sn=Ho * Wo * cout))
        tap_k = []
        for kh in range(k_h):
            for kw in range(k_w):
                if stride == 1:
                    p.taps.append((pt - kh, pl - kw, 0))
                    tap_k.append((kh, kw))
                else:
                    if (ah - kh + pt) % 2 or (aw - kw + pl) % 2:
                        continue
                    p.taps.append(((ah - kh + pt) // 2, (aw - kw + pl) // 2, 0))
                    tap_k.append((kh, kw))
        if not p.taps:
            continue
        if stride == 1:
            p.Ho, p.Wo = H, W
            p.out_off, p.out_sw, p.out_sh, p.out_sn = 0, Cx, W * Cx, H * W * Cx
        else:
            p.Ho, p.Wo = -(-(H - ah) // 2), -(-(W - aw) // 2)
            if p.Ho <= 0 or p.Wo <= 0:
                continue
            p.out_off, p.out_sw, p.out_sh, p.out_sn = (ah * W + aw) * Cx, 2 * Cx, 2 * W * Cx, H * W * Cx

        _finish(p, Cx, "dgrad", [kh * k_w + kw for kh, kw in tap_k], [(0, cout)], row_slice=(c0, c1))
        plans.append(p)
    return plans


# ------------------------------------------------------------------------------------------------
# weight gradient
# ------------------------------------------------------------------------------------------------
class WgradDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("n_maps", ctypes.c_int), ("map", TapView * KP_MAX_MAPS), ("dy", TapView),
                ("n_taps", ctypes.c_int),
                ("dh", ctypes.c_byte * KP_MAX_TAPS), ("dw", ctypes.c_byte * KP_MAX_TAPS),
                ("map_first", ctypes.c_byte * KP_MAX_TAPS), ("tap_flat", ctypes.c_int * KP_MAX_TAPS),
                ("CB", ctypes.c_int), ("Ho", ctypes.c_int), ("Wo", ctypes.c_int),
                ("Cin", ctypes.c_int), ("Cout", ctypes.c_int),
                ("dw_off", ctypes.c_longlong), ("dw_stap", ctypes.c_longlong), ("dw_sci", ctypes.c_longlong),
                ("splits", ctypes.c_int)]


class WgradPlan:
    """dW[kh,kw,c0:c1,:] of one concat source: X views + taps (as in the forward plan) and the dY view."""

    def __init__(self):
        self.N = 0
        self.views = []
        self.dy = None
        self.taps = []       # (dh, dw, map, tap_flat)
        self.CB = 16
        self.Ho = self.Wo = 0
        self.Cin = self.Cout = 0
        self.dw_off = self.dw_stap = self.dw_sci = 0

    def desc(self, splits=0):
        d = WgradDesc()
        d.N = self.N
        d.n_maps = len(self.views)
        for i, v in enumerate(self.views):
            d.map[i].src, d.map[i].C, d.map[i].Wd, d.map[i].Hd = 0, v["C"], v["Wd"], v["Hd"]
            d.map[i].off, d.map[i].sw, d.map[i].sh, d.map[i].sn = v["off"], v["sw"], v["sh"], v["sn"]
        v = self.dy
        d.dy.src, d.dy.C, d.dy.Wd, d.dy.Hd = 0, v["C"], v["Wd"], v["Hd"]
        d.dy.off, d.dy.sw, d.dy.sh, d.dy.sn = v["off"], v["sw"], v["sh"], v["sn"]
        d.n_taps = len(self.taps)
        for i, (dh, dw, mf, tf) in enumerate(self.taps):
            d.dh[i], d.dw[i], d.map_first[i], d.tap_flat[i] = dh, dw, mf, tf
        d.CB, d.Ho, d.Wo, d.Cin, d.Cout = self.CB, self.Ho, self.Wo, self.Cin, self.Cout
        d.dw_off, d.dw_stap, d.dw_sci = self.dw_off, self.dw_stap, self.dw_sci
        d.splits = splits
        return d


def plan_conv_wgrad(x_shape, k, stride, pad, cout, cin_slice=None):
    """Weight gradient w.r.t. the [k,k,cin_total,cout] HWIO kernel rows [c0,c1) given X = one source
    [N,H,W,C] and dY [N,Ho,Wo,cout] (cout must be a multiple of 8 in memory)."""
    N, H, W, Cx = x_shape
    c0, c1, cin_total = (0, Cx, Cx) if cin_slice is None else cin_slice
    assert c1 - c0 == Cx
    fwd, (_, Ho, Wo) = plan_conv_fwd([x_shape], k, stride, pad, cout)
    p = WgradPlan()
    p.N = N
    p.views = fwd.views
    p.taps = [(dh, dw, mf, tf) for (dh, dw, mf), tf in zip(fwd.taps, fwd.pack["taps"])]
    p.dy = dict(src=0, C=cout, Wd=Wo, Hd=Ho, off=0, sw=cout, sh=Wo * cout, sn=Ho * Wo * cout)
    p.CB = pick_cb([Cx, cout])
    p.Ho, p.Wo = Ho, Wo
    p.Cin, p.Cout = Cx, cout
    p.dw_off, p.dw_stap, p.dw_sci = c0 * cout, cin_total * cout, cout
    return p


# ------------------------------------------------------------------------------------------------
# weight packing descriptor (kp_pack_weights)
# ------------------------------------------------------------------------------------------------
class PackDesc(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int), ("T", ctypes.c_int), ("tap_flat", ctypes.c_int * KP_MAX_TAPS),
                ("cin", ctypes.c_int), ("cout", ctypes.c_int),
                ("nseg", ctypes.c_int), ("seg_start", ctypes.c_int * 3), ("seg_count", ctypes.c_int * 3),
                ("seg_kbase", ctypes.c_int * 3),
                ("c0", ctypes.c_int), ("rows", ctypes.c_int),
                ("Kper", ctypes.c_int), ("rows_pad", ctypes.c_int), ("Ktot", ctypes.c_int)]


class PackJob(ctypes.Structure):
    """kp_pack_job: one entry of the device job table of kp_pack_weights_batch."""
    _fields_ = [("w", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("d", PackDesc), ("block_begin", ctypes.c_int),
                ("n_blocks", ctypes.c_int)]


def pack_key(plan):
    """What the packed layout depends on (NOT the batch / image size): convolutions of the same variable at different
    batch sizes (img_discr on [real;fake] and on fake alone) share one packed copy."""
    pk = plan.pack
    extra = tuple(tuple(s) for s in pk["segs"]) if pk["mode"] == "fwd" else tuple(pk["row_slice"])
    return (pk["mode"], tuple(pk["taps"]), extra, plan.rows_pad, plan.Ktot, plan.CB)


def pack_desc(plan, w_shape):
    """Descriptor of the device weight re-layout for `plan` and an HWIO kernel of shape w_shape (same recipe as
    pack_weights_np)."""
    k1, k2, cin, cout = w_shape
    d = PackDesc()
    taps = plan.pack["taps"]
    d.T = len(taps)
    for i, t in enumerate(taps):
        d.tap_flat[i] = t
    d.cin, d.cout = cin, cout
    d.rows_pad, d.Ktot = plan.rows_pad, plan.Ktot
    d.Kper = plan.Ktot // len(taps)
    if plan.pack["mode"] == "fwd":
        d.mode = 0
        d.nseg = len(plan.pack["segs"])
        kb = 0
        for i, (c_start, c_count) in enumerate(plan.pack["segs"]):
            d.seg_start[i], d.seg_count[i], d.seg_kbase[i] = c_start, c_count, kb
            kb += round_up(c_count, plan.CB)
        assert kb == d.Kper
    else:
        d.mode = 1
        c0, c1 = plan.pack["row_slice"]
        d.c0, d.rows = c0, c1 - c0
        assert d.Kper == round_up(cout, plan.CB)
    return d
