"""Execution engine of the stage-1 path: name-keyed parameter store, backward tape, convolution layers.

This is host-side plumbing (buffers, launch order, gradient bookkeeping); every FLOP runs in
libkp_b200.so.  Parameters use the TF variable names of the reference graph
(`<scope>/conv2d/{kernel,bias}`, `<scope>/{gamma,beta,moving_mean,moving_variance}`; HWIO kernels) so a
state dict can be exchanged with the oracle and, eventually, with reference checkpoints
(models/base_model.py:74-92).
"""
import math

import torch

from . import conv as cv
from . import ops
from . import tapconv as tc

BF16 = torch.bfloat16
F32 = torch.float32


class ParamGroup:
    """One flat f32 buffer (+ grad / Adam slots) holding many named tensors: a single Adam launch and a single
    gradient all-reduce per optimiser step."""

    def __init__(self, device, trainable=True):
        self.device = device
        self.trainable = trainable
        self.specs = []          # (name, shape, offset)
        self.index = {}
        self.total = 0
        self.data = self.grad = self.m = self.v = None

    def add(self, name, shape):
        assert self.data is None and name not in self.index
        n = int(math.prod(shape))
        self.index[name] = len(self.specs)
        self.specs.append((name, tuple(shape), self.total))
        self.total += (n + 3) // 4 * 4      # keep every tensor 16-byte aligned

    def finalize(self):
        self.data = torch.zeros(self.total, device=self.device, dtype=F32)
        if self.trainable:
            self.grad = torch.zeros_like(self.data)
            self.m = torch.zeros_like(self.data)
            self.v = torch.zeros_like(self.data)

    def _view(self, buf, name):
        _, shape, off = self.specs[self.index[name]]
        return buf[off:off + int(math.prod(shape))].view(shape)

    def __contains__(self, name):
        return name in self.index

    def p(self, name):
        return self._view(self.data, name)

    def g(self, name):
        return self._view(self.grad, name)

    def names(self):
        return [s[0] for s in self.specs]


class Tape:
    """Reverse-mode tape: closures run in reverse; gradients of activations are keyed by tensor identity."""

    def __init__(self):
        self.ops = []
        self.g = {}
        self.keep = []
        self.joins = []          # contexts whose weight-gradient stream this backward pass forked (joined at its end)
        self.branch = None       # stream of a forward branch: closures recorded while it is current run on it again

    def record(self, fn):
        on_branch = self.branch is not None and torch.cuda.current_stream() == self.branch
        self.ops.append((fn, self.branch if on_branch else None))

    def grad(self, t):
        return self.g.get(id(t))

    def set_grad(self, t, g):
        self.keep.append(t)
        self.g[id(t)] = g

    def acquire(self, t, dtype=None, shape=None):
        """Gradient buffer of activation `t` and whether it already holds a gradient (=> accumulate)."""
        k = id(t)
        if k in self.g:
            return self.g[k], True
        buf = torch.empty(tuple(t.shape) if shape is None else shape, device=t.device, dtype=dtype or t.dtype)
        self.keep.append(t)
        self.g[k] = buf
        return buf, False

    def run_closures(self):
        """The recorded closures in reverse order, each on the stream it was recorded on; the gradients stay (tests read
        them before clearing)."""
        for fn, st in reversed(self.ops):
            if st is None:
                fn()
            else:
                with torch.cuda.stream(st):
                    fn()

    def backward(self):
        self.run_closures()
        for c in self.joins:
            c.join_wgrad()
        self.joins.clear()
        self.ops.clear()
        self.g.clear()
        self.keep.clear()


class Context:
    """Parameters + execution state shared by the network builders (the analogue of TF's default graph)."""

    def __init__(self, device, n_pts=40):
        self.device = torch.device(device)
        self.n_pts = n_pts
        self.G = ParamGroup(self.device)              # generator: image_encoder, pose_encoder, translator
        self.D = ParamGroup(self.device)              # img_discr
        self.S = ParamGroup(self.device, False)       # BN moving statistics
        self.V = ParamGroup(self.device, False)       # VGG19 constants
        self.version = 0                              # bumped whenever parameter values change
        self.tape = None
        self.update_moving = False
        self.train_D = False                          # accumulate img_discr weight gradients
        self.train_G = False
        self.debug = None                             # dict: layer scope -> internals (tests / probes only)
        self.trace = None                             # list of dicts, one per stored tensor in call order (tests only)
        self._zp = {}                 # slot -> zero pool state
        self._z = None                # the pool of the run being issued
        # Weight gradients on their own stream (set by the model): in the backward pass a layer's weight gradient depends
        # only on (saved input, dy) and nobody needs it before the optimizer, so it runs beside the data-gradient chain
        # (next layer's BN backward = HBM-bound, data gradient = tensor-bound) and fills the tails of those launches.
        self.wgrad_stream = None
        self.branch_stream = None     # independent sub-network (image_encoder beside pose_encoder), see Context.branch()
        self._wg_keep = []            # (dy, inputs) kept alive until the join: the side stream still reads them
        self._wg_pending = False
        self._plans = {}
        self._packed = {}
        self._jobs = {}        # id(group) -> [(key, plan, fp32 kernel view, packed bf16 tensor)]: re-packed in ONE launch
        self._tables = {}      # id(group) -> (number of jobs when built, cv.PackTable)

    # ---- pooled zero-initialised scratch (BN statistics, per-call gradient sums): ONE memset per run ----
    ZPOOL_FLOATS = 4 * 1024 * 1024

    def begin_run(self, slot=0):
        """Call at the start of every forward(/backward) run: re-zeroes the scratch pool with a single memset.
        slot: runs that overlap on different streams (the D run beside the G run) use different pools."""
        z = self._zp.get(slot)
        if z is None:
            z = self._zp[slot] = {"pool": torch.zeros(self.ZPOOL_FLOATS, device=self.device, dtype=F32), "off": 0, "high": 0,
                                  "captured": None}
        else:
            # Re-zero up to the HIGH-WATER mark of all runs so far, not just the previous run's extent: inside a captured
            # CUDA graph this memset has a fixed length, and an eager run in between (test_step) may have used more.
            z["high"] = max(z["high"], z["off"], 1)
            if torch.cuda.is_current_stream_capturing():
                z["captured"] = z["high"]
            elif z["captured"] is not None and z["high"] > z["captured"]:
                # scratch beyond what the captured memset clears was dirtied: clear it now, eagerly
                z["pool"][z["captured"]:z["high"]].zero_()
            z["pool"][:z["high"]].zero_()
        z["off"] = 0
        self._z = z

    def zeros(self, n):
        """n zeroed floats valid until the next begin_run() of the same slot (falls back to torch.zeros when the pool is
        exhausted)."""
        n4 = (n + 3) // 4 * 4
        z = self._z
        if z is None or z["off"] + n4 > z["pool"].numel():
            return torch.zeros(n, device=self.device, dtype=F32)
        out = z["pool"][z["off"]:z["off"] + n]
        z["off"] += n4
        return out

    def branch(self):
        """Context manager: run an independent sub-network on the branch stream, forward and (through the tape) backward.
        Usage: `with ctx.branch(): y = net(x)` ... independent work on the current stream ... `ctx.branch_join()`."""
        return _Branch(self)

    def branch_join(self):
        br, tape = self.branch_stream, self.tape
        if br is None:
            return
        torch.cuda.current_stream().wait_stream(br)
        if tape is not None:
            # backward: the branch's closures (recorded before this point) start after the gradients produced so far
            tape.record(lambda: br.wait_stream(torch.cuda.current_stream()))

    def join_wgrad(self, waiter=None):
        """Make `waiter` (default: the current stream) wait for the weight gradients launched so far."""
        if self._wg_pending:
            (waiter or torch.cuda.current_stream()).wait_stream(self.wgrad_stream)
            if waiter is None:
                self._wg_pending = False
                self._wg_keep.clear()

    # ---- parameter lookup ----
    def group_of(self, name):
        for g in (self.G, self.D, self.S, self.V):
            if name in g:
                return g
        raise KeyError(name)

    def p(self, name):
        return self.group_of(name).p(name)

    def has(self, name):
        return any(name in g for g in (self.G, self.D, self.S, self.V))

    def params_changed(self, group=None):
        """Invalidate packed-weight caches: all of them, or only those built from `group`'s variables."""
        self.version += 1
        groups = [group] if group is not None else [self.G, self.D, self.S, self.V]
        keep = set()
        for g in groups:
            keep.update(j[0] for j in self._jobs.get(id(g), []))
        for key in list(self._packed):
            if key in keep:
                continue
            if group is None or (isinstance(key[1], tuple) and key[1] and key[1][0] in group):
                del self._packed[key]
        for g in groups:
            jobs = self._jobs.get(id(g), [])
            if not jobs:
                continue
            # re-pack every registered kernel of this optimiser in place, now, with one launch (the packed tensors keep
            # their addresses: a captured CUDA graph keeps reading them)
            built = self._tables.get(id(g))
            if built is None or built[0] != len(jobs):
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("pack table changed during CUDA-graph capture (run two eager steps first)")
                built = (len(jobs), cv.PackTable([(pl, w, out) for _, pl, w, out in jobs], self.device))
                self._tables[id(g)] = built
            built[1].run()

    def load_state_dict(self, sd):
        """sd: name -> tensor/ndarray (HWIO kernels).  Unknown names are ignored, like BaseModel.restore."""
        loaded = []
        for name, val in sd.items():
            if self.has(name):
                self.p(name).copy_(torch.as_tensor(val).to(self.device, F32))
                loaded.append(name)
        self.params_changed()
        return loaded

    def state_dict(self):
        out = {}
        for g in (self.G, self.D, self.S, self.V):
            for n in g.names():
                out[n] = g.p(n).detach().clone()
        return out

    # ---- plan / packed-weight caches ----
    def plan(self, kind, key, builder):
        k = (kind,) + key
        if k not in self._plans:
            self._plans[k] = builder()
        return self._plans[k]

    def packed(self, key, builder):
        if key not in self._packed:
            self._packed[key] = builder()
        return self._packed[key]

    def packed_weight(self, kind, wnames, plan, w):
        """bf16 packed copy of the kernel `w` for `plan`.  A kernel that is a plain view of ONE variable is registered
        for the batched in-place re-pack that params_changed(group) runs; fused kernels (several variables
        concatenated) are rebuilt on demand."""
        key = (kind, tuple(wnames), tc.pack_key(plan))
        out = self._packed.get(key)
        if out is None:
            out = cv.pack_weights(plan, w)
            self._packed[key] = out
            if len(wnames) == 1:
                self._jobs.setdefault(id(self.group_of(wnames[0])), []).append((key, plan, w, out))
        return out


# --------------------------------------------------------------------------------------------------
# convolution layer (+ BN / activation), forward and tape entry
# --------------------------------------------------------------------------------------------------
class _Branch:
    def __init__(self, ctx):
        self.ctx = ctx
        self.cm = None

    def __enter__(self):
        ctx = self.ctx
        br, tape = ctx.branch_stream, ctx.tape
        if br is None:
            return self
        br.wait_stream(torch.cuda.current_stream())
        if tape is not None:
            tape.branch = br
            # backward: runs after the branch's closures were issued -> the main chain waits for them here
            tape.record(lambda: torch.cuda.current_stream().wait_stream(br))
        self.cm = torch.cuda.stream(br)
        self.cm.__enter__()
        return self

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)
        return False


def _conv_weights(ctx, wnames, wshape=None):
    """HWIO kernel of the layer; several TF variables may be fused along Cout (translator heads).  `wshape`
    reinterprets the (contiguous) variable in place, e.g. [7,7,3,32] as [7,1,21,32] for the W-unrolled first layer."""
    if len(wnames) == 1:
        w = ctx.p(wnames[0])
        return w if wshape is None else w.view(wshape)
    return torch.cat([ctx.p(n) for n in wnames], dim=3)


def _scope_of(wname):
    for suffix in ("/conv2d/kernel", "/filter"):
        if wname.endswith(suffix):
            return wname[:-len(suffix)]
    return wname


def _bias_vec(ctx, bnames, rows_pad):
    if bnames is None:
        return None
    parts = [ctx.p(n) for n in bnames]
    b = parts[0] if len(parts) == 1 else torch.cat(parts)
    return cv.pad_vec(b, rows_pad)


def conv_layer(ctx, srcs, wnames, bnames, k, stride, pad, *, bn=None, train_mode=False, act=tc.ACT_NONE, alpha=0.0,
               upsample=False, out_f32=False, need_input_grad=True, wshape=None, segments=1, grad_rows=None):
    """layers.conv [+ layers.batch_norm + relu] [+ resize x2] of the reference, on bf16 NHWC tensors.

    srcs: list of bf16 [N,H,W,C] tensors (virtual channel concat).  Returns the layer output (bf16, or f32 when
    out_f32).  bn: BN scope name or None.  With bn and train_mode the batch statistics are used (and the
    moving averages updated when ctx.update_moving); with bn and not train_mode BN is folded into the
    convolution.  A tape entry is recorded when ctx.tape is set.
    segments > 1: the batch holds that many calls of a shared-weight network side by side (pose_encoder on image and
    future_image, detector_translator_model.py:166-167): one launch per layer, batch-norm statistics, moving-average
    updates and the BN backward per segment, exactly as separate calls would compute them.
    grad_rows=(lo, hi): only the images [lo, hi) of the batch carry a gradient (VGG on [gt; pred], :274-279): the
    backward pass runs on that contiguous slice alone.
    """
    wnames = [wnames] if isinstance(wnames, str) else list(wnames)
    if isinstance(bnames, str):
        bnames = [bnames]
    shapes = tuple(tuple(s.shape) for s in srcs)
    w = _conv_weights(ctx, wnames, wshape)
    cout = w.shape[3]
    cin_real = w.shape[2]
    cin_src = sum(s[3] for s in shapes)
    flop_scale = w.shape[2] / float(cin_src)      # algorithmic-FLOP accounting ignores zero-padded channels
    # Sources may carry zero-padded channels (3-channel images stored as 16, the 208-channel joint embedding as 256)
    # and Cout is padded to the channel block: the pack kernels zero-fill everything outside the real [cin, cout].
    fplan, (N, Ho, Wo) = ctx.plan("fwd", (shapes, k, stride, pad, cout), lambda: tc.plan_conv_fwd(list(shapes), k, stride, pad, cout))
    fplan.flop_scale = flop_scale
    dev = ctx.device
    group = ctx.group_of(wnames[0])
    wkey = tuple(wnames)

    if bn is not None and not train_mode:
        # inference: fold BN (moving statistics) into the convolution, ReLU in the epilogue
        if ctx.tape is not None:
            raise RuntimeError("conv_layer(%s): inference-mode (folded) batch norm records no backward; gradients through "
                               "an is_training=False stack are not supported" % bn)
        def build():
            scale = ctx.p(bn + "/gamma") * torch.rsqrt(ctx.p(bn + "/moving_variance") + 1e-5)
            b = ctx.p(bnames[0]) if bnames else torch.zeros(cout, device=dev)
            bias = (b - ctx.p(bn + "/moving_mean")) * scale + ctx.p(bn + "/beta")
            return cv.pack_weights(fplan, w, row_scale=scale), cv.pad_vec(bias, fplan.rows_pad)
        wp, bias = ctx.packed(("fold", wkey, shapes), build)
        y = torch.empty((N, Ho, Wo, cout), device=dev, dtype=BF16)
        cv.run_plan(fplan, srcs, wp, bias, y, act=tc.ACT_RELU)
        if upsample:
            y = ops.bn_act_apply(y, None, None, relu=False, upsample=True)
        return y

    wp = ctx.packed_weight("fwd", wnames, fplan, w)
    bias = ctx.packed(("bias", wkey, fplan.rows_pad), lambda: _bias_vec(ctx, bnames, fplan.rows_pad))

    if bn is not None:
        # training-mode BN: conv (+bias) with statistics in the epilogue -> finalize -> normalise + ReLU (+ x2)
        assert fplan.rows_pad == cout or segments == 1, "segmented batch norm needs Cout to be a multiple of 16"
        ssum = ctx.zeros(segments * fplan.rows_pad)
        ssq = ctx.zeros(segments * fplan.rows_pad)
        y_pre = torch.empty((N, Ho, Wo, cout), device=dev, dtype=BF16)
        cv.run_plan(fplan, srcs, wp, bias, y_pre, act=tc.ACT_NONE, stats=(ssum, ssq), stat_groups=segments)
        mm = ctx.p(bn + "/moving_mean") if ctx.update_moving else None
        mv = ctx.p(bn + "/moving_variance") if ctx.update_moving else None
        out, scale, shift, mean, rstd = ops.bn_stats_apply(y_pre, ssum, ssq, bias, ctx.p(bn + "/gamma"), ctx.p(bn + "/beta"),
                                                           (N // segments) * Ho * Wo, mm, mv, relu=True, upsample=upsample,
                                                           segments=segments)
        if ctx.debug is not None:
            ctx.debug[wnames[0].replace("/conv2d/kernel", "")] = (out, y_pre, scale, shift, mean, rstd, upsample)
        if ctx.trace is not None:
            ctx.trace.append(dict(scope=_scope_of(wnames[0]), kind="bn", out=out, y_pre=y_pre, mean=mean, rstd=rstd, segments=segments))
        if ctx.tape is not None:
            tape = ctx.tape

            def bwd():
                dout = tape.grad(out)
                if dout is None:
                    return
                train = (group is ctx.G and ctx.train_G) or (group is ctx.D and ctx.train_D)
                dy, _, _ = ops.bn_act_bwd(dout, y_pre, scale, shift, mean, rstd, relu=True, upsample=upsample,
                                          gbeta_acc=group.g(bn + "/beta") if train else None,
                                          ggamma_acc=group.g(bn + "/gamma") if train else None,
                                          zeroed=ctx.zeros(2 * segments * cout), segments=segments)
                _conv_backward(ctx, tape, srcs, shapes, wnames, None, w, k, stride, pad, cout, dy, group, need_input_grad,
                               cin_real, wshape)
            tape.record(bwd)
        return out

    # plain conv + bias + activation in the epilogue
    k_h, k_w = tc._khw(k)
    if (out_f32 and act == tc.ACT_NONE and k_h == 1 and k_w == 1 and stride == 1 and pad == 0 and len(srcs) == 1
            and shapes[0][3] == 16 and cin_real == 16 and len(wnames) == 1 and cout % 4 == 0 and cout <= 40 and wshape is None):
        # HBM-bound 1x1 head (16 -> n_pts fp32 logits): dedicated kernel with coalesced fp32 stores (csrc/head1x1.cu); it reads
        # the fp32 master weights in place and rounds them to bf16 on load, like the packed copies of the tensor-core path
        y = ops.conv1x1_f32(srcs[0], w, ctx.p(bnames[0]) if bnames else None)
    else:
        y = torch.empty((N, Ho, Wo, cout), device=dev, dtype=F32 if out_f32 else BF16)
        cv.run_plan(fplan, srcs, wp, bias, y, act=act, alpha=alpha)
    if ctx.trace is not None:
        ctx.trace.append(dict(scope=_scope_of(wnames[0]), kind="plain", out=y))
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            dy = tape.grad(y)          # bf16, channels padded to a multiple of 8 for f32-output layers
            if dy is None:
                return
            yb, sb, shp_b = y, srcs, shapes
            if grad_rows is not None:
                # only a contiguous run of images carries a gradient: the whole backward works on that slice
                lo, hi = grad_rows
                dy, yb = dy[lo:hi], y[lo:hi]
                sb = [s_[lo:hi] for s_ in srcs]
                shp_b = tuple((hi - lo,) + tuple(s_[1:]) for s_ in shapes)
            if act == tc.ACT_RELU:
                dy = ops.act_mask_bwd(dy, yb, 0.0)
            elif act == tc.ACT_LEAKY:
                dy = ops.act_mask_bwd(dy, yb, alpha)
            _conv_backward(ctx, tape, sb, shp_b, wnames, bnames, w, k, stride, pad, cout, dy, group, need_input_grad,
                           cin_real, wshape, parents=srcs if grad_rows is not None else None, rows=grad_rows)
        tape.record(bwd)
    return y


def _weight_grads(ctx, srcs, shapes, wnames, bnames, k, stride, pad, cout, dy, group, cin_real, wshape, cin, cpad):
    """Weight and bias gradients of one convolution (accumulated into the group's flat gradient buffer)."""
    k_h, k_w = tc._khw(k)
    gview = (lambda n: group.g(n)) if wshape is None else (lambda n: group.g(n).view(wshape))
    direct = cpad == cout and len(wnames) == 1 and cin_real == cin
    if direct:
        gw = gview(wnames[0])
    else:
        gw = ctx.zeros(k_h * k_w * cin * cpad).view(k_h, k_w, cin, cpad)
    c0 = 0
    for s, shp in zip(srcs, shapes):
        C = shp[3]
        wplan = ctx.plan("wgrad", (shp, k, stride, pad, cpad, c0, cin),
                         lambda shp=shp, c0=c0, C=C: tc.plan_conv_wgrad(shp, k, stride, pad, cpad, cin_slice=(c0, c0 + C, cin)))
        wplan.flop_scale = (cin_real / float(cin)) * (cout / float(cpad))
        cv.run_wgrad(wplan, s, dy, gw)
        c0 += C
    if not direct:
        o = 0
        for n in wnames:
            co = ctx.p(n).shape[3]
            gview(n).add_(gw[:, :, :cin_real, o:o + co])
            o += co
    if bnames:
        gb = ctx.zeros(cpad)
        ops.channel_sum(dy, gb)
        o = 0
        for n in bnames:
            co = ctx.p(n).shape[0]
            group.g(n).add_(gb[o:o + co])
            o += co


def _conv_backward(ctx, tape, srcs, shapes, wnames, bnames, w, k, stride, pad, cout, dy, group, need_input_grad,
                           cin_real, wshape, parents=None, rows=None):
    """dy: bf16 gradient w.r.t. the convolution output, [N,Ho,Wo,cpad] with cpad = round_up(cout, 8).
    parents / rows: `srcs` are the image slices [rows[0], rows[1]) of `parents`; the data gradient goes into the same slice
    of the parents' gradient buffers."""
    cpad = dy.shape[3]
    train = group.trainable and ((group is ctx.G and ctx.train_G) or (group is ctx.D and ctx.train_D))
    cin = sum(s[3] for s in shapes)
    if train:
        ws = ctx.wgrad_stream
        if ws is not None:
            ws.wait_stream(torch.cuda.current_stream())        # dy is complete on the main stream
            ctx._wg_keep.append((dy, srcs))
            ctx._wg_pending = True
            if ctx not in tape.joins:
                tape.joins.append(ctx)
            with torch.cuda.stream(ws):
                _weight_grads(ctx, srcs, shapes, wnames, bnames, k, stride, pad, cout, dy, group, cin_real, wshape, cin, cpad)
        else:
            _weight_grads(ctx, srcs, shapes, wnames, bnames, k, stride, pad, cout, dy, group, cin_real, wshape, cin, cpad)
    if not need_input_grad:
        return
    c0 = 0
    for si, (s, shp) in enumerate(zip(srcs, shapes)):
        C = shp[3]
        plans = ctx.plan("dgrad", (shp, k, stride, pad, cpad, c0, cin),
                         lambda shp=shp, c0=c0, C=C: tc.plan_conv_dgrad(shp, k, stride, pad, cpad, cin_slice=(c0, c0 + C, cin)))
        # Stride-2 data gradients of the deep, small-image layers (img_discr conv_3..5: a few hundred pixels, K = 4 taps x
        # 512..2048 channels): skinny GEMMs that fill the machine only when K is split over CTAs.  They accumulate in fp32
        # (atomic adds into a zeroed tensor shared by the four parity launches) and are cast to bf16 afterwards.
        n_px = shp[0] * shp[1] * shp[2]
        if (parents is None and stride == 2 and tape.grad(s) is None and n_px <= 128 * 64 and C % 8 == 0
                and min(pl.Ktot for pl in plans) >= 2048):
            tmp = torch.zeros(tuple(shp), device=ctx.device, dtype=F32)
            for p in plans:
                p.flop_scale = (cin_real / float(cin)) * (cout / float(cpad))
                cv.run_plan(p, [dy], ctx.packed_weight("dgrad", wnames, p, w), None, tmp, accumulate=2)
            tape.set_grad(s, ops.pack_channels([tmp], C))
            c0 += C
            continue
        if parents is None:
            dx, acc = tape.acquire(s)
        else:
            full, acc = tape.acquire(parents[si])
            dx = full[rows[0]:rows[1]]
        for i, p in enumerate(plans):
            p.flop_scale = (cin_real / float(cin)) * (cout / float(cpad))
            wp = ctx.packed_weight("dgrad", wnames, p, w)
            cv.run_plan(p, [dy], wp, None, dx, accumulate=acc)
        c0 += C
