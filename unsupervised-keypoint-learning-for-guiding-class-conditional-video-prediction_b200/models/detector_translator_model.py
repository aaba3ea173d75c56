"""DetectorTranslatorModel — the stage-1 trainer of /root/reference/models/detector_translator_model.py on B200.

Keeps the reference's surface (`build(inputs)`, `train_step`, `test_step`, `collect_test_results`, attribute
names) and its step semantics:
  * one train_step = a D run on one batch, then a G run on ANOTHER batch (each `sess.run` pulled a new batch);
  * BN always uses batch statistics when is_training (also in test_step); moving averages are updated only in
    the G run; the shared pose_encoder normalises each of its two calls separately;
  * loss_D = BCE(D(real),1) + BCE(D(fake),0); loss_G = perceptual(VGG19, 5 taps, L1) + BCE(D(fake),1);
  * lr = start * decay^(global_step/step) (continuous), two Adam(beta1=.5, beta2=.999, eps=1e-8) optimisers,
    global_step += 1 in the G run.
Data parallelism: one process per GPU; gradients of the flat G / D buffers are summed with one NCCL all-reduce
each per step and scaled by 1/world_size inside the Adam kernel; BN statistics stay per replica.
"""
import os
import time
from datetime import datetime

import torch

from .. import dp
from .. import engine as E
from .. import networks
from .. import ops
from ..utils import model as model_utils
from .base_model import BaseModel, GlobalStep, log


class DetectorTranslatorModel(BaseModel):
    name = 'detector_translator'

    def __init__(self, config, global_step=None, is_training=True, device=None, seed=0, process_group=None):
        super().__init__(is_training)
        train_config = config['training']
        model_config = config['model']
        paths_config = config['paths']
        self.lr = train_config['lr'] if self.is_training else None
        self.batch_size = train_config['batch_size']
        self.n_points = model_config['n_pts']
        self.log_dir = paths_config['log_dir']
        self.vgg19_path = paths_config.get('vggnet')
        self.colors = model_utils.get_n_colors(model_config['n_pts'], pastel_factor=0.0)
        self.global_step = global_step if global_step is not None else GlobalStep(0)
        self.device = torch.device(device if device is not None else "cuda")
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.inputs = None
        self.t_D = 0
        self.t_G = 0
        self._graph = None
        self._lr_dev = None
        self._static = None
        self._side = None                  # second stream: the D run beside the G run's forward / perceptual chain
        self._side_branch = None
        self._d_pending = False
        self.overlap_d_update = os.environ.get("KP_OVERLAP_D_UPDATE", "1") != "0"
        # the two pose_encoder calls and VGG(gt) / VGG(pred) run as single batched passes (KP_BATCH_SHARED=0: separate calls)
        self.batch_shared_passes = os.environ.get("KP_BATCH_SHARED", "1") != "0"
        # data parallel: all-reduce the generator gradients in buckets (translator / pose_encoder / image_encoder, the
        # order in which the backward pass completes them) on the side stream, under the rest of the backward pass
        self.overlap_g_allreduce = os.environ.get("KP_OVERLAP_G_ALLREDUCE", "1") != "0"
        self._comm = None                  # communication stream of the bucketed gradient all-reduces
        self._done = {"G": [], "D": []}      # gradient slices already handed to an all-reduce in this backward pass
        # weight gradients on their own stream beside the data-gradient chain (engine.Context.wgrad_stream)
        self.overlap_wgrad = os.environ.get("KP_WGRAD_STREAM", "1") != "0"
        # outputs of the last forward pass (names follow the reference's attributes)
        self.final_output = self.crude_output = self.mask = None
        self.current_keypoints = self.future_keypoints = None
        self.loss_D = self.loss_G = None
        self.loss_D_real = self.loss_D_fake = self.loss_G_recon = self.loss_G_adv = None

        self.ctx = E.Context(self.device, n_pts=self.n_points)
        if self.overlap_wgrad and is_training and self.device.type == "cuda":
            self.ctx.wgrad_stream = torch.cuda.Stream(device=self.device)
        if os.environ.get("KP_BRANCH_STREAM", "1") != "0" and self.device.type == "cuda":
            self.ctx.branch_stream = torch.cuda.Stream(device=self.device, priority=int(os.environ.get("KP_BRANCH_PRIORITY", "0")))
        networks.build_parameters(self.ctx, self.n_points, with_vgg=True)
        self._init_weights(seed)
        if self.world > 1:
            dp.broadcast_parameters(self.ctx, 0, self.pg)      # replicas start (and stay) identical

    # ------------------------------------------------------------------------------------------
    def _init_weights(self, seed):
        """xavier-uniform kernels, zero biases (tf.contrib.layers.xavier_initializer, layers.py:8); VGG19 weights
        come from vgg19.npy when the file exists, otherwise He-normal random (benchmarks: there is no network)."""
        import math
        import os
        g = torch.Generator(device="cpu").manual_seed(seed)
        ctx = self.ctx
        for grp in (ctx.G, ctx.D):
            for n in grp.names():
                if n.endswith("/kernel"):
                    p = grp.p(n)
                    k1, k2, cin, cout = p.shape
                    lim = math.sqrt(6.0 / (k1 * k2 * (cin + cout)))
                    p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * lim).to(self.device))
        if self.vgg19_path and os.path.exists(self.vgg19_path):
            networks.vgg.load_npy_into(ctx, self.vgg19_path)
        else:
            for n in ctx.V.names():
                p = ctx.V.p(n)
                if n.endswith("/filter"):
                    p.copy_((torch.randn(p.shape, generator=g) * math.sqrt(2.0 / (9 * p.shape[2]))).to(self.device))
                else:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(self.device))
        ctx.params_changed()

    def _extra_state(self):
        return {"global_step": int(self.global_step.value), "t_D": self.t_D, "t_G": self.t_G,
                "adam": {"G_m": self.ctx.G.m.cpu(), "G_v": self.ctx.G.v.cpu(), "D_m": self.ctx.D.m.cpu(),
                         "D_v": self.ctx.D.v.cpu()}}

    def _load_extra_state(self, ex):
        self.global_step.value = int(ex.get("global_step", 0))
        self.t_D, self.t_G = int(ex.get("t_D", 0)), int(ex.get("t_G", 0))
        if "adam" in ex:
            self.ctx.G.m.copy_(ex["adam"]["G_m"]); self.ctx.G.v.copy_(ex["adam"]["G_v"])
            self.ctx.D.m.copy_(ex["adam"]["D_m"]); self.ctx.D.v.copy_(ex["adam"]["D_v"])

    # ------------------------------------------------------------------------------------------
    def build(self, inputs):
        """inputs: dict with 'image' and 'future_image' (float32 NHWC CUDA tensors in [-1,1]), or a callable
        returning such a dict — called once per run, which reproduces the reference's "every sess.run pulls a
        new batch" behaviour (train.py:46-50)."""
        self.inputs = inputs
        networks.set_context(self.ctx)

    def _next_batch(self, feed_dict=None):
        src = feed_dict if isinstance(feed_dict, dict) and 'image' in feed_dict else self.inputs
        b = src() if callable(src) else src
        return b['image'], b['future_image']

    def _define_forward_pass(self, im, future_im, for_G_run=True):
        """reference :160-184.  Returns the heads tensor too (crude+mask before compose)."""
        networks.set_context(self.ctx)
        tm = self.is_training
        # image_encoder and pose_encoder are independent until the translator: the shorter one runs on the branch stream
        with self.ctx.branch():
            embeddings = networks.image_encoder(im, tm) if for_G_run else \
                [im] + networks.encoder(networks._prep7(im), tm, _scope="image_encoder/encoder/", _n_blocks=3) + [None]
        self._grad_bucket_marker("pose_encoder/")        # runs (in the backward pass) after BOTH pose_encoder calls are done
        if self.batch_shared_passes:
            # pose_encoder(im) and pose_encoder(future_im) as one pass over [im; future_im] with per-call BN statistics
            (current_gauss_pt, current_pt_map), (future_gauss_pt, future_pt_map) = \
                networks.pose_encoder_pair_with_maps(im, future_im, self.n_points, tm, (32, 32))
        else:
            current_gauss_pt, current_pt_map = networks.pose_encoder_with_maps(im, self.n_points, tm, (32, 32))
            future_gauss_pt, future_pt_map = networks.pose_encoder_with_maps(future_im, self.n_points, tm, (32, 32))
        self.ctx.branch_join()
        joint_embedding = networks.joint_embedding(embeddings[-2], current_pt_map, future_pt_map)
        self._grad_bucket_marker("translator/")          # ... after the translator's last weight gradient
        heads = networks.translator_heads(joint_embedding, tm)
        final_output = networks.compose(im, heads)
        self.final_output = final_output
        self.crude_output, self.mask = heads[..., :3], heads[..., 3:4]
        self.current_keypoints, self.future_keypoints = current_gauss_pt, future_gauss_pt
        return final_output

    @property
    def current_keypoints_map(self):
        return model_utils.get_gaussian_maps(self.current_keypoints, [128, 128])

    @property
    def future_keypoints_map(self):
        return model_utils.get_gaussian_maps(self.future_keypoints, [128, 128])

    def _current_lr(self):
        return self.lr['start_val'] * self.lr['decay'] ** (float(self.global_step.value) / self.lr['step'])

    # ---- losses (reference :246-289) ----
    def _compute_loss_D(self, future_im_pred, future_im, backward):
        ctx = self.ctx
        B = future_im.shape[0]
        x = torch.cat([future_im, future_im_pred.detach()], dim=0)
        logits = networks.img_discr(x, marker=self._d_marker if backward else None)    # [2B,6,6,1]: real half then fake half
        loss = torch.zeros(2, device=self.device)
        d_real = ops.bce_logits(logits[:B], 1.0, 1.0, loss[0:1], backward)
        d_fake = ops.bce_logits(logits[B:], 0.0, 1.0, loss[1:2], backward)
        if backward:
            ctx.tape.set_grad(logits, torch.cat([d_real, d_fake], dim=0))
        return loss

    def _compute_loss_G(self, future_im_pred, future_im, backward):
        ctx = self.ctx
        loss = torch.zeros(2, device=self.device)             # [reconstruction, adversarial]
        tape = ctx.tape
        # The adversarial chain D(fake) and the perceptual chain VGG([gt; pred]) are independent, forward and backward,
        # and meet only in the gradient of the generated frame: D(fake) runs on the branch stream (issued FIRST, so that in
        # the backward pass its closures come last and overlap the VGG backward).  It reads the frame through an alias with
        # its own gradient buffer; the two gradients are added after the branch has joined.
        pred_D = future_im_pred.view(future_im_pred.shape)
        if backward:
            def merge_bwd():
                g = tape.grad(pred_D)
                if g is None:
                    return
                dx, acc = tape.acquire(future_im_pred, dtype=g.dtype)
                if acc:
                    dx.add_(g)
                else:
                    dx.copy_(g)
            tape.record(merge_bwd)
        with ctx.branch():
            self._join_D()                      # the discriminator update of the D run must have landed
            fake_ = networks.img_discr(pred_D, need_input_grad=backward)
            d_adv = ops.bce_logits(fake_, 1.0, 1.0, loss[1:2], backward)
            if backward:
                tape.set_grad(fake_, d_adv)
        if self.batch_shared_passes:
            # VGG on [gt; pred] in one pass like the reference (:274-279); only the generated half carries a gradient
            B = future_im.shape[0]
            u = networks.vgg.VGG_UNROLL
            xp = torch.empty((2 * B, 128, 128, u[2]), device=self.device, dtype=torch.bfloat16)
            networks.vgg.prepare(future_im, ops.VGG_PREP, out=xp[:B])
            networks.vgg.prepare(future_im_pred, ops.VGG_PREP, out=xp[B:])
            if backward:
                def prep_bwd():
                    g = tape.grad(xp)
                    if g is None:
                        return
                    dx, acc = tape.acquire(future_im_pred)
                    ops.image_prep_unrolled_bwd(g[B:], dx, u[0], u[1], ops.VGG_PREP, accumulate=acc)
                tape.record(prep_bwd)
            feats = networks.vgg.features_from_prepared(xp, need_input_grad=backward, grad_rows=(B, 2 * B))
            for f in feats:
                d = torch.empty_like(f) if backward else None
                ops.l1_pair(f[:B], f[B:], 1.0 / len(feats), loss[0:1], d[B:] if backward else None)
                if backward:
                    tape.set_grad(f, d)
        else:
            ctx.tape = None
            feat_gt = networks.vgg.features_from_prepared(networks.vgg.prepare(future_im, ops.VGG_PREP), need_input_grad=False)
            ctx.tape = tape
            xp = networks._prep_with_grad(ctx, future_im_pred, ops.VGG_PREP, networks.vgg.VGG_UNROLL) if backward else \
                networks.vgg.prepare(future_im_pred, ops.VGG_PREP)
            feat_pred = networks.vgg.features_from_prepared(xp, need_input_grad=backward)
            for fg, fp in zip(feat_gt, feat_pred):
                d = torch.empty_like(fp) if backward else None
                ops.l1_pair(fg, fp, 1.0 / len(feat_pred), loss[0:1], d)
                if backward:
                    tape.set_grad(fp, d)
        ctx.branch_join()
        return loss

    def _allreduce(self, buf):
        if self.world > 1:
            dp.allreduce_sum_(buf, self.pg)

    def _bucket_range(self, prefix):
        """[lo, hi) of the flat generator buffer holding the variables whose names start with `prefix` (contiguous: the
        variables are declared network by network, networks.build_parameters)."""
        G = self.ctx.G
        offs = [(off, off + (int(torch.tensor(shape).prod()) + 3) // 4 * 4) for name, shape, off in G.specs if name.startswith(prefix)]
        lo, hi = min(o[0] for o in offs), max(o[1] for o in offs)
        inside = sum(1 for name, _, off in G.specs if lo <= off < hi)
        assert inside == len(offs), "variables of %r are not contiguous in the flat buffer" % prefix
        return lo, hi

    def _comm_stream(self):
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.device, priority=-1)
        return self._comm

    def _allreduce_marker(self, which, lo, hi):
        """Tape entry recorded BEFORE the forward of the layers that own the flat gradient slice [lo, hi): in the backward pass
        it runs right after their last weight gradient has been issued and starts the all-reduce of that bucket on the
        communication stream while the backward pass of the layers further upstream goes on (data parallel only; the rest
        and the join are in _finish_allreduce, before Adam)."""
        ctx = self.ctx
        grad = ctx.G.grad if which == "G" else ctx.D.grad
        self._done[which].append((lo, hi))

        def fire():
            c = self._comm_stream()
            c.wait_stream(torch.cuda.current_stream())
            ctx.join_wgrad(waiter=c)                 # the bucket's weight gradients may still run on their own stream
            with torch.cuda.stream(c):
                self._allreduce(grad[lo:hi])
        ctx.tape.record(fire)

    def _grad_bucket_marker(self, prefix):
        ctx = self.ctx
        if ctx.tape is None or self.world <= 1 or not self.overlap_g_allreduce or not ctx.train_G:
            return
        lo, hi = self._bucket_range(prefix)
        self._allreduce_marker("G", lo, hi)

    def _d_marker(self, scope):
        """img_discr calls this before each layer.  Buckets in the order the backward pass completes them: [conv_5, D_logit]
        (3/4 of the 179 MB) right after the first two layers of the backward pass, then [conv_3, conv_4]; conv_0..2 are
        the rest."""
        ctx = self.ctx
        if ctx.tape is None or self.world <= 1 or not self.overlap_g_allreduce or not ctx.train_D:
            return
        D = ctx.D
        off = lambda pre: min(o for name, _, o in D.specs if name.startswith(pre))
        if scope == "img_discr/conv_5/":
            self._allreduce_marker("D", off(scope), D.total)
        elif scope == "img_discr/conv_3/":
            self._allreduce_marker("D", off(scope), off("img_discr/conv_5/"))

    def _finish_allreduce(self, which):
        """After the backward pass: all-reduce what the markers have not sent, then wait for the communication stream."""
        if self.world <= 1:
            return
        grad = self.ctx.G.grad if which == "G" else self.ctx.D.grad
        done, self._done[which] = sorted(self._done[which]), []
        pos = 0
        for lo, hi in done + [(grad.numel(), grad.numel())]:
            if lo > pos:
                self._allreduce(grad[pos:lo])
            pos = max(pos, hi)
        if done:
            torch.cuda.current_stream().wait_stream(self._comm_stream())

    # ---- the two runs of one train step ----
    def _run_D(self, im, future_im):
        if self.overlap_d_update and self.device.type == "cuda":
            # The D run (generator forward without gradient on ITS batch, img_discr forward/backward, gradient all-reduce,
            # Adam(D), weight re-pack) shares nothing with the G run's generator forward and perceptual chain - they read
            # the same generator weights and meet only at the G run's D(fake), which needs the updated discriminator
            # (_join_D).  So the whole run goes to a second stream and the two chains fill each other's launch tails and
            # pair HBM-bound with tensor-bound kernels.  Inside the captured graph this is a fork/join of nodes.
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device, priority=int(os.environ.get("KP_DCHAIN_PRIORITY", "-1")))
            if self._side_branch is None:
                self._side_branch = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(main)
            keep = self.ctx.branch_stream
            if keep is not None:
                self.ctx.branch_stream = self._side_branch
            with torch.cuda.stream(self._side):
                loss = self._run_D_body(im, future_im)
            self.ctx.branch_stream = keep
            self._d_pending = True
            return loss
        return self._run_D_body(im, future_im)

    def _run_D_body(self, im, future_im):
        ctx = self.ctx
        ctx.begin_run(0)
        ctx.tape, ctx.update_moving, ctx.train_G, ctx.train_D = None, False, False, False
        fake = self._define_forward_pass(im, future_im, for_G_run=False)
        ctx.tape, ctx.train_D = E.Tape(), True
        ctx.D.grad.zero_()
        self._done["D"] = []
        loss = self._compute_loss_D(fake, future_im, backward=True)
        ctx.tape.backward()
        ctx.tape, ctx.train_D = None, False
        self._finish_D()
        return loss

    def _finish_D(self):
        ctx = self.ctx
        self._finish_allreduce("D")
        self.t_D += 1
        ops.adam_tf(ctx.D.data, ctx.D.grad, ctx.D.m, ctx.D.v, self._current_lr(), self.t_D, grad_scale=1.0 / self.world,
                    lr_t_dev=self._lr_dev[0:1] if self._lr_dev is not None else None)
        ctx.params_changed(ctx.D)

    def _join_D(self):
        if self._d_pending:
            torch.cuda.current_stream().wait_stream(self._side)
            self._d_pending = False

    def _run_G(self, im, future_im):
        ctx = self.ctx
        ctx.begin_run(1)                   # its own scratch pool: the D run may still be running on the other stream
        ctx.tape, ctx.update_moving, ctx.train_G, ctx.train_D = E.Tape(), True, True, False
        ctx.G.grad.zero_()
        self._done["G"] = []
        fake = self._define_forward_pass(im, future_im, for_G_run=True)
        loss = self._compute_loss_G(fake, future_im, backward=True)
        ctx.tape.backward()
        ctx.tape, ctx.update_moving, ctx.train_G = None, False, False
        self._join_D()
        self._finish_allreduce("G")
        self.t_G += 1
        ops.adam_tf(ctx.G.data, ctx.G.grad, ctx.G.m, ctx.G.v, self._current_lr(), self.t_G, grad_scale=1.0 / self.world,
                    lr_t_dev=self._lr_dev[1:2] if self._lr_dev is not None else None)
        self.global_step.value += 1
        ctx.params_changed(ctx.G)
        return loss

    # ---- CUDA-graph execution of the whole train step (the ~2000 launches of a step are CPU-launch bound) ----
    def enable_cuda_graph(self, batch_size):
        """Capture D run + G run (incl. gradient all-reduces and Adam) into one CUDA graph on static input buffers.
        Subsequent train_step calls copy the two batches into those buffers and replay the graph; the step sizes
        lr_t of both optimisers are fed through a device scalar."""
        dev = self.device
        self._static = [torch.empty((batch_size, 128, 128, 3), device=dev) for _ in range(4)]
        self._lr_dev = torch.zeros(2, device=dev)
        self._set_lr_dev()
        # warm-up (eager) on a side stream: first-call kernel attributes, plan caches, NCCL communicators
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for b in self._static:
                b.uniform_(-1, 1)
            snap = (self.ctx.G.data.clone(), self.ctx.D.data.clone(), self.ctx.S.data.clone(), self.t_D, self.t_G,
                    self.global_step.value, self.ctx.G.m.clone(), self.ctx.G.v.clone(), self.ctx.D.m.clone(),
                    self.ctx.D.v.clone())
            for _ in range(2):              # step 1 registers every weight-pack job, step 2 builds the final job tables
                self._run_D(self._static[0], self._static[1])
                self._run_G(self._static[2], self._static[3])
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from .. import _lib
        n0 = _lib.load().kp_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        # the main chain is captured from a high-priority stream: where a side stream's kernel (weight gradients, the D
        # update) and the next kernel of the chain are both ready, the chain gets the SMs first
        prio = int(os.environ.get("KP_GRAPH_PRIORITY", "-1"))
        cap = torch.cuda.Stream(device=dev, priority=prio)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(self._graph, stream=cap):
            lD = self._run_D(self._static[0], self._static[1])
            lG = self._run_G(self._static[2], self._static[3])
        self.graph_launches = int(_lib.load().kp_launch_count() - n0)   # library kernels recorded per replay
        self._graph_losses = (lD, lG)
        # undo the warm-up/capture steps' host-side bookkeeping and parameter updates
        self.ctx.G.data.copy_(snap[0]); self.ctx.D.data.copy_(snap[1]); self.ctx.S.data.copy_(snap[2])
        self.ctx.G.m.copy_(snap[6]); self.ctx.G.v.copy_(snap[7]); self.ctx.D.m.copy_(snap[8]); self.ctx.D.v.copy_(snap[9])
        self.t_D, self.t_G, self.global_step.value = snap[3], snap[4], snap[5]
        self.ctx.params_changed()
        torch.cuda.synchronize()

    def _set_lr_dev(self):
        lr = self._current_lr()
        vals = torch.tensor([ops.adam_lr_t(lr, self.t_D + 1), ops.adam_lr_t(lr, self.t_G + 1)], dtype=torch.float32)
        self._lr_dev.copy_(vals, non_blocking=True)

    def train_step(self, sess=None, feed_dict=None, step=0, batch_size=None, should_write_log=False,
                   should_write_summary=False):
        start_time = time.time()
        if not self.is_training:
            raise RuntimeError("train_step on a model built with is_training=False (batch norm is folded: no gradients)")
        if self._graph is not None:
            im, fut = self._next_batch(feed_dict)
            self._static[0].copy_(im, non_blocking=True); self._static[1].copy_(fut, non_blocking=True)
            im, fut = self._next_batch(feed_dict)
            self._static[2].copy_(im, non_blocking=True); self._static[3].copy_(fut, non_blocking=True)
            self._set_lr_dev()
            self._graph.replay()
            self.t_D += 1; self.t_G += 1; self.global_step.value += 1
            loss_D, loss_G = self._graph_losses
        else:
            im, fut = self._next_batch(feed_dict)
            loss_D = self._run_D(im, fut)
            im, fut = self._next_batch(feed_dict)
            loss_G = self._run_G(im, fut)
        self._last_losses = (loss_D, loss_G)
        if should_write_log:
            ld, lg = float(loss_D.sum().item()), float(loss_G.sum().item())
            duration = time.time() - start_time
            bs = batch_size or im.shape[0]
            log_format = '%s: step %d, loss_D = %.4f, loss_G = %.4f (%.1f examples/sec) %.3f sec/batch'
            log.info(log_format % (datetime.now(), step, ld, lg, bs / float(duration), duration))
            self.loss_D, self.loss_G = ld, lg

    def test_step(self, sess=None, feed_dict=None, step=0, test_idx=0, batch_size=None):
        ctx = self.ctx
        start_time = time.time()
        im, fut = self._next_batch(feed_dict)
        ctx.begin_run()
        ctx.tape, ctx.update_moving, ctx.train_G, ctx.train_D = None, False, False, False
        fake = self._define_forward_pass(im, fut, for_G_run=False)
        lD = self._compute_loss_D(fake, fut, backward=False)
        lG = self._compute_loss_G(fake, fut, backward=False)
        vals = torch.cat([lD, lG]).tolist()
        duration = time.time() - start_time
        self.loss_D_real, self.loss_D_fake, self.loss_G_recon, self.loss_G_adv = vals
        return vals[0] + vals[1], vals[2] + vals[3], duration, batch_size or im.shape[0]

    def collect_test_results(self, results, step):
        average_loss_D = sum(x[0] for x in results) / len(results)
        average_loss_G = sum(x[1] for x in results) / len(results)
        total_duration = sum(x[2] for x in results)
        average_duration = total_duration / len(results)
        num_examples = sum(x[3] for x in results)
        log_format = 'test: %s: step %d, loss_D = %.4f, loss_G = %.4f (%.1f examples/sec) %.3f sec/batch'
        log.info(log_format % (datetime.now(), step, average_loss_D, average_loss_G, num_examples / total_duration,
                               average_duration))
        return average_loss_D, average_loss_G
