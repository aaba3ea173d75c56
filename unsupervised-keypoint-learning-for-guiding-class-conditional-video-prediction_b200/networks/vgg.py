"""VGG19 feature extractor — mirror of /root/reference/models/networks/vgg.py:7-61 on the B200 engine.

Constant weights (`tf.constant` in the reference: no weight gradients); 16 x (3x3 conv + bias + ReLU) as
tcgen05 tap-GEMMs with the bias+ReLU epilogue, 2x2 max-pools in between.  Returns
[conv1_2, conv2_2, conv3_4, conv4_4, conv5_4] as bf16 NHWC tensors.
"""
import numpy as np
import torch

from .. import engine as E
from .. import ops
from .. import tapconv as tc

VGG_LAYERS = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
              ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256),
              ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512),
              ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512)]
_ORDER = ["conv1_1", "conv1_2", "pool", "conv2_1", "conv2_2", "pool", "conv3_1", "conv3_2", "conv3_3", "conv3_4", "pool",
          "conv4_1", "conv4_2", "conv4_3", "conv4_4", "pool", "conv5_1", "conv5_2", "conv5_3", "conv5_4"]
_TAPS = ("conv1_2", "conv2_2", "conv3_4", "conv4_4", "conv5_4")


def _ctx():
    from . import get_context
    return get_context()


def load_npy_into(ctx, vgg19_path):
    """Load the reference's vgg19.npy (dict name -> [W(3,3,Cin,Cout), b]) into the context's constant group."""
    data = np.load(vgg19_path, encoding='latin1', allow_pickle=True).item()
    sd = {}
    for name, _, _ in VGG_LAYERS:
        sd["vgg/%s/filter" % name] = data[name][0]
        sd["vgg/%s/biases" % name] = data[name][1]
    return ctx.load_state_dict(sd)


VGG_UNROLL = (3, 1, 16)   # conv1_1 as a 3x1 convolution over 3 horizontal taps x 3 channels (9 of 16 channels used)


def prepare(x, prep, out=None):
    """float32 image -> W-unrolled bf16 VGG input (ops.image_prep_unrolled with `prep` = affine map + channel order)."""
    return ops.image_prep_unrolled(x.contiguous(), VGG_UNROLL[0], VGG_UNROLL[1], VGG_UNROLL[2], prep, out=out)


def features_from_prepared(xp, need_input_grad, grad_rows=None):
    """xp: W-unrolled bf16 [N,H,W,16] in VGG input space (BGR minus mean), see prepare().  grad_rows=(lo, hi): only those
    images carry a gradient (the generated half of [gt; pred], reference :274-279)."""
    from . import maxpool
    ctx = _ctx()
    feats = []
    x = xp
    first = True
    for name in _ORDER:
        if name == "pool":
            x = maxpool(x, grad_rows)
            continue
        x = E.conv_layer(ctx, [x], "vgg/%s/filter" % name, "vgg/%s/biases" % name, (3, 1) if first else 3, 1, 0,
                         act=tc.ACT_RELU, need_input_grad=(need_input_grad or not first),
                         wshape=(3, 1, 9, 64) if first else None, grad_rows=grad_rows)
        first = False
        if name in _TAPS:
            feats.append(x)
    return feats


class Vgg19:
    def __init__(self, vgg19_path=None):
        """`vgg19_path`: the reference's vgg19.npy; None keeps whatever the context holds (random-init benchmarks)."""
        if vgg19_path is not None:
            load_npy_into(_ctx(), vgg19_path)

    def build(self, rgb):
        """rgb: float32 NHWC in [0,255] (reference vgg.py:13-43) -> the five feature maps."""
        prep = ((1.0, 1.0, 1.0), tuple(-m for m in ops.VGG_MEAN), (2, 1, 0))
        return features_from_prepared(prepare(rgb, prep), need_input_grad=False)
