"""KeypointModel — pseudo-label extractor, mirror of /root/reference/models/keypoint_model.py:12-83.

Detector only, inference-mode BN (folded into the convolutions), frames are independent: no collective.
`run()` returns the reference's dict {'pts': [B,T,n_pts,2], 'idx', 'len', 'im'} (T = 663 in the reference,
:52; here the T of the input).
"""
import torch

from .. import engine as E
from .. import networks
from ..utils import model as model_utils
from .base_model import BaseModel


class KeypointModel(BaseModel):
    name = 'stage1'

    def __init__(self, config, device=None, chunk=512):
        super().__init__(False)
        model_config = config['model']
        paths_config = config['paths']
        self.n_points = model_config['n_pts']
        self.log_dir = paths_config['log_dir']
        self.colors = model_utils.get_n_colors(model_config['n_pts'], pastel_factor=0.0)
        self.device = torch.device(device if device is not None else "cuda")
        self.chunk = chunk
        self.input_im = self.input_idx = self.input_len = None
        self.ctx = E.Context(self.device, n_pts=self.n_points)
        networks.build_parameters(self.ctx, self.n_points, with_vgg=False)

    def build(self, inputs):
        self.inputs = inputs
        networks.set_context(self.ctx)

    def detect(self, frames):
        """frames: float32 [F,128,128,3] in [-1,1] -> keypoints [F,n_pts,2] (x,y)."""
        networks.set_context(self.ctx)
        self.ctx.tape = None
        outs = []
        for s in range(0, frames.shape[0], self.chunk):
            outs.append(networks.pose_encoder(frames[s:s + self.chunk].contiguous(), self.n_points, False))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def run(self, sess=None, feed_dict=None):
        b = self.inputs() if callable(self.inputs) else self.inputs
        im = b['image']
        T = im.shape[1]
        pts = self.detect(im.reshape(-1, im.shape[-3], im.shape[-2], im.shape[-1]))
        return {'pts': pts.reshape(-1, T, self.n_points, 2), 'idx': b.get('idx'), 'len': b.get('len'), 'im': im}

    def write_pseudo_labels(self, videos, out_dir, rank=None, world=None):
        """make_pseudo_labels.py:83-101: one `{idx:04d}.npy` ([len, n_pts, 2] float32) per video of this rank's shard."""
        from .. import pseudo_labels
        return pseudo_labels.write_pseudo_labels(self.detect, videos, out_dir, rank, world)

    def train_step(self, sess, feed_dict, step, batch_size, should_write_log=False, should_write_summary=False):
        """This model is not trainable"""
        raise NotImplementedError

    def test_step(self, sess, feed_dict, step, test_idx, batch_size):
        """This model has no test step"""
        raise NotImplementedError

    def collect_test_results(self, results, step):
        """This model has no test step"""
        raise NotImplementedError
