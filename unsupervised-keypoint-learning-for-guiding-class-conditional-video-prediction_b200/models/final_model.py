"""FinalModel — evaluation renderer, translator half of /root/reference/models/final_model.py:49-122.

The stage-2 motion generator (`networks.vae_decoder`, final_model.py:71-77) is outside this path: the future
keypoint trajectory is taken from inputs['pred_seq'] [B,32,n_pts,2] (or produced by a user-supplied
`motion_fn(first_pt, action_code)`).  Everything else follows the reference: first frame tiled x32, image
embedding tiled x32, 32x32 maps of the first and the predicted keypoints, translator in inference mode, mask
compose with clipping, colourised 128x128 keypoint visualisations; the output dict has the reference's keys.
"""
import torch

from .. import engine as E
from .. import networks
from .. import ops
from ..utils import model as model_utils
from .base_model import BaseModel

N_FUTURE_FRAMES = 32
IMAGE_SIZE = 128


class FinalModel(BaseModel):
    name = 'final'

    def __init__(self, config, device=None, motion_fn=None, frame_chunk=1024):
        super().__init__(False)
        model_config = config['model']
        paths_config = config['paths']
        self.log_dir = paths_config['log_dir']
        self.n_points = model_config['n_pts']
        self.cell_info = model_config.get('cell_info')
        self.vae_dim = model_config.get('vae_dim')
        self.colors = model_utils.get_n_colors(model_config['n_pts'], pastel_factor=0.0)
        self.device = torch.device(device if device is not None else "cuda")
        self.motion_fn = motion_fn
        self.frame_chunk = frame_chunk
        self.ctx = E.Context(self.device, n_pts=self.n_points)
        networks.build_parameters(self.ctx, self.n_points, with_vgg=False)

    def build(self, inputs):
        self.inputs = inputs
        networks.set_context(self.ctx)

    def run(self, sess=None, feed_dict=None, visualize=True):
        networks.set_context(self.ctx)
        self.ctx.tape = None
        b = self.inputs() if callable(self.inputs) else self.inputs
        im = b['image'].contiguous()
        B = im.shape[0]
        T = N_FUTURE_FRAMES
        K = self.n_points

        embeddings = networks.image_encoder(im, False)[-2]                       # [B,32,32,128]
        first_pt, cur_map = networks.pose_encoder_with_maps(im, K, False, (32, 32))
        if 'pred_seq' in b and b['pred_seq'] is not None:
            pred_seq = b['pred_seq']
        elif self.motion_fn is not None:
            pred_seq = self.motion_fn(first_pt.reshape(B, K * 2), b.get('action_code'))
        else:
            raise NotImplementedError("stage-2 vae_decoder is outside the stage-1 path: supply inputs['pred_seq'] "
                                      "[B,32,n_pts,2] or a motion_fn")
        pred_seq = pred_seq.reshape(B, T, K, 2).to(torch.float32).contiguous()
        pred_map = model_utils.get_gaussian_maps(pred_seq.reshape(B * T, K, 2), [32, 32])     # [BT,32,32,K]

        finals, crudes, masks = [], [], []
        vids_per_chunk = max(1, self.frame_chunk // T)
        for v0 in range(0, B, vids_per_chunk):
            v1 = min(B, v0 + vids_per_chunk)
            n = (v1 - v0) * T
            emb_t = embeddings[v0:v1].unsqueeze(1).expand(-1, T, -1, -1, -1).reshape(n, *embeddings.shape[1:])
            cur_t = cur_map[v0:v1].unsqueeze(1).expand(-1, T, -1, -1, -1).reshape(n, *cur_map.shape[1:])
            im_t = im[v0:v1].unsqueeze(1).expand(-1, T, -1, -1, -1).reshape(n, IMAGE_SIZE, IMAGE_SIZE, 3).contiguous()
            joint = networks.joint_embedding(emb_t, cur_t, pred_map[v0 * T:v1 * T])
            heads = networks.translator_heads(joint, False)
            final, crude, mask = networks.compose(im_t, heads, clip=True, want_parts=True)
            finals.append(final); crudes.append(crude); masks.append(mask)
        final = torch.cat(finals) if len(finals) > 1 else finals[0]
        crude = torch.cat(crudes) if len(crudes) > 1 else crudes[0]
        mask = torch.cat(masks) if len(masks) > 1 else masks[0]

        out = {
            'real_im_seq': b.get('real_im_seq'),
            'im': im,
            'pred_im_seq': final.reshape(B, T, IMAGE_SIZE, IMAGE_SIZE, 3),
            'mask': mask.reshape(B, T, IMAGE_SIZE, IMAGE_SIZE, 1),
            'pred_im_crude': crude.reshape(B, T, IMAGE_SIZE, IMAGE_SIZE, 3),
            'fut_pt_raw': pred_seq,
        }
        if visualize:
            # get_gaussian_maps(.,[128,128]) + colorize_point_maps (final_model.py:102-109) fused: no [.,128,128,40] maps
            out['current_points'] = model_utils.gaussian_maps_colorized(first_pt, self.colors, [IMAGE_SIZE, IMAGE_SIZE])
            out['future_points'] = model_utils.gaussian_maps_colorized(
                pred_seq.reshape(B * T, K, 2), self.colors, [IMAGE_SIZE, IMAGE_SIZE]).reshape(B, T, IMAGE_SIZE, IMAGE_SIZE, 3)
        return out

    def train_step(self, sess, feed_dict, step, batch_size, should_write_log=False, should_write_summary=False):
        """This model is not trainable"""
        raise NotImplementedError

    def test_step(self, sess, feed_dict, step, test_idx, batch_size):
        """This model has no test step"""
        raise NotImplementedError

    def collect_test_results(self, results, step):
        """This model has no test step"""
        raise NotImplementedError
