"""BaseModel — mirror of /root/reference/models/base_model.py:8-92 without TensorFlow.

Same abstract interface (build / train_step / test_step / collect_test_results / initialize_loggers /
save_checkpoint / restore).  `sess` and `feed_dict` arguments are accepted for call-site compatibility and
ignored (there is no session: kernels are enqueued eagerly on the current CUDA stream).  Checkpoints are
name-keyed `.npz` files using the TF variable names incl. the Adam slots (checkpoint.py); `restore` loads only the
names present in both the file and the model, like the reference's filtered tf.train.Saver (:83-92).
"""
import logging
import os
from abc import ABC, abstractmethod
from os import path as osp

import torch

log = logging.getLogger("kp_b200")


class GlobalStep:
    """Stand-in for the reference's `tf.Variable(0, name='global_step')` (train.py:30)."""

    def __init__(self, value=0):
        self.value = int(value)

    def __int__(self):
        return self.value


class BaseModel(ABC):
    name = 'base_model'
    trainable = True

    def __init__(self, is_training=True):
        super().__init__()
        self.is_training = is_training
        self.log_dir = None
        self.train_writer = None
        self.test_writer = None
        self.saver = None
        self.ctx = None

    @abstractmethod
    def build(self, inputs):
        raise NotImplementedError

    @abstractmethod
    def train_step(self, sess, feed_dict, step, batch_size, should_write_log=False, should_write_summary=False):
        raise NotImplementedError

    @abstractmethod
    def test_step(self, sess, feed_dict, step, test_idx, batch_size):
        raise NotImplementedError

    @abstractmethod
    def collect_test_results(self, results, step):
        raise NotImplementedError

    def initialize_loggers(self, log_dir, sess=None):
        self.log_dir = log_dir
        os.makedirs(osp.join(log_dir, self.__class__.name), exist_ok=True)

    def save_checkpoint(self, sess, step):
        """reference :74-81 -> `<log_dir>/<name>/model.ckpt-<step>.npz` (TF variable names, see checkpoint.py).
        Data parallel: the batch-norm moving statistics are per replica (each normalises its own shard), so they are
        averaged over the replicas first and only rank 0 writes the file; every rank returns the path."""
        from .. import checkpoint, dp
        name = self.__class__.name
        path = osp.join(self.log_dir, name, 'model.ckpt-%d.npz' % int(step))
        world = dp.world_size(getattr(self, "pg", None))
        if world > 1 and self.ctx.S.data is not None and self.ctx.S.data.numel():
            dp.allreduce_sum_(self.ctx.S.data, getattr(self, "pg", None))
            self.ctx.S.data.div_(world)
            self.ctx.params_changed()
        if dp.rank(getattr(self, "pg", None)) == 0:
            os.makedirs(osp.dirname(path), exist_ok=True)
            checkpoint.save_npz(path, checkpoint.export_variables(self))
        return path

    def restore(self, sess, checkpoint_path):
        """reference :83-92: restore by NAME, only what both the file and the model have.  `.npz` (this package, or a dump of a
        TF checkpoint reader); `.pt` files of round 1 are read with weights_only=True."""
        from .. import checkpoint
        if str(checkpoint_path).endswith(".pt"):
            blob = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
            sd = blob["variables"] if "variables" in blob else blob
            loaded = self.ctx.load_state_dict(sd)
            if "extra" in blob and hasattr(self, "_load_extra_state"):
                self._load_extra_state(blob["extra"])
        else:
            loaded = checkpoint.import_variables(self, checkpoint.load_npz(checkpoint_path))
        print('vars-to-RESTORE:')
        print('\n'.join(loaded))
        return loaded
