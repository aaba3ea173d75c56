"""Pseudo-label writer: the on-disk hand-over from stage 1 to stage 2 of the reference.

Reference: make_pseudo_labels.py:83-101 runs KeypointModel over every video (one zero-padded [1,663,128,128,3] clip per
`sess.run`, data/keypoint_dataloader.py:77-80) and `_save_output` writes `pseudo_labels/{idx:04d}.npy` =
`outputs['pts'][0, :len]`, float32 [len, n_pts, 2], (x, y) in [-1, 1]; data/sequence_dataloader.py:101 `np.load`s that file
and indexes it by frame.  Here:
  * the video list is sharded contiguously over the ranks (dp.shard_range) - frames are independent, no collective;
  * only the `len` real frames of a clip go through the detector (the reference computes the zero padding and throws it
    away; with inference-mode batch norm every frame is independent, so the kept rows are the same);
  * the result is copied to pinned host memory asynchronously and written by a background thread, so the device does not
    wait for the file system.
"""
import os
import queue
import threading

import numpy as np
import torch

from . import dp


class PseudoLabelWriter:
    """Background `.npy` writer.  put(idx, pts_device, event) enqueues; close() drains."""

    def __init__(self, out_dir, depth=8):
        self.out_dir = out_dir
        os.makedirs(out_dir, exist_ok=True)
        self.q = queue.Queue(maxsize=depth)
        self.error = None
        self.written = []
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    @staticmethod
    def path_for(out_dir, idx):
        return os.path.join(out_dir, '{:04d}.npy'.format(int(idx)))

    def _loop(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            idx, host, event = item
            try:
                if event is not None:
                    event.synchronize()
                arr = np.ascontiguousarray(host.numpy(), dtype=np.float32)
                path = self.path_for(self.out_dir, idx)
                np.save(path, arr)
                self.written.append(path)
            except Exception as e:   # surfaced by close()
                self.error = e

    def put(self, idx, pts):
        """pts: [len, n_pts, 2] tensor (device or host)."""
        event = None
        if pts.is_cuda:
            host = torch.empty(tuple(pts.shape), dtype=torch.float32, pin_memory=True)
            host.copy_(pts, non_blocking=True)
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(pts.device))
        else:
            host = pts.to(torch.float32)
        self.q.put((int(idx), host, event))

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error
        return sorted(self.written)


def write_pseudo_labels(detect, videos, out_dir, rank=None, world=None):
    """Run `detect` (KeypointModel.detect: frames [F,128,128,3] -> keypoints [F,n_pts,2]) over this rank's shard of `videos`
    and write one `{idx:04d}.npy` per video.

    videos: sequence of dicts {'image': [T,128,128,3] or [1,T,128,128,3] float tensor in [-1,1], 'idx': int, 'len': int}
            (the reference's keypoint_dataloader.py:33-38 element: 'len' real frames, the rest zero padding) or a callable
            i -> such a dict plus `n` via len(); or a `data.KeypointDataLoader` (make_pseudo_labels.py:55-66): this rank's shard
            of its video list is then decoded and cropped by the loader's device pipeline, one video per batch.
    Returns the list of files this rank wrote."""
    rank = dp.rank() if rank is None else rank
    world = dp.world_size() if world is None else world
    from_loader = hasattr(videos, "get_dataset")
    lo, hi = dp.shard_range(videos.length() if from_loader else len(videos), rank, world)
    writer = PseudoLabelWriter(out_dir)
    dataset = videos.get_dataset(1, sample_range=(lo, hi)) if from_loader and hi > lo else None
    try:
        for v in (dataset if from_loader else (videos[i] for i in range(lo, hi))) or ():
            im = v['image']
            if im.dim() == 5:
                im = im[0]
            n = int(v['len']) if not torch.is_tensor(v['len']) else int(v['len'].reshape(-1)[0])
            idx = int(v['idx']) if not torch.is_tensor(v['idx']) else int(v['idx'].reshape(-1)[0])
            pts = detect(im[:n].contiguous())
            writer.put(idx, pts)
    finally:
        files = writer.close()
        if dataset is not None:
            dataset.close()
    return files
