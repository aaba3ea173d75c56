"""Mirror of the reference's ``utils`` package (utils/__init__.py:8-19, utils/training.py:4-5)."""
import math
import os

import yaml


def load_config(config_path):
    with open(config_path) as fh:
        return yaml.load(fh, Loader=yaml.FullLoader)


def touch_dir(path):
    os.makedirs(path, exist_ok=True)


def get_n_iterations(total, batch_size):
    return int(math.ceil(total / float(batch_size)))


class DevicePrefetcher:
    """Input pipeline helper: wraps a callable that returns a dict of PINNED host tensors and keeps `depth` batches in
    flight to the device on a side stream, so that the host-to-device copy of the next batches overlaps the current
    train step (the reference gets the same overlap from tf.data's prefetch(1), train.py:36-43).  Calling the object
    returns the oldest batch (device tensors) after making the current stream wait for its copy."""

    def __init__(self, host_feed, device, depth=2):
        import collections
        import torch
        self._torch = torch
        self.feed, self.device = host_feed, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.queue = collections.deque()
        for _ in range(max(1, depth)):
            self._launch()

    def _launch(self):
        torch = self._torch
        host = self.feed()
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((dev, ev))

    def __call__(self):
        torch = self._torch
        dev, ev = self.queue.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev.values():
            t.record_stream(cur)
        self._launch()
        return dev
