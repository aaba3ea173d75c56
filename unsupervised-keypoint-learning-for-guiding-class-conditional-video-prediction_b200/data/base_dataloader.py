"""``BaseDataLoader`` (reference: data/base_dataloader.py:7-54) and the batch pipeline behind ``get_dataset``.

The reference builds ``tf.data.Dataset.from_generator(...).map(map_fn, 12 threads).batch(B).prefetch(1)``, its generator
doing ALL pixel work with Pillow on one Python thread.  Here ``sample_generator`` yields sample *descriptions* - which
files, and what the drawn rotation / crop / flip / filter are - and ``DeviceDataset`` turns ``batch_size`` of them into
device tensors:

    main (or prefetch) thread : pull descriptions (random draws in the reference's order, image sizes from JPEG headers)
    decode pool (12 threads)  : PIL decodes each JPEG straight into a pinned uint8 staging buffer (GIL released)
    side CUDA stream          : H2D of frames + plans, ONE kp_augment_frames launch -> float32 [n,128,128,3] in [-1,1]
    consumer                  : current stream waits on the batch's event

A frame request is a dict: ``path`` (or ``zero``), ``size`` (w, h of the JPEG), ``resize`` (W, H), ``crop`` (left, top),
``angle``, ``flip``, ``filter_id``, ``factor``.
"""
import collections
import concurrent.futures
import queue
import random
import threading
from abc import ABC, abstractmethod

import numpy as np
import torch
from PIL import Image

from .. import augment

IMAGE_SIZE = augment.IMAGE_SIZE


def frame_request(path, size, resize, crop, angle=0, flip=0):
    return {"path": path, "size": tuple(size), "resize": tuple(resize), "crop": tuple(crop), "angle": angle, "flip": flip,
            "filter_id": augment.NO_FILTER, "factor": 0.0}


def zero_frame():
    return {"zero": True}


def decode_rgb(path):
    """The decoded frame as uint8 [h, w, 3] (the reference hands PIL's decode to ``np.asarray`` as is)."""
    with Image.open(path) as im:
        return np.asarray(im if im.mode == "RGB" else im.convert("RGB"))


class _Staging:
    """One batch in flight: pinned frame bytes + plans, their device copies, the output tensor and the done event."""

    def __init__(self, device):
        self.device = device
        self.host = None
        self.dev = None
        self.event = None

    def reserve(self, nbytes):
        if self.host is None or self.host.numel() < nbytes:
            cap = max(int(nbytes * 1.25), 1 << 20)
            self.host = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            self.dev = torch.empty(cap, dtype=torch.uint8, device=self.device)
        return self.host


class DeviceDataset:
    """Iterable of batches ``{key: CUDA tensor}``; what ``BaseDataLoader.get_dataset`` returns."""

    def __init__(self, loader, batch_size, repeat, shuffle, num_preprocess_threads, prefetch, device, shuffle_buffer=2000):
        if not torch.cuda.is_available():
            raise RuntimeError("the input pipeline augments on the GPU; there is no CPU path")
        self.loader, self.batch_size, self.repeat, self.shuffle = loader, int(batch_size), repeat, shuffle
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.pool = concurrent.futures.ThreadPoolExecutor(max(1, int(num_preprocess_threads)))
        self.prefetch = 1 if prefetch else 0
        self.stream = torch.cuda.Stream(device=self.device)
        self.shuffle_buffer = shuffle_buffer
        self._shuffle_rng = random.Random(0x5eed)      # tf.data's shuffle has its own generator too (not reproducible)
        self.frames_done = 0

    # -- sample stream -------------------------------------------------------------------------------------------------
    def _samples(self):
        while True:
            yield from self.loader.sample_generator()
            if not self.repeat:
                return

    def _shuffled(self, it):
        buf = []
        for s in it:
            buf.append(s)
            if len(buf) >= self.shuffle_buffer:
                yield buf.pop(self._shuffle_rng.randrange(len(buf)))
        while buf:
            yield buf.pop(self._shuffle_rng.randrange(len(buf)))

    # -- one batch -----------------------------------------------------------------------------------------------------
    def _build(self, samples, slot):
        """samples: list of sample descriptions -> dict of device tensors (work enqueued on self.stream)."""
        keys = list(samples[0]["frames"])
        reqs = [r for k in keys for s in samples for r in s["frames"][k]]       # key-major: each key is contiguous
        n = len(reqs)
        offs, total = [], 0
        for r in reqs:
            offs.append(total)
            if not r.get("zero"):
                total += r["size"][0] * r["size"][1] * 3
        plans = augment.PlanTable(n, pin=False)
        plan_bytes = n * augment.PLAN_BYTES
        if slot.event is not None:
            slot.event.synchronize()               # the previous batch of this slot has left the staging buffer
        host = slot.reserve(total + plan_bytes + 16)
        host_np = host.numpy()

        def work(i):
            r = reqs[i]
            if r.get("zero"):
                plans.set_zero(i)
                return
            w, h = r["size"]
            px = decode_rgb(r["path"])
            if px.shape != (h, w, 3):
                raise ValueError("%s: decoded %s, header said %s" % (r["path"], px.shape, (h, w, 3)))
            host_np[offs[i]:offs[i] + px.size] = px.reshape(-1)
            plans.set(i, offs[i], w, h, r["resize"][0], r["resize"][1], r["crop"][0], r["crop"][1], r["angle"], r["flip"],
                      r["filter_id"], r["factor"])

        list(self.pool.map(work, range(n)))
        plan_off = (total + 15) // 16 * 16
        host[plan_off:plan_off + plan_bytes].copy_(plans.host[:plan_bytes])
        with torch.cuda.stream(self.stream):
            slot.dev[:plan_off + plan_bytes].copy_(host[:plan_off + plan_bytes], non_blocking=True)
            out = augment.augment_frames(slot.dev, slot.dev[plan_off:plan_off + plan_bytes], n, stream=self.stream)
            slot.event = torch.cuda.Event()
            slot.event.record(self.stream)
        batch, at = {}, 0
        for k in keys:
            per = len(samples[0]["frames"][k])
            cnt = per * len(samples)
            lead = (len(samples),) + ((per,) if samples[0].get("sequence") else ())
            batch[k] = out[at:at + cnt].view(*lead, IMAGE_SIZE, IMAGE_SIZE, 3)
            at += cnt
        for k in samples[0].get("extra", {}):
            batch[k] = torch.tensor([s["extra"][k] for s in samples], dtype=torch.int16)
        self.frames_done += n
        return batch, slot.event, total + plan_bytes

    def _batches(self):
        it = self._samples()
        if self.shuffle:
            it = self._shuffled(it)
        slots = collections.deque(_Staging(self.device) for _ in range(self.prefetch + 2))
        cur = []
        for s in it:
            cur.append(s)
            if len(cur) == self.batch_size:
                slots.rotate(-1)
                yield self._build(cur, slots[0])
                cur = []
        if cur:
            slots.rotate(-1)
            yield self._build(cur, slots[0])

    def __iter__(self):
        if not self.prefetch:
            for batch, ev, _ in self._batches():
                yield self._hand_over(batch, ev)
            return
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def producer():
            try:
                with torch.cuda.device(self.device):
                    for item in self._batches():
                        while not stop.is_set():
                            try:
                                q.put(item, timeout=0.1)
                                break
                            except queue.Full:
                                continue
                        if stop.is_set():
                            return
                q.put(None)
            except BaseException as exc:       # surfaces in the consumer
                q.put(exc)

        th = threading.Thread(target=producer, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield self._hand_over(item[0], item[1])
        finally:
            stop.set()

    def _hand_over(self, batch, ev):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in batch.values():
            if t.is_cuda:
                t.record_stream(cur)
        return batch


class BaseDataLoader(ABC):
    """data/base_dataloader.py:7-54: same abstract surface, ``get_dataset`` returns a ``DeviceDataset``."""

    @abstractmethod
    def length(self):
        raise NotImplementedError

    @abstractmethod
    def get_sample_dtype(self):
        raise NotImplementedError

    @abstractmethod
    def get_sample_shape(self):
        raise NotImplementedError

    @abstractmethod
    def sample_generator(self):
        raise NotImplementedError

    def map_fn(self, inputs):
        """[0,1] -> [-1,1] (image_pair_dataloader.py:63-69).  ``kp_augment_frames`` already applies it while writing its
        output, so the pipeline does not call this; kept for callers that hold [0,1] tensors."""
        return {k: (v * 2.0 - 1.0 if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in inputs.items()}

    def get_dataset(self, batch_size, repeat=False, shuffle=False, num_preprocess_threads=12, prefetch=True, device=None):
        return DeviceDataset(self, batch_size, repeat, shuffle, num_preprocess_threads, prefetch, device)
