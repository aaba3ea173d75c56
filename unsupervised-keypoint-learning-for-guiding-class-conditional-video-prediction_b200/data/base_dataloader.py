"""``BaseDataLoader`` (reference: data/base_dataloader.py:7-54) and the batch pipeline behind ``get_dataset``.

The reference builds ``tf.data.Dataset.from_generator(...).map(map_fn, 12 threads).batch(B).prefetch(1)``, its generator
doing ALL pixel work with Pillow on one Python thread.  Here ``sample_generator`` yields sample *descriptions* - which
files, and what the drawn rotation / crop / flip / filter are - and ``DeviceDataset`` turns ``batch_size`` of them into
device tensors:

    main (or prefetch) thread : pull descriptions (random draws in the reference's order, image sizes from JPEG headers)
    decode workers (12)       : PIL decodes each JPEG straight into the batch's staging buffer - worker PROCESSES writing
                                into a shared mapping that is registered with CUDA as pinned memory (PIL's JPEG plugin
                                holds the GIL for its Python-level parsing, threads do not scale), or threads
    side CUDA stream          : H2D of frames + plans, ONE kp_augment_frames launch -> float32 [n,128,128,3] in [-1,1]
    consumer                  : current stream waits on the batch's event

A frame request is a dict: ``path`` (or ``zero``), ``size`` (w, h of the JPEG), ``resize`` (W, H), ``crop`` (left, top),
``angle``, ``flip``, ``filter_id``, ``factor``.
"""
import atexit
import collections
import concurrent.futures
import mmap
import os
import pickle
import queue
import random
import struct
import subprocess
import sys
import tempfile
import threading
from abc import ABC, abstractmethod

import numpy as np
import torch
from PIL import Image

from .. import _lib, augment

IMAGE_SIZE = augment.IMAGE_SIZE


def frame_request(path, size, resize, crop, angle=0, flip=0):
    return {"path": path, "size": tuple(size), "resize": tuple(resize), "crop": tuple(crop), "angle": angle, "flip": flip,
            "filter_id": augment.NO_FILTER, "factor": 0.0}


def zero_frame():
    return {"zero": True}


_DIR_LEN, _JPEG_SIZE = {}, {}


def dir_len(folder):
    """len(os.listdir(folder)) (image_pair_dataloader.py:74), remembered: a video's frame count does not change."""
    n = _DIR_LEN.get(folder)
    if n is None:
        n = _DIR_LEN[folder] = len(os.listdir(folder))
    return n


def jpeg_size(path):
    """(w, h) from the JPEG header, remembered per file (the reference reads it from the opened image, :95)."""
    s = _JPEG_SIZE.get(path)
    if s is None:
        with Image.open(path) as im:
            s = _JPEG_SIZE[path] = im.size
    return s


def decode_rgb(path):
    """The decoded frame as uint8 [h, w, 3] (the reference hands PIL's decode to ``np.asarray`` as is)."""
    with Image.open(path) as im:
        return np.asarray(im if im.mode == "RGB" else im.convert("RGB"))


_STAGE_FILES = set()


def _unlink_leftovers():
    for path in list(_STAGE_FILES):
        try:
            os.unlink(path)
        except OSError:
            pass


atexit.register(_unlink_leftovers)


class _Staging:
    """One batch in flight: frame bytes + plans in a shared, CUDA-registered host mapping (decode workers write into it),
    their device copy, and the event that marks the batch's kernel."""
    _count = [0]

    def __init__(self, device):
        self.device = device
        self.host = self.dev = self.event = self.path = self._mm = None
        self._registered = False
        _Staging._count[0] += 1
        self._stem = "kp_b200_stage_%d_%d" % (os.getpid(), _Staging._count[0])
        self._gen = 0

    def reserve(self, nbytes):
        if self.host is None or self.host.numel() < nbytes:
            self.release()
            cap = max(int(nbytes * 1.25), 1 << 20)
            shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
            self._gen += 1
            self.path = os.path.join(shm, "%s_%d" % (self._stem, self._gen))
            _STAGE_FILES.add(self.path)
            with open(self.path, "w+b") as fh:
                fh.truncate(cap)
                self._mm = mmap.mmap(fh.fileno(), cap)
            self.host = torch.frombuffer(self._mm, dtype=torch.uint8)
            torch.cuda.current_stream(self.device)         # the CUDA context exists before the library registers memory
            self._registered = _lib.load().kp_host_register(self.host.data_ptr(), cap) == 0
            self._pinned = None if self._registered else torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            self.dev = torch.empty(cap, dtype=torch.uint8, device=self.device)
        return self.host

    def upload_source(self, nbytes):
        """The host tensor to copy from: the registered mapping itself, or a pinned bounce buffer if registering failed."""
        if self._registered:
            return self.host
        self._pinned[:nbytes].copy_(self.host[:nbytes])
        return self._pinned

    def release(self):
        if self.host is not None:
            if self.event is not None:
                self.event.synchronize()
            if self._registered:
                _lib.load().kp_host_unregister(self.host.data_ptr())      # status ignored: nothing to do about it here
                self._registered = False
            self.host = None
        if self.path is not None:
            try:
                os.unlink(self.path)
            except OSError:
                pass
            _STAGE_FILES.discard(self.path)
            self.path = None


class _DecodeWorkers:
    """Plain subprocesses running data/_kp_decode_worker.py (numpy + PIL only; never CUDA)."""

    def __init__(self, n):
        script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_kp_decode_worker.py")
        self.procs = [subprocess.Popen([sys.executable, script], stdin=subprocess.PIPE, stdout=subprocess.PIPE)
                      for _ in range(n)]
        atexit.register(self.close)

    def run(self, tasks):
        """tasks: list of (staging file, offset, path, w, h), dealt round-robin; raises on the first worker error."""
        self.submit(tasks)()

    def submit(self, tasks):
        """Send the tasks and return a function that waits for them.  Requests are answered in order, so several
        submissions may be outstanding (their waiters must be called in submission order)."""
        used = []
        for i, p in enumerate(self.procs):
            part = tasks[i::len(self.procs)]
            if part:
                blob = pickle.dumps(part)
                p.stdin.write(struct.pack("<I", len(blob)) + blob)
                p.stdin.flush()
                used.append(p)

        def wait():
            err = None
            for p in used:
                hdr = p.stdout.read(4)
                if len(hdr) < 4:
                    raise RuntimeError("a decode worker died (exit code %r)" % p.poll())
                err = pickle.loads(p.stdout.read(struct.unpack("<I", hdr)[0])) or err
            if err:
                raise RuntimeError("decode worker: " + err)

        return wait

    def drain(self):
        """After a failed request: stop the workers (later submissions' replies would be mis-paired otherwise)."""
        self.close()

    def close(self):
        for p in self.procs:
            try:
                p.stdin.close()
            except OSError:
                pass
        for p in self.procs:
            try:
                p.wait(timeout=2)
            except subprocess.TimeoutExpired:
                p.kill()
        self.procs = []


class DeviceDataset:
    """Iterable of batches ``{key: CUDA tensor}``; what ``BaseDataLoader.get_dataset`` returns."""

    def __init__(self, loader, batch_size, repeat, shuffle, num_preprocess_threads, prefetch, device, shuffle_buffer=2000,
                 decode="process", sample_range=None):
        if not torch.cuda.is_available():
            raise RuntimeError("the input pipeline augments on the GPU; there is no CPU path")
        if decode not in ("process", "thread"):
            raise ValueError("decode must be 'process' or 'thread'")
        self.loader, self.batch_size, self.repeat, self.shuffle = loader, int(batch_size), repeat, shuffle
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        # decode workers: what the caller asked for, leaving two cores to the consumer and the batch-building thread
        self.n_workers = max(1, min(int(num_preprocess_threads), (os.cpu_count() or 3) - 2))
        self.decode = decode
        self.sample_range = sample_range
        self.pool = concurrent.futures.ThreadPoolExecutor(self.n_workers)
        self.workers = None
        self._slots = []
        self._producer = None
        self.prefetch = 1 if prefetch else 0
        self.stream = torch.cuda.Stream(device=self.device)
        self.shuffle_buffer = shuffle_buffer
        self._shuffle_rng = random.Random(0x5eed)      # tf.data's shuffle has its own generator too (not reproducible)
        self.frames_done = 0

    # -- sample stream -------------------------------------------------------------------------------------------------
    def _samples(self):
        while True:
            if self.sample_range is None:
                yield from self.loader.sample_generator()
            else:
                yield from self.loader.sample_generator(*self.sample_range)
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
    def _begin(self, samples, slot):
        """First half of a batch: lay the frames out in the slot's staging buffer, hand the JPEGs to the decode workers
        (not waited for) and build the frame plans while they decode."""
        keys = list(samples[0]["frames"])
        reqs = [r for k in keys for s in samples for r in s["frames"][k]]       # key-major: each key is contiguous
        n = len(reqs)
        offs, total = [], 0
        for r in reqs:
            offs.append(total)
            if not r.get("zero"):
                total += r["size"][0] * r["size"][1] * 3
        plan_bytes = n * augment.PLAN_BYTES
        if slot.event is not None:
            slot.event.synchronize()               # the previous batch of this slot has left the staging buffer
        host = slot.reserve(total + plan_bytes + 16)
        live = [i for i in range(n) if not reqs[i].get("zero")]
        if self.decode == "process":
            if self.workers is None:
                self.workers = _DecodeWorkers(self.n_workers)
            waiter = self.workers.submit([(slot.path, offs[i], reqs[i]["path"]) + tuple(reqs[i]["size"]) for i in live])
        else:
            host_np = host.numpy()

            def work(i):
                w, h = reqs[i]["size"]
                px = decode_rgb(reqs[i]["path"])
                if px.shape != (h, w, 3):
                    raise ValueError("%s: decoded %s, header said %s" % (reqs[i]["path"], px.shape, (h, w, 3)))
                host_np[offs[i]:offs[i] + px.size] = px.reshape(-1)

            futures = [self.pool.submit(work, i) for i in live]
            waiter = lambda: [f.result() for f in futures]      # noqa: E731
        plans = augment.PlanTable(n, pin=False)
        plans.set_batch(reqs, offs)                 # one library call (this thread shares the GIL with the consumer)
        return dict(samples=samples, keys=keys, n=n, total=total, plans=plans, slot=slot, waiter=waiter)

    def _finish(self, pend):
        """Second half: wait for the decoded frames, then one H2D copy and one kernel launch on the side stream."""
        samples, keys, n, total, slot = pend["samples"], pend["keys"], pend["n"], pend["total"], pend["slot"]
        try:
            pend["waiter"]()
        except Exception:
            if self.workers is not None:
                self.workers.drain()
                self.workers = None
            raise
        plan_bytes = n * augment.PLAN_BYTES
        plan_off = (total + 15) // 16 * 16
        slot.host[plan_off:plan_off + plan_bytes].copy_(pend["plans"].host[:plan_bytes])
        with torch.cuda.stream(self.stream):
            up = slot.upload_source(plan_off + plan_bytes)
            slot.dev[:plan_off + plan_bytes].copy_(up[:plan_off + plan_bytes], non_blocking=True)
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
        """Two batches in flight on the host: the decode of batch i+1 is handed to the workers before batch i is waited
        for, so the workers never idle while this thread draws samples, builds plans or enqueues GPU work."""
        it = self._samples()
        if self.shuffle:
            it = self._shuffled(it)
        slots = collections.deque(_Staging(self.device) for _ in range(self.prefetch + 3))
        self._slots.extend(slots)
        cur, pending = [], None
        for s in it:
            cur.append(s)
            if len(cur) == self.batch_size:
                slots.rotate(-1)
                nxt = self._begin(cur, slots[0])
                if pending is not None:
                    yield self._finish(pending)
                pending, cur = nxt, []
        if cur:
            slots.rotate(-1)
            nxt = self._begin(cur, slots[0])
            if pending is not None:
                yield self._finish(pending)
            pending = nxt
        if pending is not None:
            yield self._finish(pending)

    def __iter__(self):
        """One iterator at a time: the staging slots and the decode workers belong to it."""
        self._stop_producer()
        if not self.prefetch:
            for batch, ev, _ in self._batches():
                yield self._hand_over(batch, ev)
            return
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def offer(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def producer():
            try:
                with torch.cuda.device(self.device):
                    for item in self._batches():
                        if not offer(item):
                            return
                offer(None)
            except BaseException as exc:       # surfaces in the consumer
                offer(exc)

        th = threading.Thread(target=producer, daemon=True)
        self._producer = (th, stop)
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

    def _stop_producer(self):
        th, stop = self._producer or (None, None)
        if th is not None:
            stop.set()
            if th is not threading.current_thread():     # the last reference may die on the producer thread itself
                th.join(timeout=10)
        self._producer = None

    def close(self):
        """Stop the batch-building thread and the decode workers, drop the staging mappings (also runs at garbage
        collection)."""
        try:
            self._stop_producer()
            if self.workers is not None:
                self.workers.close()
                self.workers = None
        finally:            # whatever happened above, page-locked staging memory must not outlive its mapping
            for sl in self._slots:
                sl.release()
            self._slots = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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

    def get_dataset(self, batch_size, repeat=False, shuffle=False, num_preprocess_threads=12, prefetch=True, device=None,
                    decode="process", sample_range=None):
        """``decode``: 'process' (worker processes, the default) or 'thread' (PIL on a thread pool inside this process).
        ``sample_range`` = (start, stop): forwarded to ``sample_generator`` by loaders that shard (KeypointDataLoader)."""
        return DeviceDataset(self, batch_size, repeat, shuffle, num_preprocess_threads, prefetch, device, decode=decode,
                             sample_range=sample_range)
