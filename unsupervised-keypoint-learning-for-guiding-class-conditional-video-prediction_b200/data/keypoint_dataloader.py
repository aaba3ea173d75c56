"""``KeypointDataLoader`` (reference: data/keypoint_dataloader.py:17-88): every frame of a video, resized and centre-cropped
by the FIRST frame's geometry, zero frames appended up to 663; ``len`` and ``idx`` ride along (make_pseudo_labels.py:83-101)."""
from os import path as osp

import numpy as np

from ..utils import data as data_utils
from .base_dataloader import IMAGE_SIZE, BaseDataLoader, dir_len, frame_request, jpeg_size, zero_frame

MIN_IMAGE_SEQ_LEN = 663


class KeypointDataLoader(BaseDataLoader):

    def __init__(self, data_dir, subset):
        super().__init__()
        self._data_dir = data_dir
        with open(osp.join(data_dir, subset + "_set.txt")) as fh:
            self._images = fh.read().splitlines()
        self._total = len(self._images)
        print(subset + "set : ", self._total)

    def length(self):
        return self._total

    def get_sample_shape(self):
        return {"image": [MIN_IMAGE_SEQ_LEN, IMAGE_SIZE, IMAGE_SIZE, 3], "len": None, "idx": None}

    def get_sample_dtype(self):
        return {"image": np.float32, "len": np.int16, "idx": np.int16}

    def sample_generator(self, start=0, stop=None):
        """All videos in list order (the reference); ``start`` / ``stop`` restrict it to one rank's shard."""
        for idx in range(start, self._total if stop is None else min(stop, self._total)):
            yield self._get_image_at(idx)

    def _get_image_at(self, idx):
        img_path = self._images[idx].split()[0]
        folder = osp.join(self._data_dir, img_path)
        file_len = dir_len(folder)
        paths = [osp.join(folder, "%06d.jpg" % (i + 1)) for i in range(file_len)]
        sizes = [jpeg_size(p) for p in paths]
        w, h = sizes[0]
        box, ratio = data_utils.center_crop((w, h), IMAGE_SIZE)
        resize = (int(w / ratio), int(h / ratio))
        frames = [frame_request(p, s, resize, box[:2]) for p, s in zip(paths, sizes)]
        frames += [zero_frame() for _ in range(MIN_IMAGE_SEQ_LEN - file_len)]
        return {"frames": {"image": frames}, "sequence": True,
                "extra": {"idx": int(img_path.split("/")[-1]), "len": file_len}}
