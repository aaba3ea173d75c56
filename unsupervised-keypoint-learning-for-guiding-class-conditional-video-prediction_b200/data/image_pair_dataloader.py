"""``ImagePairDataLoader`` (reference: data/image_pair_dataloader.py:16-165): a frame and the frame 8-11 steps later of one
video, jointly rotated / resized / cropped / flipped / filtered.  The draws come from the global ``numpy.random`` and
``random`` generators in the reference's order, so seeding both reproduces the reference's sample stream."""
import random
from os import path as osp

import numpy as np

from ..utils import data as data_utils
from .base_dataloader import IMAGE_SIZE, BaseDataLoader, dir_len, frame_request, jpeg_size


class ImagePairDataLoader(BaseDataLoader):

    def __init__(self, data_dir, subset, random_order=True, randomness=False):
        super().__init__()
        self._data_dir, self._random_order, self._randomness = data_dir, random_order, randomness
        with open(osp.join(data_dir, subset + "_set.txt")) as fh:
            self._images = fh.read().splitlines()
        self._total = len(self._images)
        print(subset + "set : ", self._total)

    def length(self):
        return self._total

    def get_sample_shape(self):
        return {"image": [IMAGE_SIZE, IMAGE_SIZE, 3], "future_image": [IMAGE_SIZE, IMAGE_SIZE, 3]}

    def get_sample_dtype(self):
        return {"image": np.float32, "future_image": np.float32}

    def sample_generator(self):
        if not self._random_order:
            for idx in range(self._total):
                yield self._get_image_at(idx)
            return
        for _ in range(self._total):
            yield self._get_image_at(np.random.randint(len(self._images)))

    def _get_image_at(self, idx):
        """The description of pair ``idx``: draws in the order of image_pair_dataloader.py:72-150."""
        img_path = self._images[idx].split()[0]
        folder = osp.join(self._data_dir, img_path)
        file_len = dir_len(folder)
        first, second = 0, 10
        if self._random_order:
            step = random.randint(8, 11)
            first = random.randint(0, file_len - 1)
            second = (first + step) % file_len
        paths = [osp.join(folder, "%06d.jpg" % (i + 1)) for i in (first, second)]
        sizes = [jpeg_size(p) for p in paths]          # header only
        w, h = sizes[0]
        angle = random.randrange(-10, 11) if self._randomness else 0
        wide = w > h
        ratio = (h if wide else w) / float(IMAGE_SIZE)
        resize = (int(w / ratio), int(h / ratio))            # both frames get the FIRST frame's target size
        flip = 0
        if self._randomness:
            shift = random.randint(0, int((w if wide else h) / ratio - IMAGE_SIZE))
            flip = random.randint(0, 1)
            crop = (shift, 0) if wide else (0, shift)
        else:
            crop = (resize[0] / 2.0 - IMAGE_SIZE // 2, 0)    # the reference centres along x in BOTH branches (:124-130,155-161)
        frames = [frame_request(p, s, resize, crop, angle, flip) for p, s in zip(paths, sizes)]
        if self._randomness:
            data_utils.apply_random_filter(frames)
        return {"frames": {"image": frames[:1], "future_image": frames[1:]}}
