"""Mirror of the reference's ``utils/data.py``.  The Pillow work itself runs on the device (csrc/augment.cu); the functions
here make the same RANDOM DRAWS, in the same order, from the same global ``random`` module, and describe the result."""
import math
import random

import numpy as np

# apply_random_filter's r_id -> the kernel's filter_id (identical numbering: utils/data.py:11-33)
_FACTOR_RANGE = {6: (0, 50), 7: (7, 20), 8: (0, 50), 9: (7, 30)}     # Sharpness, Brightness, Color, Contrast


def apply_random_filter(images):
    """utils/data.py:8-35.  ``images`` is a list of frame requests (dicts, see data/base_dataloader.py); every one gets the
    same drawn filter, exactly as the reference filters both frames of a pair alike."""
    r_id = random.randint(0, 9)
    factor = 0.0
    if r_id in _FACTOR_RANGE:
        lo, hi = _FACTOR_RANGE[r_id]
        factor = random.randint(lo, hi) * 0.1
    for im in images:
        im["filter_id"], im["factor"] = r_id, factor
    return images


def center_crop(size, target_size):
    """utils/data.py:38-59 on a (w, h) pair instead of a PIL image: (crop box, ratio)."""
    w, h = size
    half = target_size // 2
    if w > h:
        ratio = h / float(target_size)
        ox = int(w / ratio) / 2.0
        return (ox - half, 0, ox + half, target_size), ratio
    ratio = w / float(target_size)
    oy = int(h / ratio) / 2.0
    return (0, oy - half, target_size, oy + half), ratio


def rotate_keypoints(keypoints, rand_val, ox=0, oy=0):
    """utils/data.py:62-71."""
    c, s = math.cos(math.radians(-rand_val)), math.sin(math.radians(-rand_val))
    dx, dy = keypoints[..., 0] - ox, keypoints[..., 1] - oy
    return np.stack([ox + c * dx - s * dy, oy + s * dx + c * dy], axis=-1)


def create_one_hot_label(n_classes, idx):
    """utils/data.py:74-78."""
    label = np.zeros(n_classes)
    label[int(idx)] = 1
    return label
