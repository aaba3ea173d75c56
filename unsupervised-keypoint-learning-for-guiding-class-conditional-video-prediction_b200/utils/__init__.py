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
