"""B200-native stage-1 hot path (detector -> keypoints -> rendered maps -> translated frame).

Import name: ``kp_b200`` (see kp_b200.py at the repo root; the directory name carries hyphens).
The compute lives in libkp_b200.so (hand-written sm_100a CUDA behind the C ABI in include/kp_b200.h);
this package is the host-side mirror of the reference's ``utils`` / ``models.networks`` / ``models`` API.
"""
from . import _lib  # noqa: F401
from . import k1  # noqa: F401
from . import tapconv  # noqa: F401
from . import conv  # noqa: F401
from . import ops  # noqa: F401
from . import engine  # noqa: F401
from .utils import model as model_utils  # noqa: F401
from . import networks  # noqa: F401
from . import models  # noqa: F401
from . import checkpoint  # noqa: F401
from . import pseudo_labels  # noqa: F401

__all__ = ["model_utils", "k1", "networks", "models", "engine", "ops", "conv", "tapconv", "checkpoint", "pseudo_labels"]
