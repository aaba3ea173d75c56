"""Import alias: ``import kp_b200`` loads the package directory whose on-disk name carries hyphens."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "unsupervised-keypoint-learning-for-guiding-class-conditional-video-prediction_b200")
_spec = importlib.util.spec_from_file_location("kp_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["kp_b200"] = _mod
_spec.loader.exec_module(_mod)
