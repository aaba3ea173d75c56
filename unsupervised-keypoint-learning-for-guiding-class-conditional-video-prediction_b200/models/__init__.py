"""Model classes of the stage-1 path (mirror of the reference's ``models`` package, minus stage 2)."""
from .base_model import BaseModel, GlobalStep  # noqa: F401
from .detector_translator_model import DetectorTranslatorModel  # noqa: F401
from .final_model import FinalModel  # noqa: F401
from .keypoint_model import KeypointModel  # noqa: F401
