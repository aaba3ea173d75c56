"""Mirror of the reference's ``data`` package for stage 1: the loaders keep their names, constructor
arguments, random draws and output contract; JPEG decoding stays on host threads (PIL, like the reference), every pixel
operation after it runs in one CUDA kernel per batch (csrc/augment.cu)."""
from .base_dataloader import BaseDataLoader, DeviceDataset  # noqa: F401
from .image_pair_dataloader import ImagePairDataLoader  # noqa: F401
from .keypoint_dataloader import KeypointDataLoader  # noqa: F401
