"""B200-native implementation of the arcadia-microscopy-tools per-image hot path.

Same public surface as the reference for this path (ref: ``__init__.py:1-20``):
``MicroscopyImage``, ``ImageOperation``, ``Pipeline``, ``Channel`` and the two warning classes at
the top level; ``operations`` and ``masks`` as submodules.  Everything from uint16
intensities to per-cell tables runs in hand-written sm_100a CUDA kernels behind
``libamt_b200.so`` (C ABI in ``include/amt_b200.h``); importing the package does not need a
GPU, calling a compute entry point without one (or without the built library) raises.
"""

from .channels import Channel
from .exceptions import MetadataWarning, SegmentationWarning
from .microscopy import MicroscopyImage
from .pipeline import ImageOperation, Pipeline

__version__ = "0.1.0"

__all__ = [
    "Channel",
    "ImageOperation",
    "MetadataWarning",
    "MicroscopyImage",
    "Pipeline",
    "SegmentationWarning",
]
