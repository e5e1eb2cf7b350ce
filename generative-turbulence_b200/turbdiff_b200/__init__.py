"""turbdiff_b200 - B200-native (sm_100a) implementation of TurbDiff's denoising hot path.

Drop-in for the reference's ``turbdiff.models.ddpm`` classes (``DenoisingModel``,
``GaussianDiffusion``): same constructor arguments, parameter names and call signatures,
with the compute in hand-written CUDA kernels behind the C ABI of ``libturbdiff_b200.so``
(``include/turbdiff_b200.h``).  There is no CPU fallback: using the models without the
compiled library or without a CUDA device raises.
"""

from . import _lib  # noqa: F401
from .models.ddpm import DenoisingModel, GaussianDiffusion, ModelPrediction  # noqa: F401

__all__ = ["DenoisingModel", "GaussianDiffusion", "ModelPrediction"]
