"""Same export surface as the reference's gans/models/ops/__init__.py:1-5."""
from .common import *          # noqa: F401,F403
from .common import BlurVH, Conv2d, EqualLR, MinibatchStdDev, Pad, PixelNorm, Resample, filter2d
from .fourier import FourierFeature
from .fused_act.fused_act import (FusedLeakyReLU, FusedLeakyReLUFunction,
                                  FusedLeakyReLUFunctionBackward, fused, fused_leaky_relu)
from .gumbel import GumbelSigmoid
from .style import ModConv2d, NoiseInjection
