"""dusty_gan_v2_b200 -- B200-native (sm_100a) implementation of the DUSty-v2 generator /
discriminator hot path behind the reference's own Python op / module interface.

Layout:
  csrc/ + libdusty_b200.so   hand-written CUDA kernels behind the C ABI (include/dusty_b200.h)
  _cabi.py                   ctypes binding (no torch types cross the boundary)
  functional.py              autograd-aware wrappers (forward / backward / double backward)
  gans/...                   host-side mirror of the reference modules: same names, ctor
                             arguments, state_dict keys (gans.models.ops, gans.models.*,
                             gans.coords.CoordBridge, gans.augment.adaptive_augment,
                             gans.trainer, gans.inversion, gans.utils)

There is no CPU or pure-PyTorch fallback: ops raise on CPU tensors and the package fails
loudly when the shared library is missing.
"""
import importlib
import sys

from . import _cabi
from .functional import act_dtype, set_modconv_impl, set_precision  # noqa: F401
from .config import AttrDict, load_config  # noqa: F401

__version__ = "0.1.0"

_MIRRORED = [
    "gans", "gans.coords", "gans.trainer", "gans.models", "gans.models.ops",
    "gans.models.ops.common", "gans.models.ops.style", "gans.models.ops.fourier",
    "gans.models.ops.gumbel", "gans.models.ops.fused_act", "gans.models.ops.fused_act.fused_act",
    "gans.models.ops.upfirdn2d", "gans.models.ops.upfirdn2d.upfirdn2d", "gans.models.base",
    "gans.models.dusty_v1", "gans.models.dusty_v2", "gans.models.vanilla", "gans.models.builder",
    "gans.models.loss", "gans.augment", "gans.augment.adaptive_augment", "gans.inversion", "gans.utils",
]


def install_as_gans():
    """Register this package's mirror under the reference's module names, so that
    `from gans.models import ops`, `from gans.coords import CoordBridge`,
    `from gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d` ... resolve to the B200
    implementation."""
    _cabi.load()
    for name in _MIRRORED:
        sys.modules[name] = importlib.import_module(f"{__name__}.{name}")


def library_path() -> str:
    return _cabi.LIB_PATH


def launch_count() -> int:
    """Number of libdusty_b200 kernel launches so far in this process."""
    return _cabi.launch_count()
