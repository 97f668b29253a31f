"""Drop-in for gans/models/dusty_v1.py (reference 7-41): the measurement model that turns a
(range image, raydrop logit) pair into the final image, and the dusty_v1 generator built on the
transposed-convolution synthesis network of vanilla.py.

Execution differs from the reference's three ATen steps (relaxed-Bernoulli sample, hard
threshold with a straight-through gradient, lerp towards the raydrop constant): here the
Gumbel-sigmoid mask and the masked image come out of ONE kernel (dusty_gumbel_raydrop_fwd, with
dusty_gumbel_raydrop_bwd for the straight-through gradient), and the raydrop constant is kept as a
host float next to the registered buffer so that no step reads it back from the device.
"""
import torch
from torch import nn

from . import base, ops, vanilla

_REQUIRED = ("image", "raydrop_logit")


class RayDropModel(nn.Module):
    """state_dict: `raydrop_const` (0-dim buffer); sub-module `gumbel_sigmoid` as in the reference."""

    def __init__(self, raydrop_const: float, gumbel_temperature: float):
        super().__init__()
        self._const = float(raydrop_const)
        self.gumbel_sigmoid = ops.GumbelSigmoid(temperature=gumbel_temperature, straight_through=True)
        self.register_buffer("raydrop_const", torch.tensor(self._const))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # keep the host copy in step with a loaded checkpoint
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        loaded = state_dict.get(prefix + "raydrop_const")
        if loaded is not None:
            self._const = float(loaded)

    def forward(self, h):
        if not isinstance(h, dict) or any(k not in h for k in _REQUIRED):
            raise AssertionError(f"RayDropModel expects a dict with {_REQUIRED}")
        range_image = h["image"]
        mask, masked = self.gumbel_sigmoid(h["raydrop_logit"], range_image, self._const)
        h.update(raydrop_mask=mask, image_orig=range_image, image=masked)     # in place, like the reference
        return h

    def extra_repr(self):
        return f"raydrop_const={self._const}"


class Generator(base.Generator):
    """z [B, C] is used as the style directly (no mapping network); `angles` is accepted and ignored."""

    def __init__(self, synthesis_kwargs, measurement_kwargs):
        synthesis = vanilla.SynthesisNetwork(**synthesis_kwargs)
        super().__init__(mapping_network=nn.Identity(), synthesis_network=synthesis,
                         measurement_model=RayDropModel(**measurement_kwargs))

    def forward_synthesis(self, w, angles=None):
        return self.synthesis_network(w)
