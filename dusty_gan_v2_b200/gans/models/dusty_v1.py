"""Drop-in for gans/models/dusty_v1.py: RayDropModel (7-28) and the dusty_v1 generator."""
import torch
from torch import nn

from . import base, ops, vanilla


class RayDropModel(nn.Module):
    def __init__(self, raydrop_const: float, gumbel_temperature: float):
        super().__init__()
        self.gumbel_sigmoid = ops.GumbelSigmoid(temperature=gumbel_temperature,
                                                straight_through=True)
        self.register_buffer("raydrop_const", torch.tensor(float(raydrop_const)))
        self._const = float(raydrop_const)     # host copy: no device sync per step

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        key = prefix + "raydrop_const"
        if key in state_dict:
            self._const = float(state_dict[key])

    def forward(self, h):
        assert isinstance(h, dict) and ("image" in h) and ("raydrop_logit" in h)
        mask, image = self.gumbel_sigmoid(h["raydrop_logit"], h["image"], self._const)
        h["raydrop_mask"] = mask
        h["image_orig"] = h["image"]
        h["image"] = image
        return h

    def extra_repr(self):
        return f"raydrop_const={self._const}"


class Generator(base.Generator):
    def __init__(self, synthesis_kwargs, measurement_kwargs):
        super().__init__(mapping_network=nn.Identity(),
                         synthesis_network=vanilla.SynthesisNetwork(**synthesis_kwargs),
                         measurement_model=RayDropModel(**measurement_kwargs))

    def forward_synthesis(self, w, angles=None):
        return self.synthesis_network(w)
