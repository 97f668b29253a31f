"""Drop-in for gans/models/ops/gumbel.py: GumbelSigmoid (reference lines 5-32).
The uniform draw stays a PyTorch call (`torch.rand(logits.shape)`, exactly what
RelaxedBernoulli.rsample consumes) so seeds line up; the relaxed sample, threshold and
straight-through gradient are one kernel."""
import torch
from torch import nn

from .... import functional as DF


class GumbelSigmoid(nn.Module):
    def __init__(self, temperature: float = 1.0, straight_through: bool = True):
        super().__init__()
        self.temperature = temperature
        self.straight_through = straight_through

    def forward(self, logits, image=None, raydrop_const=0.0):
        """Reference call: forward(logits) -> mask.  With `image` it also returns the
        lerped image from the same kernel (used by RayDropModel)."""
        if not self.straight_through:
            raise NotImplementedError("only the straight-through estimator is on the hot path")
        u = torch.rand(logits.shape, dtype=torch.float32, device=logits.device)
        mask, out = DF.gumbel_raydrop(logits, logits if image is None else image, u,
                                      float(raydrop_const), float(self.temperature))
        return mask if image is None else (mask, out)

    def extra_repr(self):
        return f"tau={self.temperature}, straight_through={self.straight_through}"
