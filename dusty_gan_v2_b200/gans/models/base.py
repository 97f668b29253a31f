"""Drop-in for gans/models/base.py: StyleGAN-style generator template (reference 8-142)."""
import random

import torch
from torch import nn


class Generator(nn.Module):
    def __init__(self, mapping_network: nn.Module = nn.Identity(),
                 synthesis_network: nn.Module = nn.Identity(),
                 measurement_model: nn.Module = nn.Identity(), w_avg_decay: float = 0.995):
        super().__init__()
        self.mapping_network = mapping_network
        self.synthesis_network = synthesis_network
        self.measurement_model = measurement_model
        self.w_avg_decay = w_avg_decay
        self.register_buffer("w_avg", torch.zeros(1, self.synthesis_network.in_ch))

    def forward(self, z, angle=None, style_mixing=False, truncation_psi=1.0, input_w=False):
        w = z if input_w else self.forward_mapping(z, style_mixing)
        if w.ndim != 3:
            raise RuntimeError("style codes must be [B, num_styles, D]")
        if self.training:
            self.moving_average_w(w)
        else:
            w = self.truncation_trick(w, truncation_psi)
        out = self.forward_synthesis(w, angle)
        out["w"] = w
        return self.forward_measurement(out)

    def forward_mapping(self, z, style_mixing=False):
        n = self.synthesis_network.num_styles
        w1 = self.mapping_network(z)
        if not style_mixing:
            return w1.unsqueeze(1).expand(-1, n, -1)
        w2 = self.mapping_network(torch.randn_like(z))
        cut = random.randint(1, n)
        return torch.stack([w1] * cut + [w2] * (n - cut), dim=1)

    @torch.no_grad()
    def moving_average_w(self, w):
        batch_mean = w[:, 0].mean(dim=0, keepdim=True).to(self.w_avg)
        self.w_avg.lerp_(batch_mean, 1 - self.w_avg_decay)

    def truncation_trick(self, w, psi=1.0):
        if psi != 1.0:
            w = torch.lerp(self.w_avg[None].expand_as(w), w, psi)
        return w

    def forward_synthesis(self, w, angle=None):
        raise NotImplementedError

    def forward_measurement(self, x):
        return self.measurement_model(x)
