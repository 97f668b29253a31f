"""Drop-in for gans/models/ops/fourier.py: FourierFeature (reference lines 11-85).
Same buffers (`freqs[F,2,1,1]`, `phase[F]`) and the same RNG consumption at construction
(torch uniform_, numpy choice, torch rand -- seed both generators for reproducible init)."""
import numpy as np
import torch
import torch.nn as nn

from .... import functional as DF
from . import common as ops


class FourierFeature(nn.Module):
    def __init__(self, resolution, basis_scale="random", num_freqs=512, L_offset=(3, -1),
                 mapping=False, mapping_ch=64):
        super().__init__()
        self.resolution = resolution
        self.L_h = int(np.ceil(np.log2(resolution[0]))) + L_offset[0]
        self.L_w = int(np.ceil(np.log2(resolution[1]))) + L_offset[1]
        band_h, band_w = 2 ** (self.L_h - 1), 2 ** (self.L_w - 1)
        self.max_band = (band_h ** 2 + band_w ** 2) ** 0.5
        n = num_freqs // 2
        if basis_scale in ("random", "random_2"):
            f_h = torch.empty(n, 1).uniform_(-band_h, band_h)
            pool = 2 ** np.arange(self.L_w) if basis_scale == "random" else np.arange(band_w)
            pool = list(-pool) + [0] + list(pool)
            f_w = torch.from_numpy(np.random.choice(pool, size=(n, 1)))
            phase = torch.rand(n) * 2 * np.pi
            freqs = torch.cat([f_h, f_w], dim=-1)
        elif basis_scale == "logscale":
            L_min = min(self.L_h, self.L_w)
            ph, pw = torch.arange(self.L_h).exp2(), torch.arange(self.L_w).exp2()
            f_h = torch.cat([ph, torch.zeros(self.L_w), -ph[:L_min], ph[:L_min]])
            f_w = torch.cat([torch.zeros(self.L_h), pw, pw[:L_min], pw[:L_min]])
            freqs = torch.stack([f_h, f_w], dim=-1)
            phase = torch.zeros(len(f_h))
        else:
            raise ValueError(basis_scale)
        self.register_buffer("freqs", freqs[..., None, None])
        self.register_buffer("phase", phase)
        self.basis_ch = int(freqs.shape[0] * 2)
        if mapping:
            self.out_ch = mapping_ch
            self.mapping = ops.EqualLR(nn.Conv2d(self.basis_ch, self.out_ch, 1, 1, 0, bias=False))
        else:
            self.out_ch = self.basis_ch
            self.mapping = None

    def forward(self, angles, out_dtype=None):
        enc = DF.fourier_features(angles, self.freqs, self.phase, out_dtype)
        if self.mapping is not None:
            enc = self.mapping(enc)
        return enc

    def extra_repr(self):
        return f"shape={self.resolution}, num_freqs={self.basis_ch}, L=({self.L_h}, {self.L_w})"
