"""Drop-in for gans/inversion.py (reference 10-97): the losses and the optimiser wrapper of the
latent-optimisation demo (BASELINE config 5) -- `SphericalOptimizer`, `masked_loss`,
`MultiScaleMaskedLoss`, `geocross_loss`, `normalize_noise_` with the reference's signatures.

The blur-pool pyramid of `MultiScaleMaskedLoss` (ring pad + depthwise 3x3 stride-2 convolution
for the images, the same with a box filter for the mask) runs on the package's polyphase FIR
kernel (`dusty_fir2d`: boundary modes in index math, decimation folded in): one launch per
pooled tensor instead of pad + grouped cuDNN conv, with the analytic adjoint for the gradient
w.r.t. the generated image.  CUDA tensors only, like the rest of the package.
"""
import functools

import numpy as np
import torch
import torch.nn.functional as F

from .. import _cabi as K
from .. import functional as DF


class SphericalOptimizer(torch.optim.Adam):
    """reference inversion.py:10-20: Adam step followed by a projection of every parameter onto
    the sphere of unit RMS along its last axis."""

    def __init__(self, params, **kwargs):
        params = list(params)
        super().__init__(params, **kwargs)
        self.params = params

    @torch.no_grad()
    def step(self, closure=None):
        loss = super().step(closure)
        for param in self.params:
            param.data.div_(param.pow(2).mean(dim=-1, keepdim=True).add(1e-9).sqrt())
        return loss


def masked_loss(img_ref, img_gen, mask, loss_fn=F.l1_loss, relative=True):
    """reference inversion.py:23-30."""
    loss = loss_fn(img_ref, img_gen, reduction="none")
    if relative:
        loss = (loss * mask) / img_ref.add(1e-11)
    loss = (loss * mask).sum(dim=(1, 2, 3))
    loss = loss / mask.sum(dim=(1, 2, 3)).add(1e-8)
    return loss


class MultiScaleMaskedLoss(torch.nn.Module):
    """reference inversion.py:33-80 (same buffers: `blur_kernel`, `mask_kernel` [1,1,3,3])."""

    def __init__(self, loss_fn, level=None, relative=True):
        super().__init__()
        blur_kernel = torch.tensor([1, 2, 1], dtype=torch.float32)
        blur_kernel = torch.outer(blur_kernel, blur_kernel)
        blur_kernel /= blur_kernel.sum()
        self.register_buffer("blur_kernel", blur_kernel[None, None])
        self.register_buffer("mask_kernel", torch.ones_like(blur_kernel)[None, None])
        self.dissimilarity = functools.partial(masked_loss, loss_fn=loss_fn, relative=relative)
        self.level = level
        # Pad(1, replicate, ring) + 3x3 stride-2 correlation as one polyphase FIR geometry
        self._cfg = DF.FirCfg(3, 3, down=(2, 2), pad=(1, 1, 1, 1), mode=(K.PAD_REPLICATE, K.PAD_CIRCULAR))

    def _pool(self, x, kernel):
        if not x.is_cuda:
            raise RuntimeError("MultiScaleMaskedLoss: CUDA tensors only (no CPU fallback)")
        return DF.fir2d(x.float(), kernel[0, 0].to(x.device, torch.float32).contiguous(), self._cfg)

    def blurpool(self, x):
        return self._pool(x, self.blur_kernel)

    def update_mask(self, mask):
        cnt = self._pool(mask, self.mask_kernel)
        norm = 1 / cnt.masked_fill(cnt == 0, 1.0)
        norm = norm * self.mask_kernel[0].numel()
        new_mask = torch.ones_like(cnt).masked_fill(cnt == 0, 0.0)
        return norm, new_mask

    def forward(self, gen, ref, mask):
        H = gen.shape[2]
        level = int(np.log2(H)) if self.level is None else self.level
        loss = 0
        for _ in range(max(1, level)):
            loss = loss + self.dissimilarity(ref, gen, mask)
            norm, new_mask = self.update_mask(mask)
            gen = self.blurpool(gen * mask) * norm
            ref = self.blurpool(ref * mask) * norm
            mask = new_mask
        return loss


def geocross_loss(latents):
    """reference inversion.py:83-91 (PULSE)."""
    B, N, D = latents.shape
    X = latents.view(B, 1, N, D)
    Y = latents.view(B, N, 1, D)
    A = ((X - Y).pow(2).sum(-1) + 1e-9).sqrt()
    Bm = ((X + Y).pow(2).sum(-1) + 1e-9).sqrt()
    Dm = 2 * torch.atan2(A, Bm)
    return (Dm.pow(2) * Dm).mean((1, 2)) / 8.0


def normalize_noise_(noises):
    """reference inversion.py:94-97."""
    for noise in noises:
        mean = noise.mean()
        std = noise.std()
        noise.data.add_(-mean).div_(std)
