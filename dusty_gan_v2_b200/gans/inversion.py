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


# ---- the optimisation loop of demo_inversion.py (reference 89-255), without the viewer ----
def tanh_to_sigmoid(x):
    """reference gans/utils.py:102-105."""
    return (x + 1.0) / 2.0


def lr_schedule(iteration, num_steps, rampup_ratio=0.05, rampdown_ratio=0.25):
    """reference demo_inversion.py:140-146 (StyleGAN2 projector schedule)."""
    t = iteration / num_steps
    gamma = min(1.0, (1.0 - t) / rampdown_ratio)
    gamma = 0.5 - 0.5 * np.cos(gamma * np.pi)
    return gamma * min(1.0, t / rampup_ratio)


class LatentInversion:
    """Stage 1 (latent optimisation against a frozen, eval-mode generator) and stage 2 (pivotal
    tuning of the generator weights) of the reference's inversion demo for a BATCH of targets
    (BASELINE config 5: 256 range images).  Keyword names follow the demo's command line.

    `depth` is metric depth [B,1,H,W], `mask` its validity mask; `coord` a CoordBridge.
    """

    def __init__(self, G, coord, depth, mask, latent_type="w", num_steps_1st=500,
                 num_steps_2nd=500, lr_1st=5e-2, lr_1st_rampup_ratio=0.05,
                 lr_1st_rampdown_ratio=0.25, lr_2nd=5e-4, noise_ratio=0.75, noise_coef=0.05 / 10,
                 optimize_phase=False, perturb_z=False, hypersphere_z=False, seed=0,
                 num_z_samples=10_000):
        if latent_type not in ("z", "w", "w+"):
            raise ValueError(f"{latent_type=}")
        if optimize_phase:
            # demo_inversion.py's --optimize_phase needs d(image)/d(angle): the Fourier operand of
            # the modulated contraction is a non-differentiable input of our kernels
            # (functional.fourier_features / modconv_bmm detach it), so the option would
            # silently optimise nothing.  Refuse it instead.
            raise NotImplementedError(
                "LatentInversion(optimize_phase=True): the gradient w.r.t. the angle grid is not "
                "implemented by the B200 kernels (Fourier features are a constant operand)")
        if not depth.is_cuda:
            raise RuntimeError("LatentInversion: CUDA tensors only (no CPU fallback)")
        self.G, self.coord = G, coord
        self.latent_type, self.perturb_z, self.hypersphere_z = latent_type, perturb_z, hypersphere_z
        self.num_steps_1st, self.num_steps_2nd = num_steps_1st, num_steps_2nd
        self.lr_1st, self.lr_2nd = lr_1st, lr_2nd
        self.rampup, self.rampdown = lr_1st_rampup_ratio, lr_1st_rampdown_ratio
        self.noise_ratio, self.noise_coef = noise_ratio, noise_coef
        dev = depth.device
        B = depth.shape[0]
        # targets (demo 89-94)
        self.t_mask = mask.float()
        self.t_depth = coord.convert(depth.float(), "depth", "depth_norm")
        self.t_inv_depth = coord.convert(self.t_depth, "depth_norm", "inv_depth_norm") * self.t_mask
        # latent initialisation (demo 99-120)
        syn = G.synthesis_network
        with torch.no_grad():
            z_dim = G.mapping_network.in_ch
            torch.manual_seed(seed)
            z_samples = G.mapping_network(torch.randn(num_z_samples, z_dim, device=dev))
            z_avg = z_samples.mean(dim=0, keepdim=True)
            self.z_std = (((z_samples - z_avg) ** 2).sum() / num_z_samples).sqrt()
            if hypersphere_z:
                z_avg.div_(z_avg.pow(2).mean(dim=-1, keepdim=True).add(1e-9).sqrt())
        z_avg = z_avg.repeat_interleave(B, dim=0)
        if latent_type == "z":
            z = torch.randn(B, z_dim, device=dev)
        elif latent_type == "w":
            z = z_avg
        else:
            z = torch.stack([z_avg] * syn.num_styles, dim=1)
        self.z = torch.nn.Parameter(z.clone()).requires_grad_()
        self.params_1st = [self.z]
        # noise inputs (demo 122-131): only present when the generator was built with use_noise
        self.noises = []
        from .models import ops
        for m in G.modules():
            if isinstance(m, ops.NoiseInjection) and m.fixed_noise is not None:
                noise = torch.randn_like(m.fixed_noise, dtype=torch.float32)
                m.fixed_noise = noise
                if len(self.noises) < 9:
                    noise.requires_grad = True
                    self.noises.append(noise)
        self.params_1st += self.noises
        self.phase = torch.nn.Parameter(torch.zeros((B, 2, 1, 1), device=dev)).requires_grad_()
        if optimize_phase:
            self.params_1st += [self.phase]
        self.optimize_phase = optimize_phase
        self.criterion = MultiScaleMaskedLoss(loss_fn=F.l1_loss, level=2).to(dev)
        self.optim_1st = self.scheduler = self.optim_2nd = None
        self.last_valid_count = None

    # one step forward (demo 149-191)
    def forward(self, progress=0.0):
        G, coord, z = self.G, self.coord, self.z
        n = G.synthesis_network.num_styles
        if self.latent_type == "z":
            w = G.forward_mapping(z, None)
        elif self.latent_type == "w":
            w = torch.stack([z] * n, dim=1)
        else:
            w = z
        if self.perturb_z:
            t = max(0.0, 1.0 - progress / self.noise_ratio)
            w = w + self.noise_coef * self.z_std * (t ** 2) * torch.randn_like(w)
        # an un-optimised phase is identically zero: keep the batch-shared angle grid (one
        # Fourier block for the whole batch) instead of materialising angle + 0 per sample
        angle = coord.angle + self.phase if self.optimize_phase else coord.angle
        imgs = G(w, angle=angle, input_w=True)
        g_inv_depth = tanh_to_sigmoid(imgs["image"])
        g_inv_depth_orig = tanh_to_sigmoid(imgs["image_orig"])
        g_raydrop_prob = torch.sigmoid(imgs["raydrop_logit"])
        g_depth = coord.convert(g_inv_depth_orig, "inv_depth_norm", "depth_norm")
        loss = 0
        if self.latent_type == "w+":
            loss = loss + 5e-3 * geocross_loss(w)
        loss = loss + self.criterion(g_depth, self.t_depth, self.t_mask)
        loss = loss + self.criterion(g_inv_depth_orig, self.t_inv_depth, self.t_mask)
        out = dict(inv_depth=g_inv_depth, inv_depth_orig=g_inv_depth_orig,
                   raydrop_prob=g_raydrop_prob, raydrop_mask=imgs["raydrop_mask"])
        return out, loss

    def step_1st(self, step):
        """demo 199-213: one stage-1 iteration; returns (outputs, per-sample loss)."""
        if self.optim_1st is None:
            self.G.eval().requires_grad_(False)
            cls = SphericalOptimizer if self.hypersphere_z else torch.optim.Adam
            self.optim_1st = cls(params=self.params_1st, lr=self.lr_1st)
            self.scheduler = torch.optim.lr_scheduler.LambdaLR(
                self.optim_1st, lr_lambda=lambda i: lr_schedule(i, self.num_steps_1st, self.rampup,
                                                                self.rampdown))
        out, loss = self.forward(progress=step / self.num_steps_1st)
        self.optim_1st.zero_grad(set_to_none=True)
        loss.backward(gradient=torch.ones_like(loss))
        self.optim_1st.step()
        self.scheduler.step()
        normalize_noise_(self.noises)
        return out, loss.detach()

    def step_2nd(self, step):
        """demo 239-247: one pivotal-tuning iteration (generator weights trainable)."""
        if self.optim_2nd is None:
            self.G.requires_grad_(True)
            self.optim_2nd = torch.optim.Adam(params=self.G.parameters(), lr=self.lr_2nd)
            self.perturb_z = False
        out, loss = self.forward(progress=step / self.num_steps_2nd)
        self.optim_2nd.zero_grad(set_to_none=True)
        loss.backward(gradient=torch.ones_like(loss))
        self.optim_2nd.step()
        normalize_noise_(self.noises)
        return out, loss.detach()

    def run(self, callback=None):
        out = loss = None
        for step in range(self.num_steps_1st):
            out, loss = self.step_1st(step)
            if callback is not None:
                callback(1, step, out, loss)
        for step in range(self.num_steps_2nd):
            out, loss = self.step_2nd(step)
            if callback is not None:
                callback(2, step, out, loss)
        return out, loss

    @torch.no_grad()
    def point_cloud(self, out, key="inv_depth", as_set=True):
        """range -> point projection of a result (gans/coords.py:139-155); the integer count of
        valid points is left in `self.last_valid_count`."""
        pts = self.coord.convert(out[key], "inv_depth_norm", "point_set" if as_set else "point_map")
        self.last_valid_count = self.coord.last_valid_count
        return pts
