"""Drop-in for gans/augment/adaptive_augment.py: AdaptiveAugment (reference 294-623) for
the policy set the shipped configs enable (geometric + colour; imgfilter / noise / cutout
are disabled in every config and raise here).

Execution differences from the reference (results are the same function of the sampled
transform):
  * the per-sample transforms are sampled on the HOST (tiny [B,3,3] / [B,4,4] tensors) and
    uploaded, so the data-dependent padding is known without a device sync (the reference
    syncs in get_padding via .item(), adaptive_augment.py:289,486-487);
  * circular/reflect padding and the four SYM6 up/down passes are dusty_fir2d launches with
    fused boundary handling; backward and double backward use the adjoint kernel;
  * `p` keeps a host mirror refreshed in update_p() (one sync every `lazy.ada` iterations).
"""
import math
import os

import numpy as np
import torch
import torch.distributed as dist
from torch import autograd
from torch.nn import functional as F

from ... import _cabi as K
from ... import functional as DF
from ..models.ops.upfirdn2d.upfirdn2d import upfirdn2d

SYM2 = (-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025)
SYM6 = (0.015404109327027373, 0.0034907120842174702, -0.11799011114819057,
        -0.048311742585633, 0.4910559419267466, 0.787641141030194, 0.3379294217276218,
        -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148)


def reduce_sum(tensor):
    if not (dist.is_available() and dist.is_initialized()):
        return tensor
    tensor = tensor.clone()
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


# ---- host-side transform sampling -------------------------------------------------------
def _eye(n, size):
    return torch.eye(n).repeat(size, 1, 1)


def _scale2d(sx, sy):
    m = _eye(3, sx.shape[0])
    m[:, 0, 0], m[:, 1, 1] = sx, sy
    return m


def _translate2d(tx, ty):
    m = _eye(3, tx.shape[0])
    m[:, 0, 2], m[:, 1, 2] = tx, ty
    return m


def _single(rows):
    return torch.tensor(rows, dtype=torch.float32)


def _gate(p, candidate, prev, gen):
    """With probability p (per sample) compose `candidate` onto `prev`."""
    n = candidate.shape[0]
    pick = torch.empty(n).bernoulli_(min(max(p, 0.0), 1.0), generator=gen).view(n, 1, 1)
    ident = torch.eye(candidate.shape[-1]).expand_as(candidate)
    return (pick * candidate + (1 - pick) * ident) @ prev


def _lognormal(n, std, gen):
    return torch.empty(n).log_normal_(mean=0, std=std, generator=gen)


def _choice01(n, gen):
    return torch.randint(0, 2, (n,), generator=gen).float()


def padding_for(G_inv, height, width, kernel_size):
    """reference get_padding (271-291), evaluated on host tensors -> python ints."""
    cx, cy = (width - 1) / 2, (height - 1) / 2
    corners = _single([(-cx, -cy, 1), (cx, -cy, 1), (cx, cy, 1), (-cx, cy, 1)])
    cp = G_inv @ corners.T
    pad_k = kernel_size // 4
    pad = cp[:, :2, :].permute(1, 0, 2).flatten(1)
    pad = torch.cat((-pad, pad)).max(1).values
    pad = pad + _single([pad_k * 2 - cx, pad_k * 2 - cy] * 2)
    pad = pad.max(_single([0, 0] * 2)).min(_single([width - 1, height - 1] * 2))
    x1, y1, x2, y2 = (int(v) for v in pad.ceil().to(torch.int32))
    return x1, x2, y1, y2


class AdaptiveAugment(torch.nn.Module):
    def __init__(self, p_init=0.0, p_target=0.6, p_max=0.9, kimg=500, lr_flip=0.0, ud_flip=0.0,
                 int_trans=0.0, iso_scale=0.0, frac_trans=0.0, brightness=0.0, contrast=0.0,
                 luma_flip=0.0, hue=0.0, saturation=0.0, imgfilter=0.0, noise=0.0, cutout=0.0,
                 **ada_kwargs):
        super().__init__()
        self.register_buffer("p", torch.tensor(p_init).float())
        self.register_buffer("sign_cum", torch.zeros(1))
        self.register_buffer("n_pred_cum", torch.zeros(1))
        self.kimg = kimg * 1000
        self.p_target, self.p_max = p_target, p_max
        for name, val in dict(lr_flip=lr_flip, ud_flip=ud_flip, int_trans=int_trans,
                              iso_scale=iso_scale, frac_trans=frac_trans, brightness=brightness,
                              contrast=contrast, luma_flip=luma_flip, hue=hue,
                              saturation=saturation, imgfilter=imgfilter, noise=noise,
                              cutout=cutout).items():
            setattr(self, "mul_" + name, float(val))
        if self.mul_imgfilter > 0 or self.mul_noise > 0 or self.mul_cutout > 0:
            raise NotImplementedError("imgfilter / noise / cutout are disabled in all shipped configs")
        self.h_trans_factor = 0.0 if ada_kwargs.get("wonly_trans", False) else 1.0
        self.register_buffer("Hz_fbank", torch.as_tensor(self._filter_bank(), dtype=torch.float32))
        self._p_host = float(p_init)
        self.generator = None           # optional torch.Generator (CPU) for reproducible draws
        # one-channel images + axis-aligned policies: the whole augmentation as one kernel, the
        # transforms drawn on the device (csrc/ada_fused.cu); no host work, graph-capturable
        self.fused = os.environ.get("DUSTY_ADA_FUSED", "1") != "0"
        self._seed = None
        self.register_buffer("_calls", torch.zeros(1, dtype=torch.int64), persistent=False)

    @staticmethod
    def _filter_bank():
        lo = np.asarray(SYM2)
        hi = lo * ((-1) ** np.arange(lo.size))
        lo2, hi2 = np.convolve(lo, lo[::-1]) / 2, np.convolve(hi, hi[::-1]) / 2
        bank = np.eye(4, 1)
        for i in range(1, 4):
            bank = np.dstack([bank, np.zeros_like(bank)]).reshape(4, -1)[:, :-1]
            bank = np.stack([np.convolve(row, lo2) for row in bank])
            mid = bank.shape[1]
            bank[i, (mid - hi2.size) // 2: (mid + hi2.size) // 2] += hi2
        return bank

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        super()._load_from_state_dict(state_dict, prefix, *a, **k)
        if prefix + "p" in state_dict:
            self._p_host = float(state_dict[prefix + "p"])

    # -- ADA controller (reference 368-384)
    def cumulate(self, y_real):
        self.sign_cum += y_real.detach().sign().sum()
        self.n_pred_cum += len(y_real)

    def update_p(self):
        self.sign_cum = reduce_sum(self.sign_cum)
        self.n_pred_cum = reduce_sum(self.n_pred_cum)
        rt = self.sign_cum / self.n_pred_cum
        if self.p_target is not None:
            adjust = torch.sign(rt - self.p_target) * self.n_pred_cum / self.kimg
            self.p = (self.p + adjust).clamp_(0, self.p_max).reshape(())
        self.sign_cum *= 0
        self.n_pred_cum *= 0
        self._p_host = None                # read back lazily, only if the host samplers run
        return rt

    def _p(self) -> float:
        if self._p_host is None:
            self._p_host = float(self.p)   # the one host sync of the controller (host sampling only)
        return self._p_host

    # -- sampling (reference 386-469), on the host
    def sample_affine(self, size, height, width, device="cpu"):
        g, p = self.generator, self._p()
        G = _eye(3, size)
        ones = torch.ones(size)
        if self.mul_lr_flip > 0:
            G = _gate(p * self.mul_lr_flip, _scale2d(1 - 2.0 * _choice01(size, g), ones), G, g)
        if self.mul_ud_flip > 0:
            G = _gate(p * self.mul_ud_flip, _scale2d(ones, 1 - 2.0 * _choice01(size, g)), G, g)
        if self.mul_int_trans > 0:
            t = torch.empty(2, size).uniform_(-0.125, 0.125, generator=g)
            G = _gate(p * self.mul_int_trans,
                      _translate2d(torch.round(t[1] * width),
                                   torch.round(t[0] * height) * self.h_trans_factor), G, g)
        if self.mul_iso_scale > 0:
            s = _lognormal(size, 0.2 * math.log(2), g)
            G = _gate(p * self.mul_iso_scale, _scale2d(torch.ones_like(s), s), G, g)
        if self.mul_frac_trans > 0:
            t = torch.empty(2, size).normal_(0, 0.125, generator=g)
            G = _gate(p * self.mul_frac_trans,
                      _translate2d(t[1] * width, t[0] * height * self.h_trans_factor), G, g)
        return G.to(device)

    def sample_color(self, size, device="cpu"):
        g, p = self.generator, self._p()
        C = _eye(4, size)
        v = 1 / math.sqrt(3)
        axis = torch.tensor([v, v, v, 0.0])
        outer = torch.outer(axis, axis)
        if self.mul_brightness > 0:
            b = torch.empty(size).normal_(0, 0.2, generator=g)
            m = _eye(4, size)
            m[:, :3, 3] = b[:, None]
            C = _gate(p * self.mul_brightness, m, C, g)
        if self.mul_contrast > 0:
            c = _lognormal(size, 0.5 * math.log(2), g)
            m = _eye(4, size)
            m[:, 0, 0] = m[:, 1, 1] = m[:, 2, 2] = c
            C = _gate(p * self.mul_contrast, m, C, g)
        if self.mul_luma_flip > 0:
            i = _choice01(size, g)
            C = _gate(p * self.mul_luma_flip, _eye(4, size) - 2 * outer * i.view(-1, 1, 1), C, g)
        if self.mul_hue > 0:
            th = torch.empty(size).uniform_(-math.pi, math.pi, generator=g)
            cross = torch.tensor([(0, -v, v), (v, 0, -v), (-v, v, 0)])
            sin_t, cos_t = th.sin().view(-1, 1, 1), th.cos().view(-1, 1, 1)
            rot = cos_t * torch.eye(3) + sin_t * cross + (1 - cos_t) * outer[:3, :3]
            m = _eye(4, size)
            m[:, :3, :3] = rot
            C = _gate(p * self.mul_hue, m, C, g)
        if self.mul_saturation > 0:
            s = _lognormal(size, 1 * math.log(2), g)
            C = _gate(p * self.mul_saturation,
                      outer + (torch.eye(4) - outer) * s.view(-1, 1, 1), C, g)
        return C.to(device)

    # -- the deterministic part: image, inverse transform, colour matrix -> image
    def policy_vector(self):
        return [self.mul_lr_flip, self.mul_ud_flip, self.mul_int_trans, self.mul_iso_scale,
                self.mul_frac_trans, self.mul_brightness, self.mul_contrast, self.mul_luma_flip,
                self.mul_hue, self.mul_saturation, self.h_trans_factor]

    @staticmethod
    def fused_params(G_inv, C):
        """[B, 8] parameter rows of the fused kernel from an inverse transform / colour matrix
        pair, or None when a transform is not axis-aligned."""
        G_inv, C = G_inv.detach().float().cpu(), C.detach().float().cpu()
        if bool((G_inv[:, 0, 1] != 0).any()) or bool((G_inv[:, 1, 0] != 0).any()):
            return None
        Cm = C[:, :3, :].mean(dim=1)
        out = torch.zeros(G_inv.shape[0], 8)
        out[:, 0], out[:, 1] = G_inv[:, 0, 0], G_inv[:, 0, 2]
        out[:, 2], out[:, 3] = G_inv[:, 1, 1], G_inv[:, 1, 2]
        out[:, 4], out[:, 5] = Cm[:, :3].sum(dim=1), Cm[:, 3]
        return out

    def apply(self, img, G_inv, C):
        """img [B,C,H,W] fp32 CUDA; G_inv [B,3,3], C [B,4,4] host or device tensors."""
        img = img.float()
        device = img.device
        B, ch, H, W = img.shape
        if self.fused and DF.ada_fused_supported(img):
            params = self.fused_params(G_inv, C)
            if params is not None:
                return DF.ada_apply(img, params.to(device, non_blocking=True))
        nk = len(SYM6)
        k = DF.device_taps([list(SYM6)], device)[0]
        k_flip = DF.device_taps([list(SYM6[::-1])], device)[0]
        G_inv = G_inv.detach().float().cpu()
        px1, px2, py1, py2 = padding_for(G_inv, H, W, nk)
        pad_cfg = DF.FirCfg(1, 1, pad=(py1, py2, px1, px2), mode=(K.PAD_REFLECT, K.PAD_CIRCULAR))
        img = DF.fir2d(img, DF.device_taps([[1.0]], device), pad_cfg)
        G_inv = _single([(1, 0, (px1 - px2) / 2), (0, 1, (py1 - py2) / 2), (0, 0, 1)]) @ G_inv

        up0, up1 = (nk + 1) // 2, (nk - 2) // 2
        img = upfirdn2d(img, k[None], up=(2, 1), pad=(up0, up1, 0, 0))
        img = upfirdn2d(img, k[:, None], up=(1, 2), pad=(0, 0, up0, up1))
        S2, S2i = _single([(2, 0, 0), (0, 2, 0), (0, 0, 1)]), _single([(.5, 0, 0), (0, .5, 0), (0, 0, 1)])
        Tm, Tp = (_single([(1, 0, -.5), (0, 1, -.5), (0, 0, 1)]),
                  _single([(1, 0, .5), (0, 1, .5), (0, 0, 1)]))
        G_inv = Tm @ (S2 @ G_inv @ S2i) @ Tp

        pad_k = nk // 4
        shape = (B, ch, (H + pad_k * 2) * 2, (W + pad_k * 2) * 2)
        G_inv = (_single([(2 / img.shape[3], 0, 0), (0, 2 / img.shape[2], 0), (0, 0, 1)]) @ G_inv
                 @ _single([(shape[3] / 2, 0, 0), (0, shape[2] / 2, 0), (0, 0, 1)]))
        theta = G_inv[:, :2, :].contiguous().to(device, non_blocking=True)
        img = DF.affine_warp(img, theta, shape[2:])      # affine_grid + grid_sample, fused

        d_p = -pad_k * 2
        dn0, dn1 = d_p + (nk - 1) // 2, d_p + (nk - 2) // 2
        img = upfirdn2d(img, k_flip[None], down=(2, 1), pad=(dn0, dn1, 0, 0))
        img = upfirdn2d(img, k_flip[:, None], down=(1, 2), pad=(0, 0, dn0, dn1))

        C = C.detach().float()
        img = img.reshape(B, ch, H * W)
        if ch == 3:
            C = C.to(device, non_blocking=True)
            img = C[:, :3, :3] @ img + C[:, :3, 3:]
        elif ch == 1:
            Cm = C[:, :3, :].mean(dim=1, keepdim=True)
            gain = Cm[:, :, :3].sum(dim=2, keepdim=True).to(device, non_blocking=True)
            offs = Cm[:, :, 3:].to(device, non_blocking=True)
            img = img * gain + offs
        else:
            raise RuntimeError("AdaptiveAugment supports 1- or 3-channel images")
        return img.reshape(B, ch, H, W)

    def forward(self, img):
        if not img.is_cuda:
            raise RuntimeError("AdaptiveAugment runs on CUDA tensors only (no CPU fallback)")
        B, _, H, W = img.shape
        overridden = "sample_affine" in self.__dict__ or "sample_color" in self.__dict__   # pinned draws
        if self.fused and not overridden and DF.ada_fused_supported(img):
            # transforms drawn on the device: no host tensors, no sync, CUDA-graph safe
            params = torch.empty(B, 8, device=img.device, dtype=torch.float32)
            if self._seed is None:
                self._seed = int(self.generator.initial_seed() if self.generator is not None
                                 else torch.initial_seed())
            DF.ada_sample(params, self.p.reshape(1), self._seed, self._calls, H, W, self.policy_vector())
            return DF.ada_apply(img, params)
        G_inv = torch.inverse(self.sample_affine(B, H, W))
        return self.apply(img, G_inv, self.sample_color(B))
