"""Drop-in for gans/models/ops/style.py: ModConv2d (1x1) and NoiseInjection.

ModConv2d.forward(x, style) keeps the reference contract (style.py:68-126) but never
builds weight[B,O,I,1,1] x activations as a grouped conv: the modulation / demodulation /
EMA normaliser are folded into small per-sample matrices wb[B,O,I] (fp32 math), and the
heavy part is one batched contraction kernel (dusty_modconv_fwd) that
  * reads its K axis from two tensors (features and Fourier features) so torch.cat of the
    512 positional channels never happens,
  * lets the Fourier operand be shared by the whole batch,
  * applies bias + leaky-ReLU in its epilogue (FusedLeakyReLU fused).
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from .... import functional as DF
from .common import EqualLR


class ModConv2d(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, mod_ch: int, ksize: int = 3, stride: int = 1,
                 padding: int = 1, demod: bool = True, bias: bool = True, gain: float = 1.0,
                 transposed: bool = False, factorization_rank=None, ema=False,
                 ema_decay=0.9989):
        super().__init__()
        self.in_ch, self.out_ch, self.mod_ch = in_ch, out_ch, mod_ch
        self.ksize, self.stride, self.padding = _pair(ksize), _pair(stride), _pair(padding)
        if self.ksize != (1, 1) or self.stride != (1, 1) or self.padding != (0, 0):
            raise NotImplementedError(
                "dusty_b200 ModConv2d implements the 1x1 / stride 1 / padding 0 case, the only one "
                "dusty_v2 instantiates (dusty_v2.py:42-51,112-120)")
        if transposed or factorization_rank is not None:
            raise NotImplementedError("transposed / factorised ModConv2d are not on the hot path")
        self.weight = nn.Parameter(torch.randn((1, out_ch, in_ch, 1, 1)))
        self.transposed = transposed
        self.bias = nn.Parameter(torch.zeros((1, out_ch, 1, 1))) if bias else None
        self.gain = gain
        self.scale = 1.0 / np.sqrt(in_ch)
        self.factorization_rank = None
        self.mod = EqualLR(nn.Linear(mod_ch, in_ch), gain=1.0)
        self.demod = demod
        self.ema = ema
        self.ema_decay = ema_decay
        self.register_buffer("ema_var", torch.tensor(1.0))

    # -- small fp32 tensors: [B, O, I] at most 64 MiB for the widest layer, usually < 1 MiB
    def effective_weights(self, style, out_dtype=torch.float32, pe_rot=None, c1=0, late_ema=False,
                          via_handle=False):
        """wb[B,O,I]: one fused kernel pair (dusty_modprep_fwd/bwd) on CUDA.  pe_rot [B, 2F]
        rotates the Fourier columns (batch-shared Fourier block under an azimuth shift).
        late_ema: leave the EMA normaliser out of wb (the contraction's epilogue applies it), so
        that wb depends on the style and the parameters only."""
        s = self.mod(style.float())
        if s.is_cuda:
            if late_ema and self.ema:
                return DF.modprep(s, self.weight, None, self.scale, self.demod, out_dtype, pe_rot, c1,
                                  ema_late=self.ema_var, via_handle=via_handle)
            return DF.modprep(s, self.weight, self.ema_var if self.ema else None, self.scale,
                              self.demod, out_dtype, pe_rot, c1, via_handle=via_handle)
        if pe_rot is not None:
            raise RuntimeError("pe_rot needs the CUDA path")
        return self._effective_weights_composite(s).to(out_dtype)

    def _effective_weights_composite(self, s):
        """The same algebra in plain tensor ops (host-side reference of the fused kernel; used
        by the tests and for shape inference on the CPU, never on the training path)."""
        w = self.weight.float().reshape(self.out_ch, self.in_ch) * self.scale
        if self.demod:
            w = w / w.abs().amax()
            s = s / s.abs().amax(dim=1, keepdim=True)
        wb = w.unsqueeze(0) * (s + 1.0).unsqueeze(1)
        if self.demod:
            wb = wb * torch.rsqrt(wb.square().sum(dim=2, keepdim=True) + 1e-8)
        if self.ema:
            wb = wb / (torch.sqrt(self.ema_var) + 1e-8)
        return wb

    @torch.no_grad()
    def update_ema(self, x, pe=None, x_sumsq=None):
        """ema_var <- lerp(ema_var, mean(x^2), 1 - decay) over the *whole* input, Fourier
        channels included (style.py:99-102): two reductions + one scalar kernel (`x_sumsq`:
        sum(x^2) when the kernel that produced x already accumulated it)."""
        sa = x_sumsq if x_sumsq is not None else (DF.sumsq_buffer(x) if x is not None else None)
        numel = x.numel() if x is not None else 0
        sb, rep = None, 1
        if pe is not None:
            B = x.shape[0] if x is not None else pe.shape[0]
            rep = B // pe.shape[0]
            sb = DF.sumsq_buffer(pe)
            numel += pe.numel() * rep
        DF.ema_lerp_(self.ema_var, sa, sb, rep, numel, 1 - self.ema_decay)

    def ema_in_epilogue(self, src, c1, c2) -> bool:
        """Can this layer's EMA normaliser be applied by the contraction's epilogue?  (tcgen05
        kernels only: the bf16 production path.)"""
        return src.is_cuda and self.ema_in_epilogue_for(src.dtype, src.shape[-2] * src.shape[-1], c1, c2)

    def ema_in_epilogue_for(self, dtype, P, c1, c2) -> bool:
        return bool(self.ema and DF.late_ema_enabled()
                    and DF.modconv_tc_domain_of(dtype, self.out_ch, c1, c2, P))

    def forward(self, x, style, pe=None, fused_act=None, pe_rot=None, x_sumsq=None, wb=None,
                want_sumsq=False):
        """x: [B, C1, H, W] (or None when the input is `pe` alone); pe: optional Fourier
        block [B or 1, C2, H, W] appended on the channel axis; fused_act: a FusedLeakyReLU
        module to apply in the epilogue; wb: weights prepared ahead by
        effective_weights(..., late_ema=True) (SynthesisNetwork's weight bank)."""
        c1 = 0 if x is None else x.shape[1]
        c2 = 0 if pe is None else pe.shape[1]
        if c1 + c2 != self.in_ch:
            raise RuntimeError(f"expected {self.in_ch} input channels, got {c1}+{c2}")
        if self.ema and self.training:
            if x_sumsq is None and x is not None:
                x_sumsq = getattr(x, "_dusty_sumsq", None)     # left by the producing contraction
            self.update_ema(x, pe, x_sumsq)
        src = x if x is not None else pe
        late = wb is not None or self.ema_in_epilogue(src, c1, c2)
        if wb is None:
            wb = self.effective_weights(style, src.dtype, pe_rot, c1, late_ema=late, via_handle=True)
        bias = self.bias
        act, alpha, scale = 1, 0.0, float(self.gain)
        if fused_act is not None:
            assert bias is None and self.gain == 1.0
            bias, act = fused_act.bias, 3
            alpha, scale = float(fused_act.negative_slope), float(fused_act.scale)
        elif bias is not None and self.gain != 1.0:
            bias = bias * self.gain          # (h + b) * gain == h*gain + b*gain
        return DF.modconv_bmm(wb, x, pe, bias, act, alpha, scale,
                              ema_var=self.ema_var if (late and self.ema) else None,
                              want_sumsq=want_sumsq and self.training)

    def extra_repr(self):
        return (f"in_ch={self.in_ch}, out_ch={self.out_ch}, mod_ch={self.mod_ch}, "
                f"ksize={self.ksize}, demod={self.demod}, gain={self.gain}")


class NoiseInjection(nn.Module):
    """reference style.py:136-160 (unused: use_noise false in every shipped config)."""

    def __init__(self, ch: int = 1):
        super().__init__()
        self.ch = ch
        self.weight = nn.Parameter(torch.zeros(1, self.ch, 1, 1))
        self.fixed_noise = None

    def forward(self, x):
        B, _, H, W = x.shape
        noise = (torch.randn((B, 1, H, W), device=x.device, dtype=x.dtype)
                 if self.fixed_noise is None else self.fixed_noise.expand(B, -1, -1, -1))
        return x + self.weight.to(x.dtype) * noise

    def extra_repr(self):
        return f"ch={self.ch}"
