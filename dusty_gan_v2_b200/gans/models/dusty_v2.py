"""Drop-in for gans/models/dusty_v2.py (reference 13-396): MappingNetwork, Head,
SynthesisBlock, SynthesisNetwork, Generator, ResidualBlock, Discriminator with identical
constructor arguments and state_dict keys (SURVEY.md section 8a footnote).

What differs is the execution plan of a synthesis block:
  reference:  resample -> FourierFeature -> torch.cat -> ~8 ATen kernels building
              weight[B,O,I] -> grouped cuDNN conv -> fused_bias_act  (x2) -> two heads
  here:       one FIR kernel (up2), one Fourier kernel (batch-shared when the angle grid
              is), then per conv ONE contraction kernel reading [features | Fourier] as two
              K ranges with bias+lrelu in its epilogue; the two 1-channel heads run as one
              O=2 contraction.
Precision: blocks flagged `use_fp16` by the reference run in the package's activation
dtype (bf16 in production, fp32 in parity mode); head outputs and everything after them
are fp32 like the reference (dusty_v2.py:174-178).
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from ... import functional as DF
from . import base, dusty_v1, ops


class MappingNetwork(nn.Sequential):
    def __init__(self, in_ch, out_ch, depth=2):
        self.in_ch, self.out_ch, self.depth = in_ch, out_ch, depth
        stages, ch = [ops.PixelNorm()], in_ch
        for _ in range(depth):
            stages.append(nn.Sequential(
                ops.EqualLR(nn.Linear(ch, out_ch), gain=math.sqrt(2), lr_mul=0.01),
                nn.LeakyReLU(negative_slope=0.2)))
            ch = out_ch
        super().__init__(*stages)


class _HeadOut(dict):
    """name -> [B,1,H,W] views of one stacked [B,n,H,W] tensor (kept for the next level)."""
    stacked = None


class Head(nn.Module):
    def __init__(self, in_ch, mod_ch, out_ch):
        super().__init__()
        self.in_ch, self.mod_ch, self.out_ch = in_ch, mod_ch, out_ch
        self.heads = nn.ModuleDict()
        for o in out_ch:
            if o["ch"] == 0:
                continue
            if o["ch"] != 1:
                raise NotImplementedError("heads with more than one channel")
            self.heads[o["name"]] = ops.ModConv2d(out_ch=o["ch"], in_ch=in_ch, mod_ch=mod_ch,
                                                  ksize=1, stride=1, padding=0, demod=False,
                                                  ema=True)

    def late_ema_ok(self, is_cuda=True) -> bool:
        """Per-row EMA normalisers applied by the small-O contraction kernels (weights then depend
        on the style and the parameters only: SynthesisNetwork's weight bank)."""
        mods = list(self.heads.values())
        return bool(is_cuda and DF.late_ema_enabled() and 1 <= len(mods) <= 4 and all(m.ema for m in mods)
                    and DF._PRECISION["modconv_impl"] in (0, 1) and self.in_ch * len(mods) * 4 <= 48 * 1024)

    def prepare_weights(self, style, dtype, side):
        mods = list(self.heads.values())
        wb = DF.cat_wb([m.effective_weights(style, dtype, late_ema=True, via_handle=True) for m in mods])
        event = torch.cuda.Event()
        event.record(side)
        return wb, event

    def forward(self, x, style, ready=None):
        """All heads share x and the style: one contraction with O = number of heads."""
        mods = list(self.heads.values())
        if self.training:
            total = getattr(x, "_dusty_sumsq", None)
            if total is None:
                total = DF.sumsq_buffer(x)
            for m in mods:
                DF.ema_lerp_(m.ema_var, total, None, 1, x.numel(), 1 - m.ema_decay)
        bias = torch.cat([m.bias.reshape(-1) for m in mods])
        if ready is not None:
            wb, event = ready
            main = torch.cuda.current_stream()
            main.wait_event(event)
            wb.record_stream(main)
            y = DF.modconv_bmm(wb, x, None, bias, 1, 0.0, 1.0, ema_rows=[m.ema_var for m in mods])
        elif self.late_ema_ok(x.is_cuda):
            # the same formulation as the weight bank's (results do not depend on the bank)
            wb = DF.cat_wb([m.effective_weights(style, x.dtype, late_ema=True, via_handle=True) for m in mods])
            y = DF.modconv_bmm(wb, x, None, bias, 1, 0.0, 1.0, ema_rows=[m.ema_var for m in mods])
        else:
            wb = DF.cat_wb([m.effective_weights(style, x.dtype, via_handle=x.is_cuda) for m in mods])
            y = DF.modconv_bmm(wb, x, None, bias, 1, 0.0, 1.0)
        out = _HeadOut()
        out.stacked = y
        for i, name in enumerate(self.heads.keys()):
            out[name] = y[:, i:i + 1]
        return out


class SynthesisBlock(nn.Module):
    def __init__(self, in_ch, mid_ch, out_ch, mod_ch, resolution, up=2, resample_dir="hw",
                 resample_window=[1, 3, 3, 1], use_noise=True, use_pe=True, pe_type="random",
                 pe_ch=512, pe_scale_offset=(3, -1), ring=True):
        super().__init__()
        self.use_pe = use_pe
        self.use_fp16 = False
        self.is_first = in_ch == 0
        self.num_conv = 0
        if up > 1:
            self.resample = ops.Resample(up=up, window=resample_window, ring=ring,
                                         direction=resample_dir)
            self.downsample = ops.Resample(down=up, window=resample_window, ring=ring,
                                           direction=resample_dir)
        else:
            self.resample = nn.Identity()
            self.downsample = None
        if use_pe:
            self.pe = ops.FourierFeature(resolution=resolution, basis_scale=pe_type,
                                         num_freqs=pe_ch, L_offset=pe_scale_offset)
            pe_ch = self.pe.out_ch
        else:
            pe_ch = 0
        kw = dict(out_ch=mid_ch, mod_ch=mod_ch, ksize=1, stride=1, padding=0, bias=False, ema=True)
        self.conv1 = ops.ModConv2d(in_ch=in_ch + pe_ch, **kw)
        self.noise1 = ops.NoiseInjection() if use_noise else None
        self.bias_act1 = ops.FusedLeakyReLU(mid_ch)
        self.num_conv += 1
        if not self.is_first:
            self.conv2 = ops.ModConv2d(in_ch=mid_ch, **kw)
            self.noise2 = ops.NoiseInjection() if use_noise else None
            self.bias_act2 = ops.FusedLeakyReLU(mid_ch)
            self.num_conv += 1
        self.head = Head(mid_ch, mod_ch, out_ch)

    def downsample_angle(self, angle):
        if (isinstance(self.downsample, ops.Resample) and self.downsample.window == [1, 3, 3, 1]
                and self.downsample.ring and self.downsample.direction == "hw"
                and self.downsample.down_h == 2 and angle.shape[1] == 2):
            return DF.angle_down2(angle)          # fused sin/cos -> FIR -> atan2
        c = angle.shape[1]
        per = self.downsample(torch.cat([angle.sin(), angle.cos()], dim=1))
        return torch.atan2(per[:, :c], per[:, c:])

    def _conv(self, conv, noise, act, h, style, pe=None, pe_rot=None, x_sumsq=None, ready=None):
        wb = None
        if ready is not None:                  # weights prepared ahead on the side stream
            wb, event = ready
            main = torch.cuda.current_stream()
            main.wait_event(event)
            # wb was allocated on the side stream and is read here (and, saved by autograd, in this
            # stream's backward): its memory must not return to the side stream's pool before
            wb.record_stream(main)
        if noise is None:
            # the output feeds another ModConv2d (conv2 / the heads): its EMA statistic comes out
            # of this contraction's epilogue
            return conv(h, style, pe=pe, fused_act=act, pe_rot=pe_rot, x_sumsq=x_sumsq, wb=wb,
                        want_sumsq=True)
        return act(noise(conv(h, style, pe=pe, pe_rot=pe_rot, x_sumsq=x_sumsq, wb=wb)))

    def prepare_weights(self, ws, dtype, P, shift_rad, side):
        """Per-sample weights of conv1 / conv2 for this block, computed on the stream `side`
        (they depend on the styles and the parameters only once the EMA normaliser lives in the
        contraction epilogue).  Returns {name: (wb, event)} for the layers that qualify."""
        out = {}
        pe_ch = self.pe.out_ch if self.use_pe else 0
        plan = [("conv1", self.conv1, ws[0], self.conv1.in_ch - pe_ch, pe_ch)]
        if not self.is_first:
            plan.append(("conv2", self.conv2, ws[1], self.conv2.in_ch, 0))
        for name, conv, style, c1, c2 in plan:
            if not conv.ema_in_epilogue_for(dtype, P, c1, c2):
                continue
            rot = None
            if name == "conv1" and shift_rad is not None and self.use_pe:
                rot = self.pe_rotation(shift_rad)
            wb = conv.effective_weights(style, dtype, rot, c1, late_ema=True, via_handle=True)
            event = torch.cuda.Event()
            event.record(side)
            out[name] = (wb, event)
        if self.head.late_ema_ok() and (dtype == torch.bfloat16 or dtype == torch.float32):
            out["head"] = self.head.prepare_weights(ws[-1], dtype, side)
        return out

    def pe_rotation(self, shift_rad):
        """[B, 2F] table (cos | sin) of psi[b,f] = f_w[f] * shift_b for this block's basis."""
        f_w = self.pe.freqs[:, 1].reshape(1, -1).float()
        psi = shift_rad.reshape(-1, 1).float() * f_w
        return torch.cat([psi.cos(), psi.sin()], dim=1)

    def forward(self, h, skip, ws, angle, shift_rad=None, ready=None):
        ws = iter(ws)
        ready = ready or {}
        dtype = DF.act_dtype() if (self.use_fp16 and angle.is_cuda) else torch.float32
        h_ss = None
        if h is not None:
            if self.training and self.conv1.ema and isinstance(self.resample, ops.Resample):
                # conv1's EMA statistic of the upsampled tensor comes out of the upsampling kernel
                h, h_ss = self.resample.forward_with_sumsq(h.to(dtype))
            else:
                h = self.resample(h.to(dtype))
        pe = self.pe(angle, out_dtype=dtype) if self.use_pe else None
        rot = self.pe_rotation(shift_rad) if (shift_rad is not None and pe is not None
                                              and "conv1" not in ready) else None
        h = self._conv(self.conv1, self.noise1, self.bias_act1, h, next(ws), pe, rot, h_ss,
                       ready.get("conv1"))
        if not self.is_first:
            h = self._conv(self.conv2, self.noise2, self.bias_act2, h, next(ws), ready=ready.get("conv2"))
        o = self.head(h, next(ws), ready.get("head"))
        y = o.stacked.float()
        if skip is not None:
            prev = skip.stacked if isinstance(skip, _HeadOut) else torch.cat(
                [skip[k] for k in o.keys()], dim=1)
            y = y + self.resample(prev)
        out = _HeadOut()
        out.stacked = y
        for i, k in enumerate(o.keys()):
            out[k] = y[:, i:i + 1]
        return h, out

    def extra_repr(self):
        return f"use_fp16={self.use_fp16}"


_SIDE_STREAMS = {}


def _side_stream(device) -> "torch.cuda.Stream":
    """One side stream per device for the weight bank (module-level: streams do not deep-copy)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return st


def _batch_shared(angle: torch.Tensor, cache: dict) -> bool:
    """True when every sample carries the same angle grid (the trainer / demos always pass
    CoordBridge.angle repeated B times).  The check costs one sync, so it is cached per
    (storage, version, shape)."""
    if angle.shape[0] == 1 or angle.stride(0) == 0:
        return True
    key = (angle.data_ptr(), angle._version, tuple(angle.shape))
    hit = cache.get(key)
    if hit is None:
        hit = bool((angle[1:] == angle[:1]).all().item())
        cache.clear()
        cache[key] = hit
    return hit


class SynthesisNetwork(nn.Module):
    def __init__(self, in_ch, out_ch, ch_base=64, ch_max=512, resolution=(64, 256), ring=True,
                 layers=[2, 2, 2, 2], num_fp16_layers=-1, use_noise=True, pe_type="random",
                 pe_scale_offset=(3, -1), aug_coords=True, aug_coords_blitting=False,
                 output_scale=1 / 4.0):
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        self.resolution_out = np.array(_pair(tuple(resolution) if not isinstance(resolution, int)
                                             else resolution))
        self.resolution_in = self.resolution_out // int(np.prod(layers))
        n = len(layers)
        width = [min(ch_base << (n - i), ch_max) for i in range(n + 1)]
        self.layers = nn.ModuleList()
        res = self.resolution_in.copy()
        for i, scale in enumerate([1] + list(layers)):
            res = res * scale
            self.layers.append(SynthesisBlock(
                in_ch=width[i - 1] if i else 0, mid_ch=width[i], out_ch=out_ch, mod_ch=in_ch,
                resolution=res.copy(), up=scale, resample_window=[1, 3, 3, 1],
                use_noise=use_noise, use_pe=(scale > 1 or i == 0), pe_type=pe_type,
                pe_scale_offset=pe_scale_offset, ring=ring))
        for i, blk in enumerate(reversed(self.layers)):
            blk.use_fp16 = (num_fp16_layers == -1) or (i < num_fp16_layers)
        self.num_styles = len(self.layers) * 2
        self.aug_coords = aug_coords
        self.aug_coords_blitting = aug_coords_blitting
        self.output_scale = output_scale
        acts = {}
        for o in out_ch:
            a = o["act"]
            acts[o["name"]] = nn.Identity() if a is None else (eval(a)() if isinstance(a, str) else a())
        self.output_acts = nn.ModuleDict(acts)
        self._shared_cache = {}
        self._int_freq_cache = {}
        self.shared_pe_in_training = "auto"     # "auto" (bf16 mode only) | True | False
        self.weight_bank = os.environ.get("DUSTY_WEIGHT_BANK", "1") != "0"

    def _weight_bank(self, ws, pyramid, shift_rad):
        if not (ws.is_cuda and self.weight_bank and DF.late_ema_enabled()):
            return None
        side, main = _side_stream(ws.device), torch.cuda.current_stream()
        side.wait_stream(main)
        bank, i = [], 0
        with torch.cuda.stream(side):
            for blk, ang in zip(self.layers, pyramid):
                dtype = DF.act_dtype() if blk.use_fp16 else torch.float32
                P = ang.shape[-2] * ang.shape[-1]
                styles = (ws[:, i], ws[:, i + 1]) + (() if blk.is_first else (ws[:, i + 2],))
                bank.append(blk.prepare_weights(styles, dtype, P, shift_rad, side))
                i += blk.num_conv
        if not any(bank):
            main.wait_stream(side)
            return None
        return bank

    def _can_rotate(self, angle):
        """Shared-Fourier-block training path: low-precision mode, batch-shared angle grid and
        integer horizontal frequencies at every level (true for the 'random' basis,
        fourier.py:33-36).  fp32 mode keeps the literal per-sample evaluation for parity."""
        if self.shared_pe_in_training is False or not angle.is_cuda:
            return False
        if self.shared_pe_in_training == "auto" and DF.act_dtype() == torch.float32:
            return False
        if not _batch_shared(angle, self._shared_cache):
            return False
        key = tuple(blk.pe.freqs._version for blk in self.layers if blk.use_pe)
        if self._int_freq_cache.get("key") != key:
            ok = all(bool((blk.pe.freqs[:, 1] == blk.pe.freqs[:, 1].round()).all().item())
                     for blk in self.layers if blk.use_pe) and all(blk.use_pe for blk in self.layers)
            self._int_freq_cache = {"key": key, "ok": ok}
        return self._int_freq_cache["ok"]

    @staticmethod
    def translation_matrix(t):
        t = t.div(2 * np.pi)
        mat = torch.eye(3, device=t.device)[None].repeat_interleave(t.shape[0], dim=0)
        mat[:, 0, 2] = t[:, 1]
        mat[:, 1, 2] = t[:, 0]
        return mat

    def forward(self, ws, angle):
        B, N, _ = ws.shape
        if N != self.num_styles:
            raise RuntimeError(f"{self.num_styles} != {N}")
        aug = self.training and self.aug_coords
        W = int(self.resolution_out[1])
        shift01 = None
        shift_rad = None
        if aug:
            # same RNG draw as the reference: one uniform per sample (dusty_v2.py:266-274)
            shifts = torch.zeros((B, 2), device=ws.device)
            shifts[:, 1].uniform_(0, 1)
            if self.aug_coords_blitting:
                shifts[:, 1].mul_(W).round_().div_(W)
            shift01 = shifts[:, 1].contiguous()
            if self._can_rotate(angle):
                # integer horizontal frequencies: PE(angle + shift) = Rot(f_w * shift) PE(angle),
                # so keep ONE Fourier block for the batch and rotate the per-sample weights
                angle = angle[:1]
                shift_rad = shift01 * (2 * np.pi)
            else:
                angle = angle + shifts.mul(2 * np.pi)[..., None, None]
        elif _batch_shared(angle, self._shared_cache):
            angle = angle[:1]            # one pyramid + one Fourier block for the whole batch

        pyramid = [angle]
        for blk in self.layers[:0:-1]:
            if blk.downsample is not None:
                angle = blk.downsample_angle(angle)
            pyramid.insert(0, angle)

        # Weight bank: every layer's per-sample weights are functions of the styles and the
        # parameters alone (the EMA normaliser is applied by the contraction), so they are
        # prepared up-front on a side stream and overlap with the activation chain instead of
        # sitting in it as ~40 latency-bound launches (and ~100 more on the way back: autograd
        # runs their backward on the same side stream).
        bank = self._weight_bank(ws, pyramid, shift_rad)
        h, skip, i = None, None, 0
        for bi, (blk, ang) in enumerate(zip(self.layers, pyramid)):
            h, skip = blk(h, skip, (ws[:, i], ws[:, i + 1], ws[:, i + 2]), ang, shift_rad,
                          None if bank is None else bank[bi])
            i += blk.num_conv
        if bank is not None:
            torch.cuda.current_stream().wait_stream(_side_stream(ws.device))     # rejoin (graph capture)
            del bank        # the weights stay alive through autograd only

        y = skip.stacked
        if aug:      # cancel the azimuth shift in image space, output_scale folded in
            y = DF.circular_unshift(y, shift01, self.output_scale)
        else:
            y = y * self.output_scale
        out = {}
        for j, k in enumerate(skip.keys()):
            v = y[:, j:j + 1]
            out[k] = self.output_acts[k](v) if k in self.output_acts else v
        return out


class Generator(base.Generator):
    def __init__(self, mapping_kwargs, synthesis_kwargs, measurement_kwargs):
        super().__init__(mapping_network=MappingNetwork(**mapping_kwargs),
                         synthesis_network=SynthesisNetwork(**synthesis_kwargs),
                         measurement_model=dusty_v1.RayDropModel(**measurement_kwargs))

    def forward_synthesis(self, w, angle=None):
        angle = self.angle if angle is None else angle
        return self.synthesis_network(w, angle)


class ResidualBlock(nn.Module):
    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        kw = dict(bias=False, ring=True, equal_lr=True)
        self.conv1 = ops.Conv2d(in_ch, in_ch, 3, 1, 1, **kw)
        self.bias_act1 = ops.FusedLeakyReLU(in_ch)
        self.resample = ops.Resample(window=[1, 3, 3, 1], ring=True)
        self.conv2 = ops.Conv2d(in_ch, out_ch, 3, 2, 1, **kw)
        self.bias_act2 = ops.FusedLeakyReLU(out_ch)
        self.skip = ops.Conv2d(in_ch, out_ch, 1, 2, 0, **kw)
        self.fused_fork = True          # one backward kernel for the input's two consumers (NHWC)
        self.fused_act_blur = True      # conv1 + bias_act1 + blur + pad as one autograd node

    def _blur_pad_conv2(self, h):
        """conv2(resample(h)) with the blur and conv2's ring padding as one kernel (NHWC)."""
        pad = self.conv2[0] if len(self.conv2) == 2 else None
        rs = self.resample
        if (pad is not None and isinstance(pad, ops.Pad) and pad.padding == (1, 1, 1, 1)
                and pad.horizontal == "circular" and pad.vertical == "replicate"
                and getattr(rs, "_fast_up", None) == 1 and DF.blur_pad_cl_supported(h)):
            if rs._taps_host is None:
                rs._taps_host = tuple(rs.kernel.detach().float().cpu().tolist())
            return self.conv2[1](DF.blur_pad_cl(h, rs._taps_host))
        return self.conv2(rs(h))

    def residual(self, x):
        h = self.bias_act1(self.conv1(x))
        return self.bias_act2(self._blur_pad_conv2(h))

    def _skip_fast(self, x):
        rs, eq = self.resample, self.skip[-1]
        conv = getattr(eq, "module", None)
        return (len(self.skip) == 1 and isinstance(eq, ops.EqualLR) and isinstance(conv, nn.Conv2d)
                and conv.kernel_size == (1, 1) and conv.stride == (2, 2) and conv.bias is None
                and getattr(rs, "_fast_up", None) == 1 and DF.blur_down2_cl_supported(x))

    def _skip(self, x, xd=None):
        """skip(resample(x)): the 1x1 stride-2 convolution reads the blurred image at even
        positions only, so on the NHWC path the blur is evaluated there alone and the
        convolution runs at unit stride on the quarter-size tensor (`xd`: that tensor when the
        fused input fork already produced it)."""
        rs, eq = self.resample, self.skip[-1]
        if xd is not None or self._skip_fast(x):
            if rs._taps_host is None:
                rs._taps_host = tuple(rs.kernel.detach().float().cpu().tolist())
            w, w_tco = eq.prepared_weight(x.dtype, with_tco=True)
            if xd is None:
                xd = DF.blur_down2_cl(x, rs._taps_host)
            return ops.conv2d_valid(xd, w, (1, 1), w_tco)
        return self.skip(rs(x))

    def _conv1_fast(self, x):
        seq = self.conv1
        eq = seq[-1]
        conv = getattr(eq, "module", None)
        return (len(seq) == 2 and isinstance(seq[0], ops.Pad) and isinstance(eq, ops.EqualLR)
                and isinstance(conv, nn.Conv2d) and conv.bias is None and conv.stride == (1, 1)
                and x.is_cuda and x.dtype == torch.bfloat16)

    def _conv1_act(self, x, xp=None):
        """bias_act1(conv1(x)); on the halo-resident tcgen05 kernel the bias / leaky ReLU run in
        the convolution's epilogue (`xp`: the ring-padded input when the fused input fork
        already produced it)."""
        seq, act = self.conv1, self.bias_act1
        eq = seq[-1]
        if xp is not None or self._conv1_fast(x):
            w, w_tco = eq.prepared_weight(x.dtype, with_tco=True)
            if xp is None:
                xp = seq[0](x)
            if ops.conv_bias_act_supported(xp, w, (1, 1)):
                return ops.conv_bias_act(xp, w, act.bias, (1, 1), act.negative_slope, act.scale, w_tco)
            return act(ops.conv2d_valid(xp, w, (1, 1), w_tco))
        return act(seq(x))

    def _fork(self, x):
        """(Pad(1, ring)(x), blur_down2(x)) through one autograd node, so that the two
        branches' input gradients are folded and summed by one kernel; (None, None) when
        either branch is not on its NHWC fast path."""
        pad = self.conv1[0]
        if (self.fused_fork and self._conv1_fast(x) and self._skip_fast(x)
                and pad.padding == (1, 1, 1, 1) and pad.horizontal == "circular"
                and pad.vertical == "replicate" and DF.residual_fork_supported(x)):
            rs = self.resample
            if rs._taps_host is None:
                rs._taps_host = tuple(rs.kernel.detach().float().cpu().tolist())
            return DF.residual_fork(x, rs._taps_host)
        return None, None

    def _conv1_act_blur_pad(self, x, xp):
        """pad(blur(bias_act1(conv1(x)))) as ONE autograd node on the halo-resident kernel's layers:
        bias / leaky ReLU in the convolution's epilogue, then the blur + pad kernel; backward: blur
        adjoint + ring fold + activation gate + bias gradient in one kernel.  None when the block
        is not on that path (the caller then runs the stages one by one)."""
        seq, act, rs = self.conv1, self.bias_act1, self.resample
        pad2 = self.conv2[0] if len(self.conv2) == 2 else None
        if not (self.fused_act_blur and (xp is not None or self._conv1_fast(x))
                and isinstance(pad2, ops.Pad) and pad2.padding == (1, 1, 1, 1)
                and pad2.horizontal == "circular" and pad2.vertical == "replicate"
                and getattr(rs, "_fast_up", None) == 1):
            return None
        eq = seq[-1]
        if xp is None:
            xp = seq[0](x)
        w_like = torch.empty(eq.module.weight.shape, dtype=xp.dtype, device="meta")
        out_like = x if x.shape[1] == w_like.shape[0] else None
        if not (ops.conv_bias_act_supported(xp, w_like, (1, 1)) and out_like is not None
                and DF.blur_pad_cl_supported(out_like)):
            return None
        if rs._taps_host is None:
            rs._taps_host = tuple(rs.kernel.detach().float().cpu().tolist())
        w, w_tco = eq.prepared_weight(x.dtype, with_tco=True)
        return ops.conv_bias_act(xp, w, act.bias, (1, 1), act.negative_slope, act.scale, w_tco,
                                 blur_taps=rs._taps_host)

    def forward(self, x):
        xp, xd = self._fork(x)
        skip = self._skip(x, xd)
        hp = self._conv1_act_blur_pad(x, xp)
        if hp is not None:
            pre = self.conv2[1](hp)
        else:
            pre = self._blur_pad_conv2(self._conv1_act(x, xp))        # conv2 output, before bias_act2
        act = self.bias_act2
        if DF.residual_tail_supported(pre, skip):
            # bias_act2, the residual sum and the 1/sqrt(2) as one NHWC pass
            return DF.residual_tail(pre, act.bias, skip, act.negative_slope, act.scale,
                                    1.0 / math.sqrt(2))
        return (act(pre) + skip) * (1.0 / math.sqrt(2))


class _ActView:
    """Stand-in `self` for FusedLeakyReLU.forward with an explicit bias tensor (the composite
    form of the fused stem differentiates w.r.t. the bias it was handed)."""

    def __init__(self, bias, negative_slope, scale):
        self.bias, self.negative_slope, self.scale = bias, negative_slope, scale


class Discriminator(nn.Module):
    def __init__(self, in_ch: int, ch_base: int = 32, ch_max: int = 512, mbdis_group: int = 4,
                 mbdis_feat: int = 1, resolution=(64, 512), ring=True, num_fp16_layers=-1,
                 pre_blur=True):
        super().__init__()
        res_in = _pair(256 if resolution is None else tuple(resolution))
        n_down = int(np.log2(min(res_in) / 4))
        res_out = tuple(r >> n_down for r in res_in)
        ch = lambda i: min(ch_base << i, ch_max)
        kw = dict(bias=False, ring=ring, equal_lr=True)
        self.num_fp16_layers = num_fp16_layers
        # keep the residual trunk in NHWC on CUDA: the dense convs (library) run their sm_100
        # kernels without layout-conversion passes, and pad / blur / bias_act vectorise over C
        self.channels_last = True
        self.fused_stem = True          # BlurVH + 1x1 conv + bias/lrelu as one kernel (bf16 mode)
        self._stem_taps = None
        self.weight_bank = os.environ.get("DUSTY_WEIGHT_BANK", "1") != "0"
        self._bank_convs = None
        in_ch = in_ch * 2 if pre_blur else in_ch
        stack = [ops.BlurVH(ring=ring)] if pre_blur else []
        stack += [ops.Conv2d(in_ch, ch(0), 1, 1, 0, **kw), ops.FusedLeakyReLU(ch(0))]
        stack += [ResidualBlock(ch(i), ch(i + 1)) for i in range(n_down)]
        self.layers = nn.Sequential(*stack)
        self.epilogue = nn.Sequential(
            ops.MinibatchStdDev(group=mbdis_group, features=mbdis_feat),
            ops.Conv2d(ch(4) + mbdis_feat, ch(4), 3, 1, 1, **kw),
            ops.FusedLeakyReLU(ch(4)),
            nn.Flatten(),
            ops.EqualLR(nn.Linear(ch(4) * int(np.prod(res_out)), ch(4), bias=False)),
            ops.FusedLeakyReLU(ch(4)),
            ops.EqualLR(nn.Linear(ch(4), 1)),
        )

    def _fused_stem(self, x, low):
        """layers[0:3] = BlurVH -> 1x1 Conv2d(2 -> C0) -> FusedLeakyReLU as one kernel (bf16 NHWC
        output); None when the configuration does not match."""
        if not (self.fused_stem and len(self.layers) >= 3 and low == torch.bfloat16 and x.is_cuda):
            return None
        blur, conv, act = self.layers[0], self.layers[1], self.layers[2]
        if not (isinstance(blur, ops.BlurVH) and isinstance(conv, ops.Conv2d) and len(conv) == 1
                and isinstance(act, ops.FusedLeakyReLU) and self.num_fp16_layers in (-1,)):
            return None
        eq = conv[0]
        m = getattr(eq, "module", None)
        if not (isinstance(eq, ops.EqualLR) and isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1)
                and m.stride == (1, 1) and m.bias is None and m.in_channels == 2
                and DF.stem_supported(x, m.out_channels)):
            return None
        kv, kh = blur.blur_v, blur.blur_h
        if not (kv.window == [1, 2, 1] and kh.window == [1, 2, 1] and kv.ring and kh.ring
                and getattr(kv, "up_h", 1) == 1 and getattr(kv, "down_h", 1) == 1):
            return None
        if self._stem_taps is None:
            tv = kv.kernel.detach().float().cpu().reshape(-1).tolist()
            th = kh.kernel.detach().float().cpu().reshape(-1).tolist()
            if len(tv) != 3 or any(abs(a - b) > 1e-7 for a, b in zip(tv, th)):
                self._stem_taps = False
            else:
                self._stem_taps = tuple(tv)
        if not self._stem_taps:
            return None
        w = m.weight.reshape(m.out_channels, 2) * (eq.scale * eq.gain_)

        def composite(xc, wc, bc):
            hc = blur(xc.to(low))
            yc = ops.conv2d_valid(hc, wc.reshape(-1, 2, 1, 1).to(hc.dtype), (1, 1)) if hc.is_cuda \
                else torch.nn.functional.conv2d(hc, wc.reshape(-1, 2, 1, 1))
            return act.__class__.forward(_ActView(bc, act.negative_slope, act.scale), yc)

        return DF.stem(x, w, act.bias, self._stem_taps, composite, act.negative_slope, act.scale)

    def _prepare_conv_weights(self, device, dtype):
        """Filter bank: the scaled / cast / re-laid-out filters of every residual-block convolution
        depend on the parameters only, so they are prepared up-front on the side stream (one
        launch each, ~50 per pass with their adjoints) instead of inside the activation chain."""
        if not self.weight_bank:
            return None
        if self._bank_convs is None:
            self._bank_convs = [eq for blk in self.layers if isinstance(blk, ResidualBlock)
                                for conv in (blk.conv1, blk.conv2, blk.skip)
                                for eq in conv if isinstance(eq, ops.EqualLR) and isinstance(eq.module, nn.Conv2d)
                                and eq.module.bias is None and eq.module.padding == (0, 0)]
        side, main = _side_stream(device), torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for eq in self._bank_convs:
                eq._ready = None
                w, w_tco = eq.prepared_weight(dtype, with_tco=True)
                event = torch.cuda.Event()
                event.record(side)
                eq._ready = (w, w_tco, event)
        return side

    def _release_conv_weights(self, side):
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)        # rejoin (graph capture)
            for eq in self._bank_convs:
                eq._ready = None

    def forward(self, h):
        low = DF.act_dtype()
        y = self._fused_stem(h, low) if h.dim() == 4 and h.shape[1] == 1 else None
        if y is not None:
            h = y
            side = self._prepare_conv_weights(h.device, low)
            try:
                for layer in self.layers[3:]:
                    h = layer(h)
            finally:
                self._release_conv_weights(side)
            if self._fast_epilogue_ok(h, low):
                return self._epilogue_low_precision(h)
            return self.epilogue(h.to(torch.float32).contiguous())
        for i, layer in enumerate(self.layers):
            use_low = ((self.num_fp16_layers > i) or (self.num_fp16_layers == -1)) and h.is_cuda
            h = layer(h.to(low if use_low else torch.float32))
            if i == 0 and self.channels_last and h.is_cuda:
                h = h.contiguous(memory_format=torch.channels_last)
        if self._fast_epilogue_ok(h, low):
            return self._epilogue_low_precision(h)
        return self.epilogue(h.to(torch.float32).contiguous())

    def _fast_epilogue_ok(self, h, low):
        mb = self.epilogue[0]
        # bf16 trunk, or the fp32 mode with its contractions on the tensor cores (split-bf16
        # operands): either way the 513-channel filter would fall outside the tcgen05 domain
        low_ok = (low == torch.bfloat16 and h.dtype == torch.bfloat16) or \
                 (low == torch.float32 and h.dtype == torch.float32 and DF.fp32_on_tensor_cores()
                  and h.shape[1] % 8 == 0)
        return (h.is_cuda and low_ok and mb.features == 1
                and h.shape[0] % (mb.sub_batches * min(h.shape[0] // mb.sub_batches, mb.group)) == 0)

    def _epilogue_low_precision(self, h):
        """Same arithmetic as `self.epilogue`, arranged for the NHWC bf16 trunk: the appended
        MinibatchStdDev channel is constant over (H, W) and stays constant under the ring
        padding, so its contribution to the 3x3 convolution is stat_b * sum_{r,s} w[o, C, r, s]
        -- a per-sample bias.  The 512-channel remainder of the filter runs as an ordinary NHWC
        convolution; nothing with 513 channels (odd strides, fp32 NCHW pads) is materialised."""
        mb, conv, act1, _, lin1, act2, lin2 = self.epilogue
        pad, eq = conv[0], conv[1]
        C = h.shape[1]
        if mb.sub_batches > 1:
            stat = torch.cat([DF.minibatch_std_stat(c, mb.group) for c in h.chunk(mb.sub_batches, 0)])
        else:
            stat = DF.minibatch_std_stat(h, mb.group)
        w = eq.module.weight * (eq.scale * eq.gain_)                   # [O, C + 1, 3, 3] fp32
        w_main = w[:, :C].to(h.dtype).contiguous(memory_format=torch.channels_last)
        w_stat = w[:, C].sum(dim=(1, 2))                               # [O]
        y = ops.conv2d_valid(pad(h), w_main, eq.module.stride)
        y = y + (stat[:, None] * w_stat[None, :]).to(y.dtype)[:, :, None, None]
        y = act1(y)
        y = lin1(y.flatten(1))              # logical (c, h, w) order, as nn.Flatten on NCHW
        return lin2(act2(y.float()))
