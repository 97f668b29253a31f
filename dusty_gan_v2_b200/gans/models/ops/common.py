"""Drop-in for gans/models/ops/common.py: Pad, filter2d, Resample, BlurVH, EqualLR, Conv2d,
PixelNorm, MinibatchStdDev -- same constructor arguments, attribute and state_dict names.

All resampling / padding goes through one polyphase FIR kernel (dusty_fir2d) with the
boundary extension folded into index math: no padded or zero-inserted tensor exists.
The dense convolutions run on this package's own kernels (tcgen05 implicit GEMM for bf16 NHWC,
a CUDA-core family for fp32 / odd shapes); linears are cuBLAS GEMMs.  The EqualLR scale is folded
into the (small) weight instead of the activation.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.utils import _pair, _quadruple

from .... import _cabi as K
from .... import functional as DF


def _mode(name):
    return {"replicate": K.PAD_REPLICATE, "reflect": K.PAD_REFLECT, "circular": K.PAD_CIRCULAR,
            "zeros": K.PAD_ZERO, "constant": K.PAD_ZERO}[name]


class Pad(nn.Module):
    """reference common.py:10-24 -- W: circular when ring else `mode`; H: `mode`."""

    def __init__(self, padding, ring=False, mode="replicate"):
        super().__init__()
        self.padding = _quadruple(padding)
        self.horizontal = "circular" if ring else mode
        self.vertical = mode

    def forward(self, h):
        left, right, top, bottom = self.padding
        modes = (_mode(self.vertical), _mode(self.horizontal))
        if DF.pad2d_supported(h, (top, bottom, left, right), modes):
            return DF.pad2d(h, (top, bottom, left, right), modes)
        cfg = DF.FirCfg(1, 1, pad=(top, bottom, left, right),
                        mode=(_mode(self.vertical), _mode(self.horizontal)))
        return DF.fir2d(h, DF.device_taps([[1.0]], h.device), cfg)

    def extra_repr(self):
        return f"padding={self.padding}, horizontal={self.horizontal}, vertical={self.vertical}"


def filter2d(x, kernel, gain=1):
    """reference common.py:27-42: same-size separable blur, circular W / replicate H."""
    assert kernel.ndim == 1
    k = (kernel / kernel.sum()) * (gain ** 0.5)
    n = len(k)
    taps = torch.outer(k, k).to(device=x.device, dtype=torch.float32).contiguous()
    cfg = DF.FirCfg(n, n, pad=(n // 2, (n - 1) // 2, n // 2, (n - 1) // 2),
                    mode=(K.PAD_REPLICATE, K.PAD_CIRCULAR))
    return DF.fir2d(x, taps, cfg)


class Resample(nn.Module):
    """reference common.py:45-138.  One launch per call: the H and W passes are applied as a
    single separable 2-D tap set."""

    def __init__(self, up=1, down=1, window=[1, 3, 3, 1], ring=True, normalize=True,
                 direction="hw"):
        super().__init__()
        assert direction in ("h", "w", "hw")
        self.up = np.asarray(_pair(up))
        self.down = np.asarray(_pair(down))
        self.window = list(window)
        self.n_taps = len(window)
        self.ring = ring
        self.normalize = normalize
        self.direction = direction
        use_h, use_w = "h" in direction, "w" in direction
        self.k_h = self.n_taps if use_h else 1
        self.k_w = self.n_taps if use_w else 1
        self.up_h, self.down_h = (int(self.up[0]), int(self.down[0])) if use_h else (1, 1)
        self.up_w, self.down_w = (int(self.up[1]), int(self.down[1])) if use_w else (1, 1)

        kernel = torch.tensor(self.window, dtype=torch.float32)
        if normalize:
            kernel = kernel / kernel.sum()
        kernel = kernel * (self.up_h * self.up_w) ** 0.5
        self.register_buffer("kernel", kernel)

        def pads(k, u, d):
            if u > 1:
                return (k - u + 1) // 2 + u - 1, (k - u) // 2
            return (k - d + 1) // 2, (k - d) // 2

        self.ph0, self.ph1 = pads(self.k_h, self.up_h, self.down_h)
        self.pw0, self.pw1 = pads(self.k_w, self.up_w, self.down_w)
        self.margin = max(self.ph0, self.ph1, self.pw0, self.pw1)
        self._cfg = DF.FirCfg(self.k_h, self.k_w, up=(self.up_h, self.up_w),
                              down=(self.down_h, self.down_w),
                              pad=(self.ph0, self.ph1, self.pw0, self.pw1),
                              mode=(K.PAD_REPLICATE, K.PAD_CIRCULAR if ring else K.PAD_REPLICATE))
        self._taps2d = None
        # 4-tap ring "hw" blur / 2x upsample: specialised kernel (dusty_resample4)
        self._fast_up = None
        if (self.n_taps == 4 and ring and direction == "hw" and self.down_h == 1 and self.down_w == 1
                and self.up_h == self.up_w and self.up_h in (1, 2)):
            self._fast_up = self.up_h
        self._taps_host = None

    def _taps(self, device):
        t = self._taps2d
        if t is None or t.device != device:
            k = self.kernel.detach().float().to(device)
            one = torch.ones(1, device=device)
            t = torch.outer(k if self.k_h > 1 else one, k if self.k_w > 1 else one).contiguous()
            self._taps2d = t
        return t

    def _apply(self, fn, *a, **kw):          # buffers may move / change: drop the caches
        self._taps2d = None
        self._taps_host = None
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self._taps2d = None
        self._taps_host = None
        return super()._load_from_state_dict(*a, **kw)

    def forward(self, h):
        if self._fast_up is not None and DF.resample4_supported(h, self._fast_up):
            if self._taps_host is None:      # one-time host copy of the 4 taps
                self._taps_host = tuple(self.kernel.detach().float().cpu().tolist())
            return DF.resample4(h, self._taps_host, self._fast_up)
        return DF.fir2d(h, self._taps(h.device), self._cfg)

    def forward_with_sumsq(self, h):
        """(forward(h), sum(forward(h)^2) as a 1-element fp32 device buffer or None): the 2x
        upsampling kernel can emit the statistic the following ModConv2d's EMA needs."""
        if self._fast_up == 2 and DF.resample4_supported(h, 2):
            if self._taps_host is None:
                self._taps_host = tuple(self.kernel.detach().float().cpu().tolist())
            return DF.up2_with_sumsq(h, self._taps_host)
        return self.forward(h), None

    def extra_repr(self):
        return f'filter_type={self.window}, up={self.up}, down={self.down}, direction="{self.direction}"'


class BlurVH(nn.Module):
    """reference common.py:141-155."""

    def __init__(self, window=[1, 2, 1], ring=True):
        super().__init__()
        self.blur_v = Resample(window=window, ring=ring, direction="h")
        self.blur_h = Resample(window=window, ring=ring, direction="w")

    def forward(self, x):
        return torch.cat([self.blur_v(x), self.blur_h(x)], dim=1)


# ---- dense convolution with an explicit first / second order ------------------------------
# Every dense convolution of the path runs on this package's own kernels: bf16 NHWC shapes on the
# tcgen05 implicit-GEMM family (conv_tc.cu: fprop / dgrad / wgrad), everything else -- fp32 parity
# mode, NCHW, 1- / 2- / 513-channel layers, zero padding -- on the CUDA-core family
# (conv_simt.cu).  There is no library convolution anywhere.  The autograd wiring is ours too:
# conv is bilinear in (x, w), so every derivative of every order is again one of fprop / dgrad /
# wgrad (PyTorch's generic convolution double-backward falls onto slow grouped formulations,
# ~100 ms per R1 step at B=64).
def _tc_ok(x, w, stride, padding, op="fprop"):
    return tuple(padding) == (0, 0) and DF.conv_tc_supported(x, w, stride, op)


def _x3_ok(x, w, stride, padding):
    """fp32 mode: the same tcgen05 kernels on split-bf16 operands (DF.conv2d_*_x3)."""
    return (tuple(padding) == (0, 0) and DF.fp32_on_tensor_cores() and DF.conv_x3_supported(x, w, stride))


def _conv_fprop(x, w, stride, padding=(0, 0)):
    if _tc_ok(x, w, stride, padding):
        return DF.conv2d_fprop_tc(x, w, stride)
    if _x3_ok(x, w, stride, padding):
        return DF.conv2d_fprop_x3(x, w, stride)
    return DF.conv2d_fprop_simt(x, w, stride, padding).to(x.dtype)


def _conv_grads(gy, x, w, stride, need_x, need_w, w_tco=None, padding=(0, 0)):
    gx = gw = None
    same = gy.dtype == x.dtype
    x3 = same and _x3_ok(x, w, stride, padding)
    if need_x:
        if same and _tc_ok(x, w, stride, padding, "dgrad"):
            gx = DF.conv2d_dgrad_tc(gy, w, stride, x.shape[2:], w_tco)
        elif x3:
            gx = DF.conv2d_dgrad_x3(gy, w, stride, x.shape[2:])
        else:
            gx = DF.conv2d_dgrad_simt(gy, w, stride, padding, x.shape[2:], like=x).to(x.dtype)
    if need_w:
        if same and _tc_ok(x, w, stride, padding, "wgrad"):
            gw = DF.conv2d_wgrad_tc(gy, x, stride, w.shape, w.dtype)
        elif x3:
            gw = DF.conv2d_wgrad_x3(gy, x, stride, w.shape)
            if not (w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous()):
                gw = gw.contiguous()
        else:
            gw = DF.conv2d_wgrad_simt(gy, x, stride, padding, w.shape).to(w.dtype)
            if w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous():
                gw = gw.contiguous(memory_format=torch.channels_last)
    return gx, gw


class _Conv2dFn(torch.autograd.Function):
    """w_tco: optional [R*S][C][O] copy of w for the data-gradient kernels (a representation of
    the same values, not a separate autograd input)."""

    @staticmethod
    def forward(ctx, x, w, stride, w_tco=None, padding=(0, 0)):
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.padding = stride, tuple(padding)
        ctx.w_tco = w_tco
        return _conv_fprop(x, w, stride, padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx, gw = _Conv2dBwdFn.apply(gy, x, w, ctx.stride, ctx.needs_input_grad[0],
                                    ctx.needs_input_grad[1], ctx.w_tco, ctx.padding)
        return gx, gw, None, None, None


class _Conv2dBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, x, w, stride, need_x, need_w, w_tco=None, padding=(0, 0)):
        ctx.save_for_backward(gy, x, w)
        ctx.stride, ctx.need, ctx.padding = stride, (need_x, need_w), tuple(padding)
        if DF._is_cl(x):
            gy = gy.contiguous(memory_format=torch.channels_last)
        elif not (DF.conv_fans_out_to_nhwc(x, w) and DF._is_cl(gy)):   # that layer's gy stays NHWC
            gy = gy.contiguous()
        gx, gw = _conv_grads(gy, x, w, stride, need_x, need_w, w_tco, padding)
        if (gw is not None and not gw.is_contiguous()
                and not gw.is_contiguous(memory_format=torch.channels_last)):
            gw = gw.contiguous()           # odd strides -> dense (OHWI filter grads stay OHWI)
        return gx, gw

    @staticmethod
    def backward(ctx, ggx, ggw):
        gy, x, w = ctx.saved_tensors
        s, pad = ctx.stride, ctx.padding
        need_gy, need_x, need_w = ctx.needs_input_grad[:3]
        g_gy = g_x = g_w = None
        if ggx is not None:
            ggx = ggx.contiguous(memory_format=torch.channels_last) if DF._is_cl(x) else ggx.contiguous()
            if need_gy:
                g_gy = _Conv2dFn.apply(ggx, w, s, None, pad)                      # fprop
            if need_w:
                g_w = _conv_grads(gy, ggx, w, s, False, True, None, pad)[1]       # wgrad(ggx, gy)
        if ggw is not None:
            if need_gy:
                t = _Conv2dFn.apply(x, ggw, s, None, pad)
                g_gy = t if g_gy is None else g_gy + t
            if need_x:
                g_x = _conv_grads(gy, x, ggw.contiguous(), s, True, False, None, pad)[0]   # dgrad(gy, ggw)
        return g_gy, g_x, g_w, None, None, None, None, None


def _convT_tc_ok(x, w, stride):
    O, C, R, S = w.shape
    return (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and O % 8 == 0
            and C % 8 == 0 and R <= 4 and S <= 4 and stride[0] in (1, 2) and stride[1] in (1, 2))


def _pad_channels(t, dim, mult=8):
    """Zero-pad dimension `dim` to a multiple of `mult` (the tcgen05 kernels address channels in
    16-byte units); differentiable, the gradient of the padding is dropped by the slice."""
    r = (-t.shape[dim]) % mult
    if r == 0:
        return t
    return F.pad(t, [0, 0] * (t.dim() - dim - 1) + [0, r])


class _ConvTranspose2dFn(torch.autograd.Function):
    """y = conv_transpose2d(x, w, stride, padding): the data-gradient kernel of the convolution
    whose filter is w [in_ch, out_ch, R, S] (reference vanilla.py:18-27, the 4x4 stride-2
    up-convolutions of the vanilla / dusty_v1 generators); its own gradients are that
    convolution's forward and filter gradient.  bf16: the tcgen05 kernels (the un-cropped data
    gradient, then the `padding` crop as a view; the backward zero-pads the incoming gradient
    with the package's pad kernel); otherwise the CUDA-core family."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, out_hw):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, tuple(padding), tuple(out_hw))
        ph, pw = padding
        if _convT_tc_ok(x, w, stride):
            full = (out_hw[0] + 2 * ph, out_hw[1] + 2 * pw)
            y = DF.conv2d_dgrad_tc(x.contiguous(memory_format=torch.channels_last), w, stride, full)
            return y[:, :, ph:ph + out_hw[0], pw:pw + out_hw[1]]
        return DF.conv2d_dgrad_simt(x, w, stride, padding, out_hw, like=x).to(x.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, out_hw = ctx.cfg
        ph, pw = padding
        gx = gw = None
        if _convT_tc_ok(x, w, stride) and gy.dtype == x.dtype:
            gyp = gy.contiguous(memory_format=torch.channels_last)
            if ph or pw:
                cfg = DF.FirCfg(1, 1, pad=(ph, ph, pw, pw), mode=(K.PAD_ZERO, K.PAD_ZERO))
                gyp = DF.fir2d(gyp, DF.device_taps([[1.0]], gyp.device), cfg)
                gyp = gyp.contiguous(memory_format=torch.channels_last)
            xc = x.contiguous(memory_format=torch.channels_last)
            if ctx.needs_input_grad[0]:
                gx = DF.conv2d_fprop_tc(gyp, w, stride)
            if ctx.needs_input_grad[1]:
                gw = DF.conv2d_wgrad_tc(xc, gyp, stride, w.shape, w.dtype)
            return gx, gw, None, None, None
        gy = gy.contiguous(memory_format=torch.channels_last) if DF._is_cl(x) else gy.contiguous()
        if ctx.needs_input_grad[0]:
            gx = DF.conv2d_fprop_simt(gy, w, stride, padding).to(x.dtype)
        if ctx.needs_input_grad[1]:
            gw = DF.conv2d_wgrad_simt(x, gy, stride, padding, w.shape).to(w.dtype)
        return gx, gw, None, None, None


def conv_transpose2d(x, w, bias, stride, padding, output_padding=(0, 0)):
    stride, padding = tuple(stride), tuple(padding)
    R, S = w.shape[2:]
    out_hw = ((x.shape[2] - 1) * stride[0] - 2 * padding[0] + R + output_padding[0],
              (x.shape[3] - 1) * stride[1] - 2 * padding[1] + S + output_padding[1])
    y = _ConvTranspose2dFn.apply(x, w, stride, padding, out_hw)
    return y if bias is None else y + bias.to(y.dtype).view(1, -1, 1, 1)


class _ConvBiasActFn(torch.autograd.Function):
    """y = lrelu(conv2d_valid(x, w) + bias) * gain with the bias / activation in the epilogue of
    our tcgen05 convolution (one write of y instead of write + read + write).  First-order
    backward = the fused bias_act backward followed by dgrad / wgrad; under create_graph=True it
    is re-expressed through the differentiable single ops."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, alpha, gain, w_tco=None, blur_taps=None):
        """blur_taps: also apply the ring blur + Pad(1, ring) that follows in a ResidualBlock
        (DF.blur_pad_cl) and return ITS output; the backward then runs the blur's adjoint, the
        activation gate and the bias-gradient reduction as ONE kernel (dusty_blur4_cl_adj_act)
        instead of the blur adjoint followed by a bias_act backward pass over the same tensor."""
        bf = None if bias is None else bias.detach().float().contiguous()
        y = DF.conv2d_fprop_tc(x, w, stride, bf, 3, alpha, gain)
        ctx.save_for_backward(x, w, bias, y)
        ctx.cfg = (stride, alpha, gain, blur_taps)
        ctx.w_tco = w_tco
        return y if blur_taps is None else DF.blur_pad_cl(y, blur_taps)

    @staticmethod
    def backward(ctx, gy):
        x, w, bias, y = ctx.saved_tensors
        stride, alpha, gain, blur_taps = ctx.cfg
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        if torch.is_grad_enabled():
            with torch.enable_grad():
                yc = DF.bias_act(conv2d_valid(x, w, stride), bias, alpha, gain)
                if blur_taps is not None:
                    yc = DF.blur_pad_cl(yc, blur_taps)
                wanted = [t for t, n in ((x, need_x), (w, need_w), (bias, need_b)) if n and t is not None]
                grads = list(torch.autograd.grad(yc, wanted, gy, create_graph=True, allow_unused=True))
            out = [grads.pop(0) if (n and t is not None) else None
                   for t, n in ((x, need_x), (w, need_w), (bias, need_b))]
            return out[0], out[1], out[2], None, None, None, None, None
        if blur_taps is not None:
            gpre, db = DF.blur_pad_adj_act(gy, y, blur_taps, alpha, gain)
        else:
            gpre, db = DF._BiasActBackward.apply(gy, y, bias is not None, alpha, gain)
        gx, gw = _Conv2dBwdFn.apply(gpre, x, w, stride, need_x, need_w, ctx.w_tco)
        gb = db.to(bias.dtype) if (need_b and bias is not None) else None
        return gx, gw, gb, None, None, None, None, None


def conv_bias_act_supported(x, w, stride) -> bool:
    """The fused epilogue exists on the halo-resident kernel (unit-stride 3x3 of the thin layers)."""
    stride = tuple(stride) if isinstance(stride, (tuple, list)) else (stride, stride)
    return (x.is_cuda and stride == (1, 1) and DF._is_cl(x) and DF.conv_tc_supported(x, w, stride, "fprop")
            and DF.conv_halo_ok(w, "fprop"))


def conv_bias_act(x, w, bias, stride, negative_slope=0.2, gain=2 ** 0.5, w_tco=None, blur_taps=None):
    stride = tuple(stride) if isinstance(stride, (tuple, list)) else (stride, stride)
    taps = None if blur_taps is None else tuple(float(t) for t in blur_taps)
    return _ConvBiasActFn.apply(x, w, bias, stride, float(negative_slope), float(gain), w_tco, taps)


def conv2d_valid(x, w, stride, w_tco=None):
    """Un-padded, bias-free 2-D convolution with analytic higher-order gradients."""
    return conv2d(x, w, None, stride, (0, 0), w_tco)


def conv2d(x, w, bias, stride, padding=(0, 0), w_tco=None):
    """Dense convolution (zero padding, optional bias) on this package's kernels."""
    stride = tuple(stride) if isinstance(stride, (tuple, list)) else (stride, stride)
    padding = tuple(padding) if isinstance(padding, (tuple, list)) else (padding, padding)
    if not x.is_cuda:
        raise RuntimeError("dusty_gan_v2_b200 convolutions run on CUDA tensors only (no CPU fallback)")
    if w.shape[1] != x.shape[1]:
        raise RuntimeError(f"conv2d: {x.shape[1]} input channels, filter expects {w.shape[1]} (groups unsupported)")
    y = _Conv2dFn.apply(x, w, stride, w_tco, padding)
    return y if bias is None else y + bias.to(y.dtype).view(1, -1, 1, 1)


class EqualLR(nn.Module):
    """reference common.py:158-184.  y = module(x / sqrt(fan_in)) * gain * lr_mul, computed
    with the scale folded into the weight (a [O, fan_in] tensor) rather than applied to the
    activation tensor."""

    def __init__(self, module, gain: float = 1.0, lr_mul=1.0):
        super().__init__()
        self.module = module
        self.gain = gain
        self.lr_mul = lr_mul
        self.gain_ = gain * lr_mul
        self.scale = 1.0 / math.sqrt(self.module.weight[0].numel())
        nn.init.normal_(self.module.weight, 0.0, 1.0 / lr_mul)
        if getattr(self.module, "bias", None) is not None:
            nn.init.constant_(self.module.bias, 0.0)

    def forward(self, x):
        m = self.module
        if isinstance(m, nn.Linear) and x.dtype == m.weight.dtype and x.numel() <= m.weight.numel():
            # small activations (style / latent vectors): scale x, not the weight -- one tiny
            # launch, exactly the reference's order of operations (common.py:180-181)
            y = F.linear(x * self.scale, m.weight, m.bias)
            return y if self.gain_ == 1.0 else y * self.gain_
        if (isinstance(m, nn.Linear) and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16)
                and DF.act_dtype() == torch.bfloat16 and m.weight.numel() >= (1 << 22)):
            # the 65536 -> 512 linear of D's epilogue (33.5 M weights, 134 MB in fp32) in
            # low-precision mode: our tcgen05 kind::tf32 GEMM straight on the fp32 master weight
            # (linear_tc.cu), the EqualLR scale applied to the [B, 512] result.  No weight-sized
            # elementwise pass exists in either direction (scaling + casting the weight to bf16
            # cost four 134 MB passes per forward/backward pair); TF32 keeps 3 more mantissa bits
            # than bf16.  Data and weight gradient read the same two tensors in place.
            y = DF.linear_nt(x, m.weight) * (self.scale * self.gain_)
            return y if m.bias is None else y + m.bias * self.gain_
        if (isinstance(m, nn.Linear) and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
                and DF.act_dtype() == torch.bfloat16 and m.out_features <= 8):
            # the 512 -> 1 head: tiny GEMMs on the CUDA-core kernel
            y = DF.linear_nt(x, m.weight) * (self.scale * self.gain_)
            return y if m.bias is None else y + m.bias * self.gain_
        if (isinstance(m, nn.Conv2d) and x.is_cuda and x.dtype == torch.bfloat16 and DF._is_cl(x)
                and m.bias is None and m.padding == (0, 0) and m.dilation == (1, 1) and m.groups == 1):
            w, w_tco = self.prepared_weight(x.dtype, with_tco=True)
            return conv2d_valid(x, w, m.stride, w_tco)
        w = (m.weight * (self.scale * self.gain_)).to(x.dtype)
        b = None if m.bias is None else (m.bias * self.gain_).to(x.dtype)
        if isinstance(m, nn.Linear):
            return F.linear(x, w, b)
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.dilation != (1, 1) or m.groups != 1 or isinstance(m.padding, str):
                raise NotImplementedError("dilated / grouped convolutions are not part of the path")
            if x.is_cuda and DF.act_dtype() == torch.bfloat16:
                return self._conv_low_precision(x, w, b)
            if isinstance(m, nn.Conv2d):
                if DF._is_cl(x):        # NHWC activations: OHWI filter memory
                    w = w.contiguous(memory_format=torch.channels_last)
                return conv2d(x, w, b, m.stride, m.padding)
            return conv_transpose2d(x, w, b, m.stride, m.padding, m.output_padding)
        return m(x * self.scale) * self.gain_

    def _conv_low_precision(self, x, w, b):
        """The dense (transposed) convolutions of the vanilla / dusty_v1 baselines (reference
        vanilla.py:7-105) in bf16 mode: activations become bf16 NHWC and stay so, channel counts
        are zero-padded to multiples of 8 (1-channel heads, the 2-channel BlurVH pair) so that
        every layer runs on the tcgen05 kernels; a convolution whose kernel covers the whole map
        (the generator's projection from [B, C, 1, 1], the discriminator's logit) is a GEMM."""
        m = self.module
        bf = torch.bfloat16
        transposed = isinstance(m, nn.ConvTranspose2d)
        B = x.shape[0]
        if (transposed and tuple(x.shape[2:]) == (1, 1) and m.stride == (1, 1) and m.padding == (0, 0)):
            cin, cout, R, S = w.shape
            y = DF.linear_nt(x.reshape(B, cin).float(), w.float().reshape(cin, cout * R * S).t())
            y = y.reshape(B, cout, R, S).to(bf).contiguous(memory_format=torch.channels_last)
        elif (not transposed and tuple(x.shape[2:]) == tuple(m.kernel_size) and m.padding == (0, 0)):
            O = w.shape[0]
            y = DF.linear_nt(x.float().flatten(1), w.float().flatten(1)).reshape(B, O, 1, 1)
        else:
            xb = _pad_channels(x.to(bf), 1).contiguous(memory_format=torch.channels_last)
            wb = _pad_channels(_pad_channels(w.to(bf), 0), 1)
            if transposed:
                cout = w.shape[1]
                y = conv_transpose2d(xb, wb, None, m.stride, m.padding, m.output_padding)[:, :cout]
            else:
                O = w.shape[0]
                y = conv2d(xb, wb.contiguous(memory_format=torch.channels_last), None, m.stride, m.padding)[:, :O]
            if y.shape[1] % 8 == 0:
                y = y.contiguous(memory_format=torch.channels_last)
        return y if b is None else y + b.to(y.dtype).view(1, -1, 1, 1)

    def prepared_weight(self, dtype, with_tco=False):
        """Scaled conv filter in `dtype`, channels_last memory: one kernel (scale + cast + layout)
        forward, one backward (DF.prep_conv_weight) instead of three / four ATen kernels.
        with_tco=True: (w, w_tco) with the [R*S][C][O] form the data-gradient kernels read (made
        by the same launch when our tcgen05 convolutions are in use, else None)."""
        m = self.module
        ready = getattr(self, "_ready", None)
        if ready is not None and with_tco and ready[0].dtype == dtype:
            # prepared ahead on the discriminator's side stream (filter bank, dusty_v2.py): wait
            # for it, and keep its memory out of the side stream's pool while this stream reads it
            w, w_tco, event = ready
            self._ready = None
            main = torch.cuda.current_stream()
            main.wait_event(event)
            w.record_stream(main)
            if w_tco is not None:
                w_tco.record_stream(main)
            return w, w_tco
        if m.weight.is_cuda and m.weight.dim() == 4:
            # the [R*S][C][O] copy feeds the data-gradient kernels only: skip it when no
            # backward pass can follow
            want = with_tco and torch.is_grad_enabled()
            r = DF.prep_conv_weight(m.weight, self.scale * self.gain_, dtype, want)
            if with_tco:
                return r if want else (r, None)
            return r
        w = (m.weight * (self.scale * self.gain_)).to(dtype).contiguous(memory_format=torch.channels_last)
        return (w, None) if with_tco else w

    def extra_repr(self):
        return f"gain={self.gain}, lr_mul={self.lr_mul}"


class Conv2d(nn.Sequential):
    """reference common.py:187-210: custom padding + Conv2d + EqualLR."""

    def __init__(self, in_ch, out_ch, kernel_size, stride, padding, bias=True, ring=False,
                 equal_lr=False, gain=1.0, lr_mul=1.0):
        layers = []
        if padding != 0:
            layers.append(Pad(padding=padding, ring=ring))
        conv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, 0, bias=bias)
        layers.append(EqualLR(conv, gain, lr_mul) if equal_lr else conv)
        super().__init__(*layers)


class PixelNorm(nn.Module):
    """reference common.py:213-223 (on [B, 512] latents: negligible work, plain torch)."""

    def forward(self, x, alpha: float = 1e-8):
        return x / x.pow(2.0).mean(dim=1, keepdim=True).add(alpha).sqrt()


class MinibatchStdDev(nn.Module):
    """reference common.py:226-253; features == 1 only (all shipped configs)."""

    def __init__(self, group=4, features=1):
        super().__init__()
        if features != 1:
            raise NotImplementedError("MinibatchStdDev: only features == 1 is implemented")
        self.group = group
        self.features = features
        # number of independent mini-batches stacked along dim 0 (the trainer evaluates real and
        # fake images in one pass; statistics must not mix them)
        self.sub_batches = 1

    def forward(self, x, alpha: float = 1e-8):
        if self.sub_batches > 1:
            if x.shape[0] % self.sub_batches:
                raise RuntimeError("batch is not divisible by sub_batches")
            return torch.cat([DF.minibatch_stddev(c, self.group, alpha)
                              for c in x.chunk(self.sub_batches, dim=0)], dim=0)
        return DF.minibatch_stddev(x, self.group, alpha)

    def extra_repr(self):
        return f"group={self.group}, features={self.features}"
