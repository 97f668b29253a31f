"""Autograd-aware host wrappers around the C ABI (include/dusty_b200.h).

PyTorch supplies device memory, the current stream and the autograd tape; every
computation below is a kernel of libdusty_b200.so.  CPU tensors are rejected -- the CPU
restatement of the reference lives in oracle/ and is test infrastructure only.

Autograd conventions follow the reference: explicit Function pairs whose
backward is itself a Function, so second order (the R1 penalty) works:
  bias_act      <-> FusedLeakyReLUFunction / ...Backward   (fused_act.py:20-90)
  fir2d / adj   <-> UpFirDn2d / UpFirDn2dBackward          (upfirdn2d.py:20-145)
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _cabi as K

# --------------------------------------------------------------------------- precision
_PRECISION = {"act": torch.float32, "modconv_impl": 0, "fp32_tc": True}


def set_precision(name: str):
    """'fp32': fp32 storage + exact-FMA SIMT contractions (parity mode).
    'bf16': bf16 activations inside the blocks the reference runs under fp16 autocast
    (dusty_v2.yaml num_fp16_layers: -1), fp32 master weights and accumulation."""
    if name in ("fp32", "float32"):
        _PRECISION["act"] = torch.float32
    elif name in ("bf16", "bfloat16"):
        _PRECISION["act"] = torch.bfloat16
    else:
        raise ValueError(f"unknown precision {name!r}")


def set_fp32_tensor_cores(enabled: bool):
    """fp32 mode: run the dense contractions on tcgen05 with split-bf16 operands (default), or
    on the CUDA-core kernels (exact fp32 FMA order of the parity oracle)."""
    _PRECISION["fp32_tc"] = bool(enabled)


def fp32_on_tensor_cores() -> bool:
    return _PRECISION["fp32_tc"]


def act_dtype() -> torch.dtype:
    return _PRECISION["act"]


def set_modconv_impl(impl: int):
    """0 auto, 1 SIMT, 2 tcgen05, 3 tcgen05 without the batch-fused tiles (see dusty_modconv_fwd)."""
    _PRECISION["modconv_impl"] = int(impl)


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


_CL = torch.channels_last


def _is_cl(t: torch.Tensor) -> bool:
    """Stored NHWC (torch.channels_last) and not simultaneously NCHW-contiguous."""
    return t.ndim == 4 and (not t.is_contiguous()) and t.is_contiguous(memory_format=_CL)


def _cl_vec_ok(t: torch.Tensor, pow2: bool = False) -> bool:
    v = 16 // t.element_size()
    c = t.shape[1]
    if c % v:
        return False
    cv = c // v
    return (not pow2) or (cv <= 256 and 256 % cv == 0)


def _canon(t: torch.Tensor, need_pow2: bool = False) -> torch.Tensor:
    """Keep NCHW-contiguous and supported NHWC tensors as they are; copy anything else to NCHW."""
    if t.is_contiguous():
        return t
    if _is_cl(t) and t.dtype in (torch.float32, torch.bfloat16) and _cl_vec_ok(t, need_pow2):
        return t
    return t.contiguous()


def _like(t: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """t in the memory format of ref (NHWC or NCHW)."""
    if _is_cl(ref):
        return t if _is_cl(t) else t.contiguous(memory_format=_CL)
    return _contig(t)


# --------------------------------------------------------------------------- bias + act
def _bias_act_raw(x, bias, ref, act, grad, alpha, scale):
    K.require_cuda(x, bias, ref)
    x = _canon(x, need_pow2=True)
    y = torch.empty_like(x)
    C = x.shape[1] if x.ndim >= 2 else 1
    if _is_cl(x):
        if bias is not None and bias.numel() == 0:
            bias = None
        if bias is not None:
            bias = _contig(bias.to(x.dtype))
            if bias.numel() != C:
                raise RuntimeError(f"bias has {bias.numel()} elements, expected {C}")
        if ref is not None and ref.numel() == 0:
            ref = None
        if ref is not None:
            if ref.shape != x.shape or ref.dtype != x.dtype:
                raise RuntimeError("refer must match input shape and dtype")
            ref = _like(ref, x)
        K.call("dusty_bias_act_cl", K.ptr(x), K.ptr(bias), K.ptr(ref), K.ptr(y), x.numel(), C, act,
               grad, alpha, scale, K.dtype_code(x), K.stream_of(x))
        return y
    inner = 1
    for s in x.shape[2:]:
        inner *= s
    if bias is not None:
        if bias.numel() == 0:
            bias = None
        else:
            bias = _contig(bias.to(x.dtype))
            if bias.numel() != C:
                raise RuntimeError(f"bias has {bias.numel()} elements, expected {C}")
    if ref is not None and ref.numel() == 0:
        ref = None
    if ref is not None:
        ref = _contig(ref)
        if ref.shape != x.shape or ref.dtype != x.dtype:
            raise RuntimeError("refer must match input shape and dtype")
    K.call("dusty_bias_act", K.ptr(x), K.ptr(bias), K.ptr(ref), K.ptr(y), x.numel(), C, inner,
           act, grad, alpha, scale, K.dtype_code(x), K.stream_of(x))
    return y


def fused_bias_act(input, bias, refer, act: int, grad: int, alpha: float, scale: float):
    """Same contract as the reference's pybind `fused.fused_bias_act`
    (fused_bias_act.cpp:18-32): empty `bias` / `refer` tensors mean "absent"."""
    if not input.is_cuda:
        raise RuntimeError("input must be a CUDA tensor")
    if not input.is_contiguous():
        raise RuntimeError("input must be contiguous")
    return _bias_act_raw(input, bias, refer, act, grad, float(alpha), float(scale))


class _BiasActBackward(Function):
    @staticmethod
    def forward(ctx, gy, out, has_bias, alpha, scale):
        gy = _like(gy, out)
        gx = torch.empty_like(gy)
        C = gy.shape[1]
        inner = 1
        for s in gy.shape[2:]:
            inner *= s
        db = torch.zeros(C, device=gy.device, dtype=torch.float32) if has_bias else None
        if _is_cl(out):
            K.call("dusty_bias_act_bwd_cl", K.ptr(gy), K.ptr(out), K.ptr(gx), K.ptr(db),
                   gy.numel() // C, C, alpha, scale, K.dtype_code(gy), K.stream_of(gy))
        else:
            K.call("dusty_bias_act_bwd", K.ptr(gy), K.ptr(out), K.ptr(gx), K.ptr(db), gy.shape[0], C,
                   inner, alpha, scale, K.dtype_code(gy), K.stream_of(gy))
        ctx.save_for_backward(out)
        ctx.alpha, ctx.scale = alpha, scale
        return gx, db

    @staticmethod
    def backward(ctx, ggx, ggb):
        (out,) = ctx.saved_tensors
        b = None if ggb is None else ggb.to(ggx.dtype)
        gg_out = _bias_act_raw(ggx, b, out, 3, 1, ctx.alpha, ctx.scale)
        return gg_out, None, None, None, None


class _BiasAct(Function):
    @staticmethod
    def forward(ctx, x, bias, alpha, scale):
        out = _bias_act_raw(x, bias, None, 3, 0, alpha, scale)
        ctx.save_for_backward(out)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = bias.dtype if bias is not None else None
        ctx.alpha, ctx.scale = alpha, scale
        return out

    @staticmethod
    def backward(ctx, gy):
        (out,) = ctx.saved_tensors
        gx, db = _BiasActBackward.apply(gy, out, ctx.has_bias, ctx.alpha, ctx.scale)
        if ctx.has_bias:
            db = db.to(ctx.bias_dtype)
        else:
            db = None
        return gx, db, None, None


def bias_act(x, bias=None, negative_slope: float = 0.2, scale: float = 2 ** 0.5):
    K.require_cuda(x, bias)
    return _BiasAct.apply(_canon(x, need_pow2=True), bias, float(negative_slope), float(scale))


class _ResidualTail(Function):
    """y = (lrelu(x + bias) * gain + skip) * c, NHWC, one pass each way (cl_ops.cu).  Saves the
    pre-activation x (the activated tensor is never materialised).  Under create_graph=True
    the backward is re-expressed through the differentiable single ops."""

    @staticmethod
    def forward(ctx, x, bias, skip, alpha, gain, c):
        xc = _canon(x, need_pow2=True)
        sk = _like(skip, xc)
        C = xc.shape[1]
        b = None if bias is None else _contig(bias.detach().to(xc.dtype))
        y = torch.empty_like(xc)
        K.call("dusty_bias_act_add_cl", K.ptr(xc), K.ptr(b), K.ptr(sk), K.ptr(y), xc.numel(), C, alpha,
               gain, c, K.dtype_code(xc), K.stream_of(xc))
        ctx.save_for_backward(x, bias, skip)
        ctx.cfg = (alpha, gain, c)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, bias, skip = ctx.saved_tensors
        alpha, gain, c = ctx.cfg
        need_x, need_b, need_s = ctx.needs_input_grad[:3]
        if torch.is_grad_enabled():
            with torch.enable_grad():
                yc = (bias_act(x, bias, alpha, gain) + skip) * c
                wanted = [t for t, n in ((x, need_x), (bias, need_b), (skip, need_s)) if n and t is not None]
                grads = list(torch.autograd.grad(yc, wanted, gy, create_graph=True, allow_unused=True))
            out = [grads.pop(0) if (n and t is not None) else None
                   for t, n in ((x, need_x), (bias, need_b), (skip, need_s))]
            return out[0], out[1], out[2], None, None, None
        xc = _canon(x.detach(), need_pow2=True)
        g = _like(gy, xc)
        C = xc.shape[1]
        b = None if bias is None else _contig(bias.detach().to(xc.dtype))
        dx = torch.empty_like(xc)
        dskip = torch.empty_like(xc) if need_s else None
        db = torch.zeros(C, device=xc.device, dtype=torch.float32) if (need_b and bias is not None) else None
        K.call("dusty_bias_act_add_bwd_cl", K.ptr(g), K.ptr(xc), K.ptr(b), K.ptr(dx), K.ptr(dskip), K.ptr(db),
               xc.numel() // C, C, alpha, gain, c, K.dtype_code(xc), K.stream_of(xc))
        return (dx if need_x else None, None if db is None else db.to(bias.dtype), dskip, None, None, None)


def residual_tail_supported(x: torch.Tensor, skip: torch.Tensor) -> bool:
    return (x.is_cuda and x.dim() == 4 and _is_cl(x) and _cl_vec_ok(x, pow2=True) and skip.shape == x.shape
            and skip.dtype == x.dtype)


def residual_tail(x, bias, skip, negative_slope: float = 0.2, gain: float = 2 ** 0.5,
                  c: float = 2 ** -0.5):
    K.require_cuda(x, bias, skip)
    return _ResidualTail.apply(x, bias, skip, float(negative_slope), float(gain), float(c))


# --------------------------------------------------------------------------- FIR family
_TAPS_CACHE = {}


def device_taps(taps2d: Sequence[Sequence[float]], device) -> torch.Tensor:
    key = (tuple(tuple(float(v) for v in r) for r in taps2d), str(device))
    t = _TAPS_CACHE.get(key)
    if t is None:
        t = torch.tensor(key[0], dtype=torch.float32, device=device)
        _TAPS_CACHE[key] = t
    return t


class FirCfg:
    """Geometry of one polyphase FIR (see dusty_fir2d in the header)."""
    __slots__ = ("kh", "kw", "flip", "up_y", "up_x", "down_y", "down_x", "pad_y0", "pad_y1",
                 "pad_x0", "pad_x1", "mode_y", "mode_x")

    def __init__(self, kh, kw, flip=0, up=(1, 1), down=(1, 1), pad=(0, 0, 0, 0),
                 mode=(K.PAD_ZERO, K.PAD_ZERO)):
        self.kh, self.kw, self.flip = int(kh), int(kw), int(flip)
        self.up_y, self.up_x = int(up[0]), int(up[1])
        self.down_y, self.down_x = int(down[0]), int(down[1])
        self.pad_y0, self.pad_y1, self.pad_x0, self.pad_x1 = (int(p) for p in pad)
        self.mode_y, self.mode_x = int(mode[0]), int(mode[1])

    def out_hw(self, h, w):
        oh = (h * self.up_y + self.pad_y0 + self.pad_y1 - self.kh + self.down_y) // self.down_y
        ow = (w * self.up_x + self.pad_x0 + self.pad_x1 - self.kw + self.down_x) // self.down_x
        return oh, ow

    def args(self):
        return (self.up_y, self.up_x, self.down_y, self.down_x, self.pad_y0, self.pad_x0,
                self.mode_y, self.mode_x)


def _fir_raw(x, taps, cfg: FirCfg, adjoint: bool, in_hw=None):
    K.require_cuda(x, taps)
    x = _contig(x)
    lead = x.shape[:-2]
    n = 1
    for s in lead:
        n *= s
    if not adjoint:
        h, w = x.shape[-2:]
        oh, ow = cfg.out_hw(h, w)
        if oh < 0 or ow < 0:
            raise RuntimeError("fir2d: negative output size")
        y = torch.empty(*lead, oh, ow, device=x.device, dtype=x.dtype)
        K.call("dusty_fir2d", K.ptr(x), K.ptr(y), K.ptr(taps), cfg.kh, cfg.kw, cfg.flip, n, h, w,
               oh, ow, *cfg.args(), K.dtype_code(x), K.stream_of(x))
        return y
    h, w = in_hw
    oh, ow = x.shape[-2:]
    assert (oh, ow) == cfg.out_hw(h, w), "adjoint: gradient shape does not match geometry"
    y = torch.empty(*lead, h, w, device=x.device, dtype=x.dtype)
    K.call("dusty_fir2d_adj", K.ptr(x), K.ptr(y), K.ptr(taps), cfg.kh, cfg.kw, cfg.flip, n, h, w,
           oh, ow, *cfg.args(), K.dtype_code(x), K.stream_of(x))
    return y


class _Fir(Function):
    @staticmethod
    def forward(ctx, x, taps, cfg):
        ctx.cfg, ctx.in_hw = cfg, tuple(x.shape[-2:])
        ctx.save_for_backward(taps)
        return _fir_raw(x, taps, cfg, False)

    @staticmethod
    def backward(ctx, gy):
        (taps,) = ctx.saved_tensors
        return _FirAdj.apply(gy, taps, ctx.cfg, ctx.in_hw), None, None


class _FirAdj(Function):
    @staticmethod
    def forward(ctx, gy, taps, cfg, in_hw):
        ctx.cfg = cfg
        ctx.save_for_backward(taps)
        return _fir_raw(gy, taps, cfg, True, in_hw)

    @staticmethod
    def backward(ctx, ggx):
        (taps,) = ctx.saved_tensors
        return _Fir.apply(ggx, taps, ctx.cfg), None, None, None


def fir2d(x: torch.Tensor, taps: torch.Tensor, cfg: FirCfg) -> torch.Tensor:
    """x: [..., H, W]; taps: fp32 device tensor [kh, kw] (no gradient flows to the taps,
    as in the reference: upfirdn2d.py:101,145)."""
    if taps.dtype != torch.float32 or tuple(taps.shape) != (cfg.kh, cfg.kw):
        raise RuntimeError("taps must be an fp32 [kh, kw] tensor")
    return _Fir.apply(x, _contig(taps.detach()), cfg)


# 4-tap separable fast path (blur / 2x upsample, circular W, replicate H)
def _resample4_raw(x, taps4, up, adjoint):
    if up == 1 and _is_cl(x) and _cl_vec_ok(x):
        y = torch.empty_like(x)
        B, C, H, W = x.shape
        K.call("dusty_blur4_cl", K.ptr(x), K.ptr(y), taps4[0], taps4[1], taps4[2], taps4[3], B, H, W,
               C, 1 if adjoint else 0, 0, K.dtype_code(x), K.stream_of(x))
        return y
    x = _contig(x)
    lead = x.shape[:-2]
    n = 1
    for s_ in lead:
        n *= s_
    h, w = x.shape[-2:]
    if adjoint:
        H, W = h // up, w // up
        y = torch.empty(*lead, H, W, device=x.device, dtype=x.dtype)
    else:
        H, W = h, w
        y = torch.empty(*lead, H * up, W * up, device=x.device, dtype=x.dtype)
    K.call("dusty_resample4", K.ptr(x), K.ptr(y), taps4[0], taps4[1], taps4[2], taps4[3], n, H, W,
           up, 1 if adjoint else 0, K.dtype_code(x), K.stream_of(x))
    return y


class _Resample4(Function):
    @staticmethod
    def forward(ctx, x, taps4, up, adjoint):
        ctx.cfg = (taps4, up, adjoint)
        return _resample4_raw(x, taps4, up, adjoint)

    @staticmethod
    def backward(ctx, g):
        taps4, up, adjoint = ctx.cfg
        return _Resample4.apply(g, taps4, up, not adjoint), None, None, None


class _BlurPadCL(Function):
    """blur (4-tap, ring) followed by Pad(1, ring) on an NHWC tensor, one kernel each way."""

    @staticmethod
    def forward(ctx, x, taps4, adjoint):
        B, C = x.shape[:2]
        x = x if _is_cl(x) else x.contiguous(memory_format=_CL)     # any layout may arrive here
        if adjoint:
            H, W = x.shape[2] - 2, x.shape[3] - 2
            y = torch.empty((B, C, H, W), device=x.device, dtype=x.dtype, memory_format=_CL)
        else:
            H, W = x.shape[2:]
            y = torch.empty((B, C, H + 2, W + 2), device=x.device, dtype=x.dtype, memory_format=_CL)
        K.call("dusty_blur4_cl", K.ptr(x), K.ptr(y), taps4[0], taps4[1], taps4[2], taps4[3], B, H, W,
               C, 1 if adjoint else 0, 1, K.dtype_code(x), K.stream_of(x))
        ctx.cfg = (taps4, adjoint)
        return y

    @staticmethod
    def backward(ctx, g):
        taps4, adjoint = ctx.cfg
        return _BlurPadCL.apply(g, taps4, not adjoint), None, None


def blur_pad_cl_supported(x: torch.Tensor) -> bool:
    return (x.is_cuda and _is_cl(x) and _cl_vec_ok(x) and x.dtype in (torch.float32, torch.bfloat16)
            and x.shape[-2] >= 2 and x.shape[-1] >= 4)


def blur_pad_adj_act(g_pad: torch.Tensor, y_act: torch.Tensor, taps4, alpha: float, scale: float):
    """(d loss / d pre, d loss / d bias) for  pre -> y = lrelu(pre + bias) * scale -> blur -> ring pad,
    given the gradient of the PADDED blurred tensor: blur adjoint + ring fold + activation gate +
    bias-gradient reduction in one kernel (no autograd: the caller's backward owns the chain)."""
    K.require_cuda(g_pad, y_act)
    g_pad = g_pad if _is_cl(g_pad) else g_pad.contiguous(memory_format=_CL)
    if g_pad.dtype != y_act.dtype:
        g_pad = g_pad.to(y_act.dtype)
    B, C, H, W = y_act.shape
    if tuple(g_pad.shape) != (B, C, H + 2, W + 2) or not _is_cl(y_act):
        raise RuntimeError("blur_pad_adj_act: gradient must be [B, C, H+2, W+2], activation NHWC [B, C, H, W]")
    gpre = torch.empty((B, C, H, W), device=y_act.device, dtype=y_act.dtype, memory_format=_CL)
    db = torch.zeros(64, C, device=y_act.device, dtype=torch.float32)      # 64 replicas, see the header
    K.call("dusty_blur4_cl_adj_act", K.ptr(g_pad), K.ptr(y_act), K.ptr(gpre), K.ptr(db), taps4[0], taps4[1],
           taps4[2], taps4[3], B, H, W, C, alpha, scale, K.dtype_code(y_act), K.stream_of(y_act))
    return gpre, db.sum(dim=0)


def blur_pad_cl(x: torch.Tensor, taps4) -> torch.Tensor:
    """Pad(1, ring)(Resample([k0..k3], ring)(x)) for an NHWC tensor."""
    return _BlurPadCL.apply(x, tuple(float(t) for t in taps4), False)


class _BlurDown2CL(Function):
    """Resample([k0..k3], ring)(x)[:, :, ::2, ::2] on an NHWC tensor without evaluating the
    discarded positions (skip branch of the residual blocks), and its adjoint."""

    @staticmethod
    def forward(ctx, x, taps4, adjoint):
        B, C = x.shape[:2]
        x = x if _is_cl(x) else x.contiguous(memory_format=_CL)
        if adjoint:
            H, W = x.shape[2] * 2, x.shape[3] * 2
            y = torch.empty((B, C, H, W), device=x.device, dtype=x.dtype, memory_format=_CL)
        else:
            H, W = x.shape[2:]
            y = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=x.dtype, memory_format=_CL)
        K.call("dusty_blur4_down2_cl", K.ptr(x), K.ptr(y), taps4[0], taps4[1], taps4[2], taps4[3],
               B, H, W, C, 1 if adjoint else 0, K.dtype_code(x), K.stream_of(x))
        ctx.cfg = (taps4, adjoint)
        return y

    @staticmethod
    def backward(ctx, g):
        taps4, adjoint = ctx.cfg
        return _BlurDown2CL.apply(g, taps4, not adjoint), None, None


def blur_down2_cl_supported(x: torch.Tensor) -> bool:
    return (blur_pad_cl_supported(x) and x.shape[-2] % 2 == 0 and x.shape[-1] % 2 == 0)


def blur_down2_cl(x: torch.Tensor, taps4) -> torch.Tensor:
    return _BlurDown2CL.apply(x, tuple(float(t) for t in taps4), False)


class _ResidualFork(Function):
    """ResidualBlock input fork on an NHWC tensor: x -> (Pad(1, ring)(x), blur_down2(x)).  Both
    consumers' gradients arrive in ONE backward, which evaluates pad_adjoint + blur adjoint as a
    single gather pass (dusty_residual_fork_bwd_cl) instead of two adjoint launches and the
    autograd engine's accumulation add.  Under create_graph (R1) the backward is re-expressed
    through the differentiable single ops."""

    @staticmethod
    def forward(ctx, x, taps4):
        x = x if _is_cl(x) else x.contiguous(memory_format=_CL)
        B, C, H, W = x.shape
        # one pass over x writes both consumers' inputs (the pad is a by-product of the blur's loads)
        xp = torch.empty((B, C, H + 2, W + 2), device=x.device, dtype=x.dtype, memory_format=_CL)
        xd = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=x.dtype, memory_format=_CL)
        K.call("dusty_residual_fork_fwd_cl", K.ptr(x), K.ptr(xp), K.ptr(xd), taps4[0], taps4[1], taps4[2],
               taps4[3], B, H, W, C, K.dtype_code(x), K.stream_of(x))
        ctx.cfg = (taps4, (H, W))
        ctx.set_materialize_grads(False)
        return xp, xd

    @staticmethod
    def backward(ctx, g_pad, g_down):
        taps4, (H, W) = ctx.cfg
        if g_pad is None and g_down is None:
            return None, None
        if g_pad is None:
            return _BlurDown2CL.apply(g_down, taps4, True), None
        if g_down is None:
            return _PadAdj.apply(g_pad, _FORK_PADS, _FORK_MODES, (H, W)), None
        if torch.is_grad_enabled():      # create_graph=True: stay differentiable
            return (_PadAdj.apply(g_pad, _FORK_PADS, _FORK_MODES, (H, W))
                    + _BlurDown2CL.apply(g_down, taps4, True)), None
        g_pad = g_pad if _is_cl(g_pad) else g_pad.contiguous(memory_format=_CL)
        g_down = g_down if _is_cl(g_down) else g_down.contiguous(memory_format=_CL)
        B, C = g_down.shape[:2]
        dx = torch.empty((B, C, H, W), device=g_pad.device, dtype=g_pad.dtype, memory_format=_CL)
        K.call("dusty_residual_fork_bwd_cl", K.ptr(g_pad), K.ptr(g_down), K.ptr(dx), taps4[0],
               taps4[1], taps4[2], taps4[3], B, H, W, C, K.dtype_code(dx), K.stream_of(dx))
        return dx, None


def residual_fork_supported(x: torch.Tensor) -> bool:
    return blur_down2_cl_supported(x) and pad2d_supported(x, _FORK_PADS, _FORK_MODES)


def residual_fork(x: torch.Tensor, taps4):
    """(Pad(1, ring)(x), Resample([k0..k3], ring)(x)[:, :, ::2, ::2]) for an NHWC tensor."""
    return _ResidualFork.apply(x, tuple(float(t) for t in taps4))


def resample4_supported(x: torch.Tensor, up: int) -> bool:
    if x.ndim < 3 or not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
        return False
    vec = 16 // x.element_size()
    if up == 1 and _is_cl(x) and _cl_vec_ok(x) and x.shape[-2] >= 2 and x.shape[-1] >= 4:
        return True
    return x.shape[-1] % vec == 0 and x.shape[-2] >= 2 and x.shape[-1] >= 4 and up in (1, 2)


def resample4(x: torch.Tensor, taps4, up: int) -> torch.Tensor:
    """Blur (up=1) or 2x upsample (up=2) with per-axis taps `taps4` (python floats)."""
    K.require_cuda(x)
    return _Resample4.apply(x, tuple(float(t) for t in taps4), int(up), False)


class _Up2SumSq(Function):
    """2x upsampling (dusty_resample4 up=2) whose forward also yields sum(y^2) as a 1-element
    fp32 buffer (no grad): the EMA statistic of the ModConv2d that consumes y."""

    @staticmethod
    def forward(ctx, x, taps4):
        x = _contig(x)
        lead = x.shape[:-2]
        n = 1
        for s_ in lead:
            n *= s_
        H, W = x.shape[-2:]
        y = torch.empty(*lead, 2 * H, 2 * W, device=x.device, dtype=x.dtype)
        ss = torch.zeros(1, device=x.device, dtype=torch.float32)
        K.call("dusty_up2_sumsq", K.ptr(x), K.ptr(y), K.ptr(ss), taps4[0], taps4[1], taps4[2],
               taps4[3], n, H, W, K.dtype_code(x), K.stream_of(x))
        ctx.cfg = taps4
        ctx.mark_non_differentiable(ss)
        return y, ss

    @staticmethod
    def backward(ctx, g, _g_ss):
        return _Resample4.apply(g, ctx.cfg, 2, True), None


def up2_with_sumsq(x: torch.Tensor, taps4):
    """(Resample(up=2)(x), sum of its squares) in one launch."""
    K.require_cuda(x)
    return _Up2SumSq.apply(x, tuple(float(t) for t in taps4))


# single-axis zero-padded FIR (ADA's SYM6 passes) and the fused affine warp
class _Fir1d(Function):
    @staticmethod
    def forward(ctx, x, taps, axis, up, down, pad0, pad1, flip):
        x = _contig(x.float())
        lead = x.shape[:-2]
        n = 1
        for s_ in lead:
            n *= s_
        h, w = x.shape[-2:]
        k = taps.numel()
        n_in = w if axis == 1 else h
        n_out = (n_in * up + pad0 + pad1 - k + down) // down
        if n_out < 1:
            raise RuntimeError("fir1d: empty output")
        y = torch.empty(*lead, *((h, n_out) if axis == 1 else (n_out, w)), device=x.device,
                        dtype=torch.float32)
        K.call("dusty_fir1d", K.ptr(x), K.ptr(y), K.ptr(taps), k, flip, n, h, w, axis, up, down,
               pad0, pad1, K.stream_of(x))
        ctx.save_for_backward(taps)
        ctx.cfg = (axis, up, down, pad0, pad1, flip, n_in, n_out, k)
        return y

    @staticmethod
    def backward(ctx, g):
        (taps,) = ctx.saved_tensors
        axis, up, down, pad0, pad1, flip, n_in, n_out, k = ctx.cfg
        gp0 = k - pad0 - 1
        gp1 = n_in * up - n_out * down + pad0 - up + 1
        return (_Fir1d.apply(g, taps, axis, down, up, gp0, gp1, 1 - flip),) + (None,) * 7


def fir1d_supported(x, kernel, up, down, pad) -> bool:
    """upfirdn2d call that is a single-axis FIR with up/down in {1,2} on an fp32 CUDA tensor."""
    if not (x.is_cuda and x.dtype == torch.float32 and kernel.ndim == 2):
        return False
    kh, kw = kernel.shape
    (ux, uy), (dx, dy), (px0, px1, py0, py1) = up, down, pad
    if kh == 1 and kw <= 64:
        return uy == 1 and dy == 1 and py0 == 0 and py1 == 0 and ux in (1, 2) and dx in (1, 2)
    if kw == 1 and kh <= 64:
        return ux == 1 and dx == 1 and px0 == 0 and px1 == 0 and uy in (1, 2) and dy in (1, 2)
    return False


def fir1d(x, kernel, up, down, pad):
    """upfirdn2d(x, kernel[1,k] or [k,1], up=(ux,uy), down=(dx,dy), pad=(px0,px1,py0,py1))."""
    kh, kw = kernel.shape
    taps = _contig(kernel.detach().float().reshape(-1))
    if kh == 1:
        return _Fir1d.apply(x, taps, 1, int(up[0]), int(down[0]), int(pad[0]), int(pad[1]), 1)
    return _Fir1d.apply(x, taps, 0, int(up[1]), int(down[1]), int(pad[2]), int(pad[3]), 1)


# ---- f1: AdaptiveAugment as one device-side op (csrc/ada_fused.cu) ------------------------------
class _AdaApply(Function):
    """out = gain * A(img) + offset for per-sample axis-aligned transforms (params [B, 8]); linear
    in img: the backward is the adjoint kernel, the backward of that the forward without offset."""

    @staticmethod
    def forward(ctx, img, params, mode):
        img = _contig(img.float())
        B, ch, H, W = img.shape
        out = torch.empty_like(img)
        K.call("dusty_ada_apply", K.ptr(img), K.ptr(out), K.ptr(params), B * ch, H, W, mode, K.stream_of(img))
        ctx.save_for_backward(params)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, g):
        (params,) = ctx.saved_tensors
        return _AdaApply.apply(g, params, 2 if ctx.mode == 1 else 1), None, None


def ada_fused_supported(img: torch.Tensor) -> bool:
    if not (img.is_cuda and img.dim() == 4 and img.shape[1] == 1):
        return False
    return int(K.load().dusty_ada_apply_smem(int(img.shape[2]), int(img.shape[3]))) <= 227 * 1024


def ada_apply(img: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
    """params: fp32 CUDA [B, 8] = ax, tx, dy, ty of the inverse transform, colour gain, offset."""
    K.require_cuda(img, params)
    if params.dtype != torch.float32 or tuple(params.shape) != (img.shape[0], 8):
        raise RuntimeError("ada_apply: params must be fp32 [B, 8]")
    return _AdaApply.apply(img, _contig(params), 0)


def ada_sample(params: torch.Tensor, p: torch.Tensor, seed: int, counter: torch.Tensor, H: int, W: int,
               policy):
    """Draw B transforms on the device into params [B, 8] (see dusty_ada_sample)."""
    K.require_cuda(params, p, counter)
    if counter.dtype != torch.int64 or p.dtype != torch.float32:
        raise RuntimeError("ada_sample: int64 counter, fp32 p")
    pol = (K.C.c_float * 11)(*[float(v) for v in policy])
    K.call("dusty_ada_sample", K.ptr(params), K.ptr(p), int(seed) & 0xFFFFFFFFFFFFFFFF, K.ptr(counter),
           params.shape[0], H, W, pol, K.stream_of(params))
    return params


class _AffineWarp(Function):
    @staticmethod
    def forward(ctx, img, theta, out_hw, in_hw, adjoint):
        img = _contig(img.float())
        theta = _contig(theta.detach().float())
        N, C = img.shape[:2]
        (Ho, Wo), (Hi, Wi) = out_hw, in_hw
        out = torch.empty((N, C, Hi, Wi) if adjoint else (N, C, Ho, Wo), device=img.device,
                          dtype=torch.float32)
        K.call("dusty_affine_warp", K.ptr(img), K.ptr(out), K.ptr(theta), N, C, Hi, Wi, Ho, Wo,
               1 if adjoint else 0, K.stream_of(img))
        ctx.save_for_backward(theta)
        ctx.cfg = (out_hw, in_hw, adjoint)
        return out

    @staticmethod
    def backward(ctx, g):
        (theta,) = ctx.saved_tensors
        out_hw, in_hw, adjoint = ctx.cfg
        return _AffineWarp.apply(g, theta, out_hw, in_hw, not adjoint), None, None, None, None


def affine_warp(img, theta, out_hw):
    """grid_sample(img, affine_grid(theta, [N,C,*out_hw])), bilinear / zeros /
    align_corners=False, linear in img (first and higher orders through the adjoint)."""
    K.require_cuda(img, theta)
    return _AffineWarp.apply(img, theta, tuple(out_hw), tuple(img.shape[-2:]), False)


# small-halo padding fast path
_FORK_PADS = (1, 1, 1, 1)                       # (top, bottom, left, right)
_FORK_MODES = (K.PAD_REPLICATE, K.PAD_CIRCULAR)   # (mode_y, mode_x): Pad(1, ring=True)
def _pad_raw(x, pads, modes, adjoint, in_hw=None):
    if _is_cl(x) and _cl_vec_ok(x):
        B, C = x.shape[:2]
        pt, pb, pl, pr = pads
        H, W = in_hw if adjoint else x.shape[-2:]
        oshape = (B, C, H, W) if adjoint else (B, C, H + pt + pb, W + pl + pr)
        y = torch.empty(oshape, device=x.device, dtype=x.dtype, memory_format=_CL)
        K.call("dusty_pad2d_cl", K.ptr(x), K.ptr(y), B, H, W, C, pt, pb, pl, pr, modes[0], modes[1],
               1 if adjoint else 0, K.dtype_code(x), K.stream_of(x))
        return y
    x = _contig(x)
    lead = x.shape[:-2]
    n = 1
    for s_ in lead:
        n *= s_
    pt, pb, pl, pr = pads
    if adjoint:
        H, W = in_hw
        y = torch.empty(*lead, H, W, device=x.device, dtype=x.dtype)
    else:
        H, W = x.shape[-2:]
        y = torch.empty(*lead, H + pt + pb, W + pl + pr, device=x.device, dtype=x.dtype)
    K.call("dusty_pad2d", K.ptr(x), K.ptr(y), n, H, W, pt, pb, pl, pr, modes[0], modes[1],
           1 if adjoint else 0, K.dtype_code(x), K.stream_of(x))
    return y


class _Pad(Function):
    @staticmethod
    def forward(ctx, x, pads, modes):
        ctx.cfg = (pads, modes, tuple(x.shape[-2:]))
        return _pad_raw(x, pads, modes, False)

    @staticmethod
    def backward(ctx, g):
        pads, modes, in_hw = ctx.cfg
        return _PadAdj.apply(g, pads, modes, in_hw), None, None


class _PadAdj(Function):
    @staticmethod
    def forward(ctx, g, pads, modes, in_hw):
        ctx.cfg = (pads, modes)
        return _pad_raw(g, pads, modes, True, in_hw)

    @staticmethod
    def backward(ctx, gg):
        pads, modes = ctx.cfg
        return _Pad.apply(gg, pads, modes), None, None, None


def pad2d_supported(x, pads, modes) -> bool:
    if not x.is_cuda or x.ndim < 3 or x.dtype not in (torch.float32, torch.bfloat16):
        return False
    H, W = x.shape[-2:]
    pt, pb, pl, pr = pads
    return (all(0 <= p <= 4 for p in pads) and max(pt, pb) < H and max(pl, pr) < W
            and modes[0] in (K.PAD_REPLICATE, K.PAD_REFLECT)
            and modes[1] in (K.PAD_CIRCULAR, K.PAD_REPLICATE, K.PAD_REFLECT))


def pad2d(x, pads, modes):
    """pads = (top, bottom, left, right); modes = (mode_y, mode_x)."""
    K.require_cuda(x)
    return _Pad.apply(x, tuple(int(p) for p in pads), tuple(int(m) for m in modes))


# --------------------------------------------------------------------------- Fourier features
def fourier_features(angle: torch.Tensor, freqs: torch.Tensor, phase: torch.Tensor,
                     out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """angle [Ba,2,H,W] fp32 -> [Ba, 2F, H, W]; no gradient (angles are data)."""
    K.require_cuda(angle, freqs, phase)
    angle = _contig(angle.detach().float())
    f = _contig(freqs.detach().reshape(-1, 2).float())
    ph = _contig(phase.detach().float())
    ba, _, h, w = angle.shape
    nf = f.shape[0]
    out_dtype = out_dtype or torch.float32
    out = torch.empty(ba, 2 * nf, h, w, device=angle.device, dtype=out_dtype)
    K.call("dusty_fourier", K.ptr(angle), K.ptr(f), K.ptr(ph), K.ptr(out), ba, nf, h * w,
           K.dtype_code(out), K.stream_of(angle))
    return out


def angle_down2(angle: torch.Tensor) -> torch.Tensor:
    K.require_cuda(angle)
    angle = _contig(angle.detach().float())
    ba, c, h, w = angle.shape
    assert c == 2
    out = torch.empty(ba, 2, h // 2, w // 2, device=angle.device, dtype=torch.float32)
    K.call("dusty_angle_down2", K.ptr(angle), K.ptr(out), ba, h, w, K.stream_of(angle))
    return out


# --------------------------------------------------------------------------- modulated 1x1 conv
def _modconv_x3_ok(wb, src, O, c1, c2, P, op) -> bool:
    """fp32 mode (g1): does the split-bf16 form of this contraction lie in the tcgen05 kernels'
    domain (modconv_tc.cu: modconv_{fwd,dx,dw}_tc_supported on the tripled axis)?"""
    if not (_PRECISION["fp32_tc"] and _PRECISION["modconv_impl"] in (0, 2, 3)):
        return False
    if src.dtype != torch.float32 or wb.dtype != torch.float32 or not src.is_cuda:
        return False
    if c1 % 8 or c2 % 8 or P % 128:
        return False
    if op == "dx":
        return O % 8 == 0 and c1 >= 32
    if O < 32 or O % 16:
        return False
    if op == "fwd":
        return c2 == 0 or (3 * c1) % 64 == 0
    return c2 == 0 or c1 % 64 == 0           # dw: 64-channel boxes must not straddle the sources


def _dw_fused_ok(gpre, x1, x2, O, c1, c2, b2, P) -> bool:
    """dW of a contraction whose Fourier block is batch-shared: dense-GEMM formulation (bf16
    tcgen05 kernels; shapes inside both kernels' domains)."""
    if not (_PRECISION.get("dw_fused", True) and _PRECISION["modconv_impl"] in (0, 2)):
        return False
    if x2 is None or b2 != 1 or c2 < 64 or gpre.dtype != torch.bfloat16 or x2.dtype != torch.bfloat16:
        return False
    if P % 128 or c2 % 8 or c1 % 4 or (c1 + c2) % 4 or O < 32 or O % 16:
        return False
    return c1 == 0 or (x1 is not None and x1.dtype == torch.bfloat16 and c1 % 8 == 0)


def set_dw_fused(enabled: bool):
    _PRECISION["dw_fused"] = bool(enabled)


def _split_planes(x: torch.Tensor, pattern: int) -> torch.Tensor:
    """fp32 [B, C, H, W] (contiguous) -> bf16 [B, 3C, H, W]: parts stacked on the channel axis."""
    x = _contig(x)
    B, C = x.shape[:2]
    P = x[0, 0].numel()
    out = torch.empty((B, 3 * C) + tuple(x.shape[2:]), dtype=torch.bfloat16, device=x.device)
    return split_bf16x3(x, out, B, C, P, (C * P, P, 1), (3 * C * P, P, 1), pattern)


def _split_pixels(x: torch.Tensor, pattern: int) -> torch.Tensor:
    """fp32 [B, C, P...] -> bf16 [B, C, 3P]: parts side by side on the pixel axis."""
    x = _contig(x)
    B, C = x.shape[:2]
    P = x[0, 0].numel()
    out = torch.empty((B, C, 3 * P), dtype=torch.bfloat16, device=x.device)
    return split_bf16x3(x, out, B * C, 1, P, (P, 0, 1), (3 * P, P, 1), pattern)


def _split_wb_k(wb: torch.Tensor, c1: int, c2: int) -> torch.Tensor:
    """fp32 wb[B, O, c1 + c2] -> bf16 [B, O, 3*c1 + 3*c2], each source's columns tripled in place
    ([hi|lo|hi] of the feature block, then of the Fourier block)."""
    B, O, Kt = wb.shape
    out = torch.empty(B, O, 3 * Kt, dtype=torch.bfloat16, device=wb.device)
    if c1:
        split_bf16x3(wb, out, B * O, c1, 1, (Kt, 1, 0), (3 * Kt, 1, 0), 1)
    if c2:
        split_bf16x3(wb[:, :, c1:], out[:, :, 3 * c1:], B * O, c2, 1, (Kt, 1, 0), (3 * Kt, 1, 0), 1)
    return out


class _ModConvBmm(Function):
    """y[b] = act(wb[b] @ cat(x1[b], x2[b or 0]) + bias) ; x2 (Fourier features) has no grad."""

    @staticmethod
    def forward(ctx, wb, x1, x2, bias, act, alpha, scale, ema_var=None, ema_rows=None, parts=None,
                *handles):
        B, O, Kt = wb.shape
        ctx.parts = parts            # [(slot, row0, row1)] matching `handles` (see _ModPrep)
        want_ss, ctx_ss = _SUMSQ_REQUEST.pop("want", False), None
        c1 = 0 if x1 is None else x1.shape[1]
        c2 = 0 if x2 is None else x2.shape[1]
        assert c1 + c2 == Kt, (c1, c2, Kt)
        src = x1 if x1 is not None else x2
        H, W = src.shape[-2:]
        P = H * W
        b2 = 1 if x2 is None else x2.shape[0]
        y = torch.empty(B, O, H, W, device=src.device, dtype=src.dtype)
        biasf = None if bias is None else _contig(bias.detach().float().reshape(-1))
        ev = None if ema_var is None else ema_var.detach().float().reshape(1)
        ctx.ema = ev
        if ev is not None and not modconv_tc_domain(wb, src, O, c1, c2, P):
            raise RuntimeError("modconv_bmm: ema_var is applied by the tcgen05 kernels only")
        # heads: one device scalar per output row (kept alive by ctx; read again by dX)
        rows = None
        if ema_rows is not None:
            if len(ema_rows) != O or O > 4 or ev is not None:
                raise RuntimeError("modconv_bmm: ema_rows is one scalar per output row, O <= 4")
            rows = [r.detach().float().reshape(1) for r in ema_rows]
        ctx.ema_rows = rows
        rows_p = None if rows is None else (K.C.c_void_p * O)(*[r.data_ptr() for r in rows])
        if ev is None and rows is None and _modconv_x3_ok(wb, src, O, c1, c2, P, "fwd"):
            # fp32 mode on tcgen05: K axis tripled, [x_hi|x_hi|x_lo] . [w_hi|w_lo|w_hi]
            x1s = None if x1 is None else _split_planes(x1, 0)
            x2s = None if x2 is None else _split_planes(x2, 0)
            wbs = _split_wb_k(wb, c1, c2)
            K.call("dusty_modconv_fwd", K.ptr(wbs), K.ptr(x1s), K.ptr(x2s), K.ptr(biasf), K.ptr(y), B, O,
                   3 * c1, 3 * c2, b2, P, act, alpha, scale, K.BF16, K.BF16, 4, None, None, None,
                   K.stream_of(src))
        else:
            if want_ss and rows is None and modconv_tc_domain(wb, src, O, c1, c2, P):
                ctx_ss = torch.zeros(1, device=src.device, dtype=torch.float32)
            K.call("dusty_modconv_fwd", K.ptr(wb), K.ptr(x1), K.ptr(x2), K.ptr(biasf), K.ptr(y), B, O,
                   c1, c2, b2, P, act, alpha, scale, K.dtype_code(src), K.dtype_code(wb),
                   _PRECISION["modconv_impl"], K.ptr(ev), rows_p, K.ptr(ctx_ss), K.stream_of(src))
        _SUMSQ_REQUEST["got"] = ctx_ss
        ctx.save_for_backward(wb, x1, x2, y if act == 3 else None)
        ctx.cfg = (act, alpha, scale, bias is not None, None if bias is None else bias.shape,
                   None if bias is None else bias.dtype)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        wb, x1, x2, y = ctx.saved_tensors
        act, alpha, scale, has_bias, bshape, bdtype = ctx.cfg
        B, O, Kt = wb.shape
        gy = _contig(gy)
        H, W = gy.shape[-2:]
        P = H * W
        st = K.stream_of(gy)
        dt = K.dtype_code(gy)
        db = None
        if act == 3:
            gpre = torch.empty_like(gy)
            dbf = torch.zeros(O, device=gy.device, dtype=torch.float32) if has_bias else None
            K.call("dusty_bias_act_bwd", K.ptr(gy), K.ptr(y), K.ptr(gpre), K.ptr(dbf), B, O, P,
                   alpha, scale, dt, st)
            db = dbf
        else:
            gpre = gy if scale == 1.0 else gy * scale
            if has_bias:
                db = gpre.float().sum(dim=(0, 2, 3))
        gx1 = None
        c1 = 0 if x1 is None else x1.shape[1]
        c2 = 0 if x2 is None else x2.shape[1]
        b2 = 1 if x2 is None else x2.shape[0]
        if x1 is not None and ctx.needs_input_grad[1]:
            gx1 = torch.empty_like(x1)
            if ctx.ema is None and ctx.ema_rows is None and _modconv_x3_ok(wb, gpre, O, c1, c2, P, "dx"):
                gs = _split_planes(gpre, 0)                       # [B, 3O, P]
                wbo = torch.empty(B, 3 * O, Kt, device=wb.device, dtype=torch.bfloat16)
                split_bf16x3(wb, wbo, B, O, Kt, (O * Kt, Kt, 1), (3 * O * Kt, Kt, 1), 1)
                K.call("dusty_modconv_bwd_dx", K.ptr(wbo), K.ptr(gs), K.ptr(gx1), B, 3 * O, c1, Kt, P,
                       K.BF16, K.BF16, 4, None, None, st)
            else:
                rows = ctx.ema_rows
                rows_p = None if rows is None else (K.C.c_void_p * O)(*[r.data_ptr() for r in rows])
                K.call("dusty_modconv_bwd_dx", K.ptr(wb), K.ptr(gpre), K.ptr(gx1), B, O, c1, Kt, P, dt,
                       K.dtype_code(wb), _PRECISION["modconv_impl"], K.ptr(ctx.ema), rows_p, st)
        gwb = None
        n_fixed = 10
        want_h = [bool(f) for f in ctx.needs_input_grad[n_fixed:]]
        if ctx.needs_input_grad[0] or any(want_h):
            gw32 = torch.empty(B, O, Kt, device=gy.device, dtype=torch.float32)
            if _modconv_x3_ok(wb, gpre, O, c1, c2, P, "dw"):
                # contraction over pixels: the three terms side by side on the pixel axis
                gp = _split_pixels(gpre, 0)
                x1p = None if x1 is None else _split_pixels(x1, 1)
                x2p = None if x2 is None else _split_pixels(x2, 1)
                K.call("dusty_modconv_bwd_dw", K.ptr(gp), K.ptr(x1p), K.ptr(x2p), K.ptr(gw32), B, O, c1,
                       c2, b2, 3 * P, K.BF16, 2, 0, st)
            elif _dw_fused_ok(gpre, x1, x2, O, c1, c2, b2, P):
                # batch-shared Fourier block: its columns of dW for ALL samples are one dense GEMM
                # [(B*O), P] x [P, C2] (the block crosses L2 -> SM once instead of once per sample);
                # the feature columns stay per sample, written at the pitch of the full tensor
                K.call("dusty_gemm_bf16", K.ptr(gpre), K.ptr(x2), gw32.data_ptr() + 4 * c1, B * O, c2, P, 0, 0,
                       P, P, Kt, 1.0, 0, st)
                if c1:
                    K.call("dusty_modconv_bwd_dw", K.ptr(gpre), K.ptr(x1), None, K.ptr(gw32), B, O, c1, 0, 1, P,
                           dt, 2, Kt, st)
            else:
                K.call("dusty_modconv_bwd_dw", K.ptr(gpre), K.ptr(x1), K.ptr(x2), K.ptr(gw32), B, O, c1,
                       c2, b2, P, dt, _PRECISION["modconv_impl"], 0, st)
            if ctx.needs_input_grad[0]:
                gwb = gw32.to(wb.dtype)
        gh = []
        if ctx.parts:
            for (slot, r0, r1), want in zip(ctx.parts, want_h):
                if want:
                    slot["gw32"] = gw32 if (r0, r1) == (0, O) else gw32[:, r0:r1]
                    gh.append(gw32.reshape(-1)[:1])        # the edge's token; its value is unused
                else:
                    gh.append(None)
        if has_bias and ctx.needs_input_grad[3]:
            db = db.reshape(bshape).to(bdtype)
        else:
            db = None
        return (gwb, gx1, None, db, None, None, None, None, None, None) + tuple(gh)


def set_late_ema(enabled: bool):
    """ModConv2d: apply the EMA normaliser in the contraction's epilogue (default) or fold it into
    the per-sample weights as the reference does (the two differ by bf16 rounding of wb only)."""
    _PRECISION["late_ema"] = bool(enabled)


def late_ema_enabled() -> bool:
    return _PRECISION.get("late_ema", True)


def modconv_tc_domain(wb, src, O, c1, c2, P) -> bool:
    return src.is_cuda and wb.dtype == src.dtype and modconv_tc_domain_of(src.dtype, O, c1, c2, P)


def modconv_tc_domain_of(dtype, O, c1, c2, P) -> bool:
    """Will dusty_modconv_fwd (and, with x1, dusty_modconv_bwd_dx) take the tcgen05 kernels for
    this contraction?  Mirrors modconv_{fwd,dx}_tc_supported (modconv_tc.cu) and the dispatch of
    modconv.cu; needed where a caller relies on an epilogue only those kernels have (ema_var)."""
    if _PRECISION["modconv_impl"] == 1:
        return False
    if dtype != torch.bfloat16:
        return False
    if O < 32 or O % 16 or O > 768 or P % 128 or c1 % 8 or c2 % 8:
        return False
    if c1 % 64 and c2:
        return False
    return c1 == 0 or c1 >= 32


_SUMSQ_REQUEST = {}      # side channel of modconv_bmm(want_sumsq=True): not an autograd output


def modconv_bmm(wb, x1, x2=None, bias=None, act: int = 1, alpha: float = 0.2, scale: float = 1.0,
                ema_var=None, ema_rows=None, want_sumsq=False):
    """want_sumsq: also accumulate sum(y^2) in the epilogue (tcgen05 path; y._dusty_sumsq).
    ema_var: apply the layer's EMA normaliser 1 / (sqrt(ema_var) + 1e-8) to the product (the
    weights wb then carry none: modprep(..., ema_var=None, ema_late=ema_var)).  ema_rows: the
    same per output row (heads: O <= 4 rows, each its own ModConv2d), a list of O buffers."""
    K.require_cuda(wb, x1, x2, bias)
    parts = getattr(wb, "_dusty_parts", None)
    wb = _contig(wb)
    x1 = None if x1 is None else _contig(x1)
    x2 = None if x2 is None else _contig(x2.detach())
    handles = () if not parts else tuple(h for _, h, _, _ in parts)
    _SUMSQ_REQUEST["want"] = bool(want_sumsq)
    _SUMSQ_REQUEST["got"] = None
    y = _ModConvBmm.apply(wb, x1, x2, bias, int(act), float(alpha), float(scale), ema_var,
                          None if ema_rows is None else list(ema_rows),
                          None if not parts else [(s_, a, b) for s_, _, a, b in parts], *handles)
    ss = _SUMSQ_REQUEST.pop("got", None)
    if ss is not None:
        # sum(y^2) accumulated by the contraction's epilogue: the next ModConv2d's EMA statistic
        # (a plain attribute: it rides along with this tensor object only)
        y._dusty_sumsq = ss
    return y


class _ModPrep(Function):
    @staticmethod
    def forward(ctx, slin, weight, ema_var, scale, demod, out_dtype, rot, c1, ema_late=None, slot=None):
        slin = _contig(slin.float())
        w2 = _contig(weight.float().reshape(weight.shape[-4], weight.shape[-3]) if weight.ndim == 5
                     else weight.float())
        B, I = slin.shape
        O = w2.shape[0]
        wb = torch.empty(B, O, I, device=slin.device, dtype=out_dtype)
        stats = torch.empty(B + 2 + B * O, device=slin.device, dtype=torch.float32)
        ev = None if ema_var is None else ema_var.detach().float().reshape(1)
        nf = 0
        if rot is not None:
            rot = _contig(rot.detach().float())
            nf = rot.shape[1] // 2
            if rot.shape[0] != B or c1 + 2 * nf != I:
                raise RuntimeError("rotation table does not match the layer")
        K.call("dusty_modprep_fwd", K.ptr(slin), K.ptr(w2), K.ptr(ev), K.ptr(wb), K.ptr(stats), B, O,
               I, scale, 1 if demod else 0, K.dtype_code(wb), K.ptr(rot), c1, nf, K.stream_of(slin))
        ctx.save_for_backward(slin, w2, stats, rot)
        ctx.cfg = (scale, demod, tuple(weight.shape), weight.dtype, c1, nf)
        # the buffer itself, read at backward time (no forward runs between a layer's forward
        # and its backward, so the value is the one the contraction used)
        ctx.ema_late = None if ema_late is None else ema_late.detach()
        ctx.slot = slot
        if slot is not None:
            # Gradient hand-over outside autograd's dtype rules: wb (bf16) is marked
            # non-differentiable and a 1-element fp32 HANDLE carries the graph edge; the
            # contraction's backward leaves the fp32 weight gradient in `slot` (modconv_bmm), this
            # node picks it up.  (Through autograd the gradient of a bf16 tensor must be bf16:
            # an fp32 -> bf16 -> fp32 round trip of [B, O, K] per layer.)
            handle = torch.empty(1, device=slin.device, dtype=torch.float32)
            ctx.mark_non_differentiable(wb)
            ctx.set_materialize_grads(False)       # no zero-filled stand-in for wb's (absent) gradient
            return wb, handle
        return wb

    @staticmethod
    @once_differentiable
    def backward(ctx, gwb, ghandle=None):
        slin, w2, stats, rot = ctx.saved_tensors
        scale, demod, wshape, wdtype, c1, nf = ctx.cfg
        B, I = slin.shape
        O = w2.shape[0]
        if ctx.slot is not None:
            gwb = ctx.slot.pop("gw32")
        # this node may run on the weight bank's side stream while gwb was produced (and its
        # memory is owned) by the stream of the contraction's backward
        gwb.record_stream(torch.cuda.current_stream())
        gwb = _contig(gwb.float())
        dslin = torch.empty_like(slin)
        dw = torch.empty_like(w2)
        work = torch.empty(B * O + B * I + O * I + B + 1, device=slin.device, dtype=torch.float32)
        K.call("dusty_modprep_bwd", K.ptr(gwb), K.ptr(slin), K.ptr(w2), K.ptr(stats), K.ptr(dslin),
               K.ptr(dw), K.ptr(work), B, O, I, scale, 1 if demod else 0, K.ptr(rot), c1, nf,
               K.ptr(ctx.ema_late), K.stream_of(slin))
        return dslin, dw.reshape(wshape).to(wdtype), None, None, None, None, None, None, None, None


def cat_wb(wbs):
    """torch.cat(wbs, dim=1) for weights made with modprep(via_handle=True): the row ranges keep
    their gradient slots."""
    out = torch.cat(wbs, dim=1)
    parts, o0 = [], 0
    for w in wbs:
        for slot, handle, a, b in getattr(w, "_dusty_parts", []):
            parts.append((slot, handle, o0 + a, o0 + b))
        o0 += w.shape[1]
    if parts:
        out._dusty_parts = parts
    return out


def modprep(slin, weight, ema_var, scale: float, demod: bool, out_dtype=torch.float32, rot=None,
            c1: int = 0, ema_late=None, via_handle=False):
    """Per-sample effective weights wb[B,O,I] of a modulated 1x1 conv (see dusty_modprep_fwd).
    rot: optional [B, 2F] (cos | sin) rotation of the Fourier columns starting at c1.
    ema_late: the layer's ema_var buffer when the EMA normaliser is applied by the contraction
    (modconv_bmm(ema_var=...)) instead of being folded into wb (then ema_var must be None)."""
    K.require_cuda(slin, weight, ema_var, rot)
    if ema_late is not None and ema_var is not None:
        raise RuntimeError("modprep: ema_var and ema_late are exclusive")
    if via_handle:
        # wb for modconv_bmm only: its gradient reaches this node through a slot, in fp32
        slot = {}
        wb, handle = _ModPrep.apply(slin, weight, ema_var, float(scale), bool(demod), out_dtype, rot,
                                    int(c1), ema_late, slot)
        wb._dusty_parts = [(slot, handle, 0, wb.shape[1])]
        return wb
    return _ModPrep.apply(slin, weight, ema_var, float(scale), bool(demod), out_dtype, rot, int(c1),
                          ema_late)


def sumsq_total(x: torch.Tensor) -> torch.Tensor:
    """sum(x^2) as a 0-dim fp32 device tensor (no grad): ModConv2d's EMA statistic."""
    K.require_cuda(x)
    x = _contig(x.detach())
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    K.call("dusty_sumsq_rows", K.ptr(x), K.ptr(out), 1, x.numel(), 0, K.dtype_code(x),
           K.stream_of(x))
    return out[0]


def sumsq_buffer(x: torch.Tensor) -> torch.Tensor:
    """sum(x^2) as a 1-element fp32 device tensor (no grad)."""
    K.require_cuda(x)
    x = _canon(x.detach())
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    K.call("dusty_sumsq_rows", K.ptr(x), K.ptr(out), 1, x.numel(), 0, K.dtype_code(x),
           K.stream_of(x))
    return out


def ema_lerp_(ema_var: torch.Tensor, sum_a, sum_b, rep_b: float, numel: int, weight: float):
    """In place: ema_var <- lerp(ema_var, (sum_a + rep_b*sum_b)/numel, weight); sums are the
    1-element tensors of sumsq_buffer (either may be None)."""
    if ema_var.dtype != torch.float32 or not ema_var.is_cuda:
        raise RuntimeError("ema_var must be an fp32 CUDA tensor")
    K.call("dusty_ema_lerp", K.ptr(ema_var), K.ptr(sum_a), K.ptr(sum_b), float(rep_b),
           1.0 / float(numel), float(weight), K.stream_of(ema_var))


class _SumSqRows(Function):
    @staticmethod
    def forward(ctx, x):
        x = _contig(x)
        rows = x.shape[0]
        out = torch.empty(rows, device=x.device, dtype=torch.float32)
        K.call("dusty_sumsq_rows", K.ptr(x), K.ptr(out), rows, x.numel() // rows, 0,
               K.dtype_code(x), K.stream_of(x))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return x * (2.0 * g.reshape(-1, *([1] * (x.ndim - 1)))).to(x.dtype)


def sumsq_rows(x: torch.Tensor) -> torch.Tensor:
    """Per-sample sum of squares [B] (R1 penalty reduction, trainer.py:440)."""
    K.require_cuda(x)
    return _SumSqRows.apply(x)


# --------------------------------------------------------------------------- raydrop
class _Raydrop(Function):
    @staticmethod
    def forward(ctx, logit, image, u, rconst, temperature):
        logit, image, u = _contig(logit.float()), _contig(image.float()), _contig(u.float())
        mask = torch.empty_like(logit)
        out = torch.empty_like(logit)
        dsoft = torch.empty_like(logit)
        K.call("dusty_gumbel_raydrop_fwd", K.ptr(logit), K.ptr(image), K.ptr(u), K.ptr(mask),
               K.ptr(out), K.ptr(dsoft), None, logit.numel(), rconst, temperature,
               K.stream_of(logit))
        ctx.save_for_backward(image, mask, dsoft)
        ctx.rconst = rconst
        return mask, out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_mask, g_out):
        image, mask, dsoft = ctx.saved_tensors
        g_logit = torch.empty_like(image)
        g_image = torch.empty_like(image)
        g_out = None if g_out is None else _contig(g_out.float())
        g_mask = None if g_mask is None else _contig(g_mask.float())
        K.call("dusty_gumbel_raydrop_bwd", K.ptr(g_out), K.ptr(g_mask), K.ptr(image), K.ptr(mask),
               K.ptr(dsoft), K.ptr(g_logit), K.ptr(g_image), image.numel(), ctx.rconst,
               K.stream_of(image))
        return g_logit, g_image, None, None, None


def gumbel_raydrop(logit, image, u, rconst: float, temperature: float = 1.0):
    """-> (mask with straight-through gradient, lerp(image, rconst, 1 - mask))."""
    K.require_cuda(logit, image, u)
    return _Raydrop.apply(logit, image, u, float(rconst), float(temperature))


def raydrop_count(logit, image, u, rconst: float, temperature: float = 1.0):
    """No-grad variant that also returns the integer number of kept rays."""
    K.require_cuda(logit, image, u)
    logit, image, u = _contig(logit.float()), _contig(image.float()), _contig(u.float())
    mask, out, dsoft = torch.empty_like(logit), torch.empty_like(logit), torch.empty_like(logit)
    count = torch.zeros(1, device=logit.device, dtype=torch.int32)
    K.call("dusty_gumbel_raydrop_fwd", K.ptr(logit), K.ptr(image), K.ptr(u), K.ptr(mask),
           K.ptr(out), K.ptr(dsoft), K.ptr(count), logit.numel(), float(rconst),
           float(temperature), K.stream_of(logit))
    return mask, out, count


# --------------------------------------------------------------------------- projection
def point_project(x: torch.Tensor, trig: torch.Tensor, min_depth: float, max_depth: float,
                  tol: float = 1e-11, point_set: bool = False):
    """x [B,1,H,W] inverse-depth-normalised -> (points, valid_count[int64 tensor]).
    trig: [4, H*W] = cos el, sin el, cos az, sin az."""
    K.require_cuda(x, trig)
    x = _contig(x.detach().float())
    B, _, H, W = x.shape
    HW = H * W
    out = torch.empty((B, HW, 3) if point_set else (B, 3, H, W), device=x.device,
                      dtype=torch.float32)
    count = torch.zeros(1, device=x.device, dtype=torch.int64)
    K.call("dusty_point_project", K.ptr(x), K.ptr(trig), K.ptr(out), K.ptr(count), B, HW,
           float(min_depth), float(max_depth), float(tol), 1 if point_set else 0, K.stream_of(x))
    return out, count


# --------------------------------------------------------------------------- minibatch stddev
def _mbstd_composite_grad(gy, x, group, alpha):
    """Analytic backward written with differentiable torch ops (used only when a second
    derivative is being recorded, i.e. the R1 step)."""
    B, C, H, W = x.shape
    G = min(B, group)
    M = B // G
    xf = x.float().reshape(G, M, C, H, W)
    mean = xf.mean(0, keepdim=True)
    d = xf - mean
    inv_sd = torch.rsqrt(d.pow(2).mean(0, keepdim=True) + alpha)
    gstat = gy[:, C].float().reshape(G, M, H * W).sum(dim=(0, 2))      # [M]
    coef = gstat.reshape(1, M, 1, 1, 1) / float(C * H * W * G)
    gx = gy[:, :C].float().reshape(G, M, C, H, W) + coef * d * inv_sd
    return gx.reshape(B, C, H, W).to(x.dtype)


class _MbStd(Function):
    @staticmethod
    def forward(ctx, x, group, alpha):
        x = _contig(x)
        B, C, H, W = x.shape
        G = min(B, group)
        y = torch.empty(B, C + 1, H, W, device=x.device, dtype=x.dtype)
        stat = torch.empty(B // G, device=x.device, dtype=torch.float32)
        K.call("dusty_minibatch_std_fwd", K.ptr(x), K.ptr(y), K.ptr(stat), B, C, H * W, group,
               alpha, K.dtype_code(x), K.stream_of(x))
        ctx.save_for_backward(x)
        ctx.group, ctx.alpha = group, alpha
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        if torch.is_grad_enabled():   # create_graph=True: stay differentiable
            return _mbstd_composite_grad(gy, x, ctx.group, ctx.alpha), None, None
        gy = _contig(gy)
        B, C, H, W = x.shape
        G = min(B, ctx.group)
        gx = torch.empty_like(x)
        dstat = torch.empty(B // G, device=x.device, dtype=torch.float32)
        K.call("dusty_minibatch_std_bwd", K.ptr(gy), K.ptr(x), K.ptr(gx), K.ptr(dstat), B, C,
               H * W, ctx.group, ctx.alpha, K.dtype_code(x), K.stream_of(x))
        return gx, None, None


def minibatch_stddev(x, group: int = 4, alpha: float = 1e-8):
    K.require_cuda(x)
    B = x.shape[0]
    G = min(B, group)
    if B % G != 0:
        raise RuntimeError(f"batch {B} is not divisible by the group size {G}")
    return _MbStd.apply(x, int(group), float(alpha))


class _MbStdStat(Function):
    """stat[m] of MinibatchStdDev only (the appended channel is constant over C, H, W, so the
    consumer can fold it in analytically); x in any dense per-sample layout."""

    @staticmethod
    def forward(ctx, x, group, alpha):
        B, C, H, W = x.shape
        G = min(B, group)
        stat = torch.empty(B // G, device=x.device, dtype=torch.float32)
        K.call("dusty_minibatch_std_fwd", K.ptr(x), None, K.ptr(stat), B, C, H * W, group, alpha,
               K.dtype_code(x), K.stream_of(x))
        ctx.save_for_backward(x)
        ctx.group, ctx.alpha = group, alpha
        return stat

    @staticmethod
    def backward(ctx, gstat):
        (x,) = ctx.saved_tensors
        B, C, H, W = x.shape
        G = min(B, ctx.group)
        M = B // G
        if torch.is_grad_enabled():   # create_graph=True (R1): differentiable composite
            xf = x.float().reshape(G, M, C, H, W)
            d = xf - xf.mean(0, keepdim=True)
            inv_sd = torch.rsqrt(d.pow(2).mean(0, keepdim=True) + ctx.alpha)
            coef = gstat.float().reshape(1, M, 1, 1, 1) / float(C * H * W * G)
            gx = (coef * d * inv_sd).reshape(B, C, H, W).to(x.dtype)
            return (gx.contiguous(memory_format=torch.channels_last) if _is_cl(x) else gx), None, None
        gx = torch.empty_like(x)
        K.call("dusty_minibatch_std_bwd", None, K.ptr(x), K.ptr(gx), K.ptr(gstat.float().contiguous()),
               B, C, H * W, ctx.group, ctx.alpha, K.dtype_code(x), K.stream_of(x))
        return gx, None, None


def minibatch_std_stat(x, group: int = 4, alpha: float = 1e-8):
    """Per-sample value of MinibatchStdDev's appended channel, [B] fp32 (sample b of a group
    slot m = b % (B/G) carries stat[m], common.py:237-250)."""
    K.require_cuda(x)
    B = x.shape[0]
    G = min(B, group)
    if B % G != 0:
        raise RuntimeError(f"batch {B} is not divisible by the group size {G}")
    if not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)):
        x = x.contiguous()
    return _MbStdStat.apply(x, int(group), float(alpha)).repeat(G)


# --------------------------------------------------------------------------- circular un-shift
class _CircShift(Function):
    @staticmethod
    def forward(ctx, v, shift01, scale, adjoint):
        v = _contig(v.float())
        B, C, H, W = v.shape
        out = torch.empty_like(v)
        K.call("dusty_circular_shift", K.ptr(v), K.ptr(shift01), K.ptr(out), B, C, H, W, scale,
               adjoint, K.stream_of(v))
        ctx.save_for_backward(shift01)
        ctx.scale, ctx.adjoint = scale, adjoint
        return out

    @staticmethod
    def backward(ctx, g):
        (shift01,) = ctx.saved_tensors
        return _CircShift.apply(g, shift01, ctx.scale, 1 - ctx.adjoint), None, None, None


def circular_unshift(v, shift01, scale: float = 1.0):
    """v [B,C,H,W] fp32, shift01 [B] fp32 in [0,1) (fraction of a full turn)."""
    K.require_cuda(v, shift01)
    return _CircShift.apply(v, _contig(shift01.detach().float()), float(scale), 0)


# ---- a11: EqualLR weight preparation (scale + cast + OIHW -> OHWI), one kernel each way -------
class _WeightPrep(Function):
    """fp32 master filter [O, C, R, S] -> scaled filter in `dtype`, channels_last memory, plus
    (non-differentiable second output) the same filter as [R*S][C][O] for the data-gradient
    kernels -- both written by one launch.  A linear map: its backward is the adjoint kernel,
    whose backward is this kernel again, so second-order terms (R1) flow through exactly."""

    @staticmethod
    def forward(ctx, w, scale, dtype, with_tco):
        O, C, R, S = w.shape
        wf = _contig(w.detach().float())
        out = torch.empty((O, C, R, S), dtype=dtype, device=w.device, memory_format=torch.channels_last)
        tco = torch.empty((R * S, C, O), dtype=dtype, device=w.device) if with_tco else None
        K.call("dusty_weight_prep", K.ptr(wf), K.ptr(out), K.ptr(tco), O, C, R * S, scale,
               K.dtype_code(out), K.stream_of(wf))
        ctx.cfg = (scale, w.dtype)
        if with_tco:
            ctx.mark_non_differentiable(tco)
            return out, tco
        return out

    @staticmethod
    def backward(ctx, g, *unused):
        scale, wdtype = ctx.cfg
        return _WeightPrepAdj.apply(g, scale).to(wdtype), None, None, None


class _WeightPrepAdj(Function):
    @staticmethod
    def forward(ctx, g, scale):
        O, C, R, S = g.shape
        nhwc = 1
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        if g.is_contiguous(memory_format=torch.channels_last):
            gsrc = g
        elif g.is_contiguous():
            gsrc, nhwc = g, 0
        else:
            gsrc = g.contiguous(memory_format=torch.channels_last)
        gw = torch.empty((O, C, R, S), dtype=torch.float32, device=g.device)
        K.call("dusty_weight_prep_adj", K.ptr(gsrc), K.ptr(gw), O, C, R * S, scale, K.dtype_code(gsrc), nhwc,
               K.stream_of(gsrc))
        ctx.cfg = (scale, g.dtype)
        return gw

    @staticmethod
    def backward(ctx, gg):
        scale, gdtype = ctx.cfg
        return _WeightPrep.apply(gg, scale, gdtype, False), None


def prep_conv_weight(w: torch.Tensor, scale: float, dtype: torch.dtype, with_tco: bool = False):
    """EqualLR-scaled convolution filter in `dtype` and channels_last memory; with_tco=True
    also returns its [R*S][C][O] form (see _WeightPrep)."""
    K.require_cuda(w)
    return _WeightPrep.apply(w, float(scale), dtype, bool(with_tco))


# ---- a11: discriminator stem (BlurVH -> 1x1 conv 2 -> O -> bias + leaky ReLU), stem.cu -------
class _Stem(Function):
    """y (bf16, NHWC) = lrelu(conv1x1(cat(blur_v(x), blur_h(x)), w) + bias) * gain in one pass.
    First order: one fused backward pass (+ a small adjoint-blur kernel when x needs a
    gradient).  Under create_graph=True (the R1 penalty) the backward is re-expressed through
    `composite`, the same math built from the differentiable ops of this module, so every
    higher-order term stays exact."""

    @staticmethod
    def forward(ctx, x, w, bias, taps, alpha, scale, composite):
        B, _, H, W = x.shape
        O = w.shape[0]
        xin = _contig(x.detach())
        if xin.dtype not in (torch.float32, torch.bfloat16):
            xin = xin.float()
        wf = _contig(w.detach().float().reshape(O, 2))
        bf = None if bias is None else _contig(bias.detach().float().reshape(O))
        y = torch.empty((B, O, H, W), dtype=torch.bfloat16, device=x.device,
                        memory_format=torch.channels_last)
        K.call("dusty_stem_fwd", K.ptr(xin), K.ptr(wf), K.ptr(bf), K.ptr(y), B, H, W, O, taps[0], taps[1],
               taps[2], alpha, scale, K.dtype_code(xin), K.stream_of(xin))
        ctx.save_for_backward(x, w, bias, y)
        ctx.cfg = (taps, alpha, scale, composite)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, bias, y = ctx.saved_tensors
        taps, alpha, scale, composite = ctx.cfg
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        if torch.is_grad_enabled():
            with torch.enable_grad():
                wanted = [t for t, n in ((x, need_x), (w, need_w), (bias, need_b)) if n and t is not None]
                yc = composite(x, w, bias)
                grads = list(torch.autograd.grad(yc, wanted, gy.to(yc.dtype), create_graph=True,
                                                 allow_unused=True))
            out = []
            for t, n in ((x, need_x), (w, need_w), (bias, need_b)):
                out.append(grads.pop(0) if (n and t is not None) else None)
            return out[0], out[1], out[2], None, None, None, None
        B, _, H, W = x.shape
        O = w.shape[0]
        gy = gy.contiguous(memory_format=torch.channels_last)
        if gy.dtype != torch.bfloat16:
            gy = gy.to(torch.bfloat16)
        xin = _contig(x.detach())
        if xin.dtype not in (torch.float32, torch.bfloat16):
            xin = xin.float()
        wf = _contig(w.detach().float().reshape(O, 2))
        dvh = torch.empty((B, 2, H, W), dtype=torch.float32, device=x.device) if need_x else None
        dwb = torch.empty((O, 3), dtype=torch.float32, device=x.device)
        st = K.stream_of(gy)
        K.call("dusty_stem_bwd", K.ptr(gy), K.ptr(y), K.ptr(xin), K.ptr(wf), K.ptr(dvh), K.ptr(dwb), B, H, W,
               O, taps[0], taps[1], taps[2], alpha, scale, K.dtype_code(xin), st)
        gx = None
        if need_x:
            gx = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
            K.call("dusty_stem_dx", K.ptr(dvh), K.ptr(gx), B, H, W, taps[0], taps[1], taps[2], st)
            gx = gx.to(x.dtype)
        gw = dwb[:, :2].reshape(w.shape).to(w.dtype) if need_w else None
        gb = dwb[:, 2].reshape(bias.shape).to(bias.dtype) if (need_b and bias is not None) else None
        return gx, gw, gb, None, None, None, None


def stem_supported(x: torch.Tensor, out_ch: int) -> bool:
    return (x.is_cuda and x.dim() == 4 and x.shape[1] == 1 and x.shape[2] >= 2 and x.shape[3] >= 2
            and out_ch in (8, 16, 32, 64) and _PRECISION["act"] == torch.bfloat16)


def stem(x, w, bias, taps, composite, negative_slope: float = 0.2, scale: float = 2 ** 0.5):
    """Fused discriminator stem; `composite(x, w, bias)` must compute the same function from
    differentiable ops (used only for second-order gradients)."""
    K.require_cuda(x, w, bias)
    return _Stem.apply(x, w, bias, tuple(float(t) for t in taps), float(negative_slope), float(scale),
                       composite)


# ---- a11: dense NHWC convolutions on tcgen05 (conv_tc.cu) --------------------------------
_CONV_IMPL = {"halo": True}


def set_conv_halo(enabled: bool):
    """Allow / forbid the halo-resident kernel (tests exercise both formulations)."""
    _CONV_IMPL["halo"] = bool(enabled)


def conv_tc_supported(x: torch.Tensor, w: torch.Tensor, stride, op: str = "fprop") -> bool:
    """Domain of the tcgen05 kernels: bf16 NHWC activations, channel counts that are multiples
    of 8 (16-byte TMA strides), filters of at most 4x4 taps, strides 1 or 2.  Everything else
    runs on the CUDA-core kernels of the same family (conv_simt.cu); there is no library path."""
    if not x.is_cuda or x.dim() != 4:
        return False
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        return False
    O, C, R, S = w.shape
    if C != x.shape[1] or C % 8 or O % 8 or R > 4 or S > 4:
        return False
    if stride[0] not in (1, 2) or stride[1] not in (1, 2):
        return False
    if x.shape[2] < R or x.shape[3] < S:
        return False
    return True


def conv_halo_ok(w: torch.Tensor, op: str) -> bool:
    """Unit-stride fprop / dgrad of a thin layer: the halo-resident kernel (each input pixel
    landed in shared memory once) applies when the contracted channel count is 32 or 64."""
    if not _CONV_IMPL["halo"]:
        return False
    O, C, R, S = w.shape
    cin, cout = (C, O) if op == "fprop" else (O, C)
    bn = 128 if cout > 64 else (64 if cout > 32 else 32)      # resident filter: <= 72 KiB
    return (R <= 3 and S <= 3 and cin in (32, 64) and cout % 8 == 0 and 8 <= cout <= 128
            and R * S * bn * cin * 2 <= 72 * 1024)


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous(memory_format=torch.channels_last)


def _ints(vals):
    return (K.C.c_int * len(vals))(*vals)


def _ohwi(w: torch.Tensor) -> torch.Tensor:
    """The filter with OHWI memory ([O][R][S][C], what channels_last gives): no copy when it is
    already laid out that way (prepared weights are)."""
    return w if w.is_contiguous(memory_format=torch.channels_last) else \
        w.contiguous(memory_format=torch.channels_last)


def filter_tco(w: torch.Tensor) -> torch.Tensor:
    """[R*S][C][O] copy of a filter (O contiguous): the layout of the data-gradient kernels.
    Prepared weights carry it already (one kernel makes both layouts); this is the fallback."""
    O, C, R, S = w.shape
    return w.permute(2, 3, 1, 0).reshape(R * S, C, O).contiguous()


def conv2d_fprop_tc(x, w, stride, bias=None, act: int = 1, alpha: float = 0.2, scale: float = 1.0,
                    out_dtype=None):
    """y = conv2d(x, w, stride) (valid, no padding) [+ bias, leaky-ReLU, scale]; NHWC in/out.
    The filter is read in place from its OHWI memory through strided tensor maps.
    out_dtype=torch.float32: fp32 output (split-bf16 operands of the fp32 mode)."""
    K.require_cuda(x, w)
    x = _nhwc(x)
    B, C, H, W = x.shape
    O, _, R, S = w.shape
    sh, sw = stride
    Ho, Wo = (H - R) // sh + 1, (W - S) // sw + 1
    out_dtype = out_dtype or x.dtype
    y = torch.empty((B, O, Ho, Wo), dtype=out_dtype, device=x.device,
                    memory_format=torch.channels_last)
    wo = _ohwi(w)
    if sh == 1 and sw == 1 and out_dtype == x.dtype and conv_halo_ok(w, "fprop"):
        # element (t, n, c) at  n * (R*S*C) + t * C + c
        K.call("dusty_conv2d_halo_tc", K.ptr(x), K.ptr(wo), K.ptr(bias), K.ptr(y), B, H, W, C,
               Ho, Wo, O, R, S, 0, 0, 0, Ho * Wo * O, Wo * O, O, act, alpha, scale,
               R * S * C, C, 0, K.stream_of(x))
        return y
    # window mode: group r, element (n, s*C + c) at  n * (R*S*C) + r * (S*C) + s*C + c
    K.call("dusty_conv2d_tc", K.ptr(x), K.ptr(wo), K.ptr(bias), K.ptr(y),
           B, H, W, C, Ho, Wo, O, 1, R, _ints(list(range(R))), _ints([0] * R), S, sh, sw,
           0, Ho * Wo * O, Wo * O, O, act, alpha, scale, R * S * C, S * C, None, 0, K.dtype_code(y),
           K.stream_of(x))
    return y


def conv2d_dgrad_tc(gy, w, stride, in_hw, w_tco=None, out_dtype=None, w_shape=None):
    """Gradient of the valid convolution w.r.t. its input ([B, C, H, W] NHWC).  Every variant
    (halo-resident, unit stride, the parity classes of a strided convolution) reads the same
    [R*S][C][O] filter tensor `w_tco` through tap-index maps: no flipping / stacking copies."""
    K.require_cuda(gy, w if w is not None else w_tco)
    gy = _nhwc(gy)
    B, O, Ho, Wo = gy.shape
    _, C, R, S = w_shape if w is None else w.shape        # w=None: only the [R*S][C][O] form exists
    H, W = in_hw
    sh, sw = stride
    if w_tco is None or w_tco.dtype != gy.dtype or tuple(w_tco.shape) != (R * S, C, O):
        w_tco = filter_tco(w)
    out_dtype = out_dtype or gy.dtype
    gx = torch.empty((B, C, H, W), dtype=out_dtype, device=gy.device,
                     memory_format=torch.channels_last)
    st = K.stream_of(gy)
    if sh == 1 and sw == 1 and out_dtype == gy.dtype and w is not None and conv_halo_ok(w, "dgrad"):
        K.call("dusty_conv2d_halo_tc", K.ptr(gy), K.ptr(w_tco), None, K.ptr(gx), B, Ho, Wo, O, H, W,
               C, R, S, -(R - 1), -(S - 1), 0, H * W * C, W * C, C, 1, 0.0, 1.0, 0, 0, 1, st)
        return gx
    # strided: one class per output parity (ph, pw), all of them in ONE launch
    cls_G, dh, dw, wtap, cls_H, cls_W, cls_off = [], [], [], [], [], [], []
    empty_class = False
    for ph in range(sh):
        rs = [r for r in range(R) if r % sh == ph]
        for pw in range(sw):
            ss = [s for s in range(S) if s % sw == pw]
            Hc, Wc = (H - ph + sh - 1) // sh, (W - pw + sw - 1) // sw
            if not (rs and ss) or Hc <= 0 or Wc <= 0:
                empty_class = empty_class or (Hc > 0 and Wc > 0)
                continue
            taps = [(r, s) for r in rs for s in ss]
            cls_G.append(len(taps))
            dh += [(ph - r) // sh for r, _ in taps]
            dw += [(pw - s) // sw for _, s in taps]
            wtap += [r * S + s for r, s in taps]
            cls_H.append(Hc)
            cls_W.append(Wc)
            cls_off.append((ph * W + pw) * C)
    if empty_class:                      # positions no filter tap reaches (e.g. 1x1, stride 2)
        gx.zero_()
    if len(cls_G) > 2:
        # a CTA walks the classes of a patch in order and its range may end between them: put
        # the heaviest and the lightest class side by side (3x3 stride 2: 4, 1, 2, 2 taps)
        order = sorted(range(len(cls_G)), key=lambda c: -cls_G[c])
        order = [order[0], order[-1]] + order[1:-1]
        starts = [sum(cls_G[:c]) for c in range(len(cls_G))]
        pick = lambda v: [x for c in order for x in v[starts[c]:starts[c] + cls_G[c]]]   # noqa: E731
        dh, dw, wtap = pick(dh), pick(dw), pick(wtap)
        cls_H, cls_W, cls_off = [cls_H[c] for c in order], [cls_W[c] for c in order], [cls_off[c] for c in order]
        cls_G = [cls_G[c] for c in order]
    if cls_G:
        K.call("dusty_conv2d_tc_classes", K.ptr(gy), K.ptr(w_tco), K.ptr(gx), B, Ho, Wo, O, C,
               len(cls_G), _ints(cls_G), _ints(dh), _ints(dw), _ints(wtap), _ints(cls_H), _ints(cls_W),
               (K.C.c_longlong * len(cls_off))(*cls_off), H * W * C, sh * W * C, sw * C, 0, 0, R * S,
               K.dtype_code(gx), st)
    return gx


def conv2d_wgrad_tc(gy, x, stride, w_shape, out_dtype):
    """Gradient of the valid convolution w.r.t. its filter ([O, C, R, S], contiguous)."""
    K.require_cuda(gy, x)
    gy, x = _nhwc(gy), _nhwc(x)
    B, C, H, W = x.shape
    O, _, R, S = w_shape
    _, _, Ho, Wo = gy.shape
    sh, sw = stride
    n_ws = int(K.load().dusty_conv2d_wgrad_tc_workspace(B, Ho, Wo, C, O, R, S))
    ws = torch.empty(max(n_ws, 1), dtype=torch.float32, device=x.device)
    dwp = torch.empty((R, S, C, O), dtype=torch.float32, device=x.device)
    K.call("dusty_conv2d_wgrad_tc", K.ptr(x), K.ptr(gy), K.ptr(dwp), K.ptr(ws),
               n_ws, B, H, W, C, Ho, Wo, O, R, S, sh, sw, K.stream_of(x))
    if out_dtype in (torch.float32, torch.bfloat16):
        # [R,S,C,O] fp32 -> [O,C,R,S] in out_dtype with OHWI (channels_last) memory, one kernel
        gw = torch.empty((O, C, R, S), dtype=out_dtype, device=x.device, memory_format=torch.channels_last)
        K.call("dusty_filter_rsco_to_ohwi", K.ptr(dwp), K.ptr(gw), O, C, R * S, K.dtype_code(gw),
               K.stream_of(x))
        return gw
    return dwp.permute(3, 2, 0, 1).to(out_dtype).contiguous()


# ---- g1: fp32 convolutions on the tensor cores (split-bf16 operands, csrc/split3.cu) -----------
def _ll3(vals):
    return (K.C.c_longlong * 3)(*[int(v) for v in vals])


def split_bf16x3(src: torch.Tensor, dst: torch.Tensor, outer, k, inner, src_strides, dst_strides,
                 pattern: int):
    """dst[o, p*K + k, i] = part p of src[o, k, i] (hi,hi,lo for pattern 0; hi,lo,hi for 1)."""
    K.require_cuda(src, dst)
    if src.dtype != torch.float32 or dst.dtype != torch.bfloat16:
        raise RuntimeError("split_bf16x3: fp32 source, bf16 destination")
    K.call("dusty_split_bf16x3", K.ptr(src), K.ptr(dst), outer, k, inner, _ll3(src_strides),
           _ll3(dst_strides), pattern, K.stream_of(src))
    return dst


def _collapsible_hw(t: torch.Tensor) -> bool:
    return t.stride(2) == t.shape[3] * t.stride(3)


def _split_act(x: torch.Tensor, pattern: int) -> torch.Tensor:
    """fp32 [B, C, H, W] (NCHW or NHWC memory) -> bf16 NHWC [B, 3C, H, W]."""
    if not _collapsible_hw(x):
        x = x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    out = torch.empty((B, 3 * C, H, W), dtype=torch.bfloat16, device=x.device,
                      memory_format=torch.channels_last)
    return split_bf16x3(x, out, B, C, H * W, (x.stride(0), x.stride(1), x.stride(3)),
                        (H * W * 3 * C, 1, 3 * C), pattern)


def _split_batch(x: torch.Tensor, pattern: int) -> torch.Tensor:
    """fp32 [B, C, H, W] -> bf16 NHWC [3B, C, H, W]: the three terms stacked on the batch axis
    (the weight gradient contracts over batch x pixels)."""
    x = x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    out = torch.empty((3 * B, C, H, W), dtype=torch.bfloat16, device=x.device,
                      memory_format=torch.channels_last)
    n = x.numel()
    return split_bf16x3(x, out, 1, 1, n, (0, 0, 1), (0, n, 1), pattern)


def conv_x3_supported(x: torch.Tensor, w: torch.Tensor, stride) -> bool:
    """fp32 tensors whose split form lies in the tcgen05 kernels' domain."""
    if not (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and w.dtype == torch.float32):
        return False
    O, C, R, S = w.shape
    if C != x.shape[1] or C % 8 or O % 8 or R > 4 or S > 4:
        return False
    if stride[0] not in (1, 2) or stride[1] not in (1, 2):
        return False
    return x.shape[2] >= R and x.shape[3] >= S


def conv2d_fprop_x3(x, w, stride):
    """fp32 y = conv2d(x, w, stride) (valid) on tcgen05: [x_hi|x_hi|x_lo] * [w_hi|w_lo|w_hi]."""
    O, C, R, S = w.shape
    xs = _split_act(x, 0)
    ws = torch.empty((O, 3 * C, R, S), dtype=torch.bfloat16, device=w.device,
                     memory_format=torch.channels_last)
    # w[o, c, r, s] -> OHWI [o][t][3C]: outer = O, K = C, inner = R*S taps
    wc = w.contiguous()
    split_bf16x3(wc, ws, O, C, R * S, (C * R * S, R * S, 1), (R * S * 3 * C, 1, 3 * C), 1)
    return conv2d_fprop_tc(xs, ws, stride, out_dtype=torch.float32)


def conv2d_dgrad_x3(gy, w, stride, in_hw):
    """fp32 gradient w.r.t. the input: contraction over (tap, O) with the O axis tripled."""
    O, C, R, S = w.shape
    gs = _split_act(gy, 0)
    # [R*S][C][3*O] from w[o, c, r, s]: outer = taps (stride 1), K = O, inner = C
    wc = w.contiguous()
    w_tco = torch.empty((R * S, C, 3 * O), dtype=torch.bfloat16, device=w.device)
    split_bf16x3(wc, w_tco, R * S, O, C, (1, C * R * S, R * S), (C * 3 * O, 1, 3 * O), 1)
    return conv2d_dgrad_tc(gs, None, stride, in_hw, w_tco, out_dtype=torch.float32,
                           w_shape=(3 * O, C, R, S))


def conv2d_wgrad_x3(gy, x, stride, w_shape):
    """fp32 gradient w.r.t. the filter: the three products as three batches of one launch."""
    return conv2d_wgrad_tc(_split_batch(gy, 0), _split_batch(x, 1), stride, w_shape, torch.float32)


# ---- a11 / a15: the same family on the CUDA cores (conv_simt.cu): fp32 parity mode, odd shapes --
def _ll4(vals):
    return (K.C.c_longlong * 4)(*[int(v) for v in vals])


def _simt_dtype(*ts):
    """Common element type of the operands: bf16 only if all are, else fp32."""
    return torch.bfloat16 if all(t.dtype == torch.bfloat16 for t in ts) else torch.float32


def _empty_like_layout(shape, ref, dtype):
    fmt = torch.channels_last if _is_cl(ref) else torch.contiguous_format
    return torch.empty(shape, dtype=dtype, device=ref.device, memory_format=fmt)


def conv_fans_out_to_nhwc(x, w) -> bool:
    """The 1x1 convolution that lifts a few-channel NCHW image into the NHWC bf16 feature stack
    (the discriminator stem's 2 -> C0 layer as single ops: functional._Stem's composite, which
    the R1 double backward runs): its OUTPUT and the gradients w.r.t. that output are NHWC like
    every tensor downstream -- the strided CUDA-core kernels read / write either layout, and an
    NCHW result here cost four 134 MB layout copies per R1 iteration further down the chain."""
    return (x.dim() == 4 and w.dim() == 4 and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
            and x.shape[1] <= 4 and w.shape[2] == 1 and w.shape[3] == 1 and w.shape[0] % 8 == 0
            and _PRECISION["act"] == torch.bfloat16)


def conv2d_fprop_simt(x, w, stride, padding=(0, 0)):
    """y = conv2d(x, w, stride, zero padding); any layout, fp32 or bf16 (fp32 accumulation)."""
    K.require_cuda(x, w)
    dt = _simt_dtype(x, w)
    fan_out = conv_fans_out_to_nhwc(x, w)
    x, w = x.detach().to(dt), w.detach().to(dt)
    B, C, H, W = x.shape
    O, _, R, S = w.shape
    sh, sw = stride
    ph, pw = padding
    Ho, Wo = (H + 2 * ph - R) // sh + 1, (W + 2 * pw - S) // sw + 1
    if fan_out:
        y = torch.empty((B, O, Ho, Wo), dtype=dt, device=x.device, memory_format=torch.channels_last)
    else:
        y = _empty_like_layout((B, O, Ho, Wo), x, dt)
    K.call("dusty_conv2d_simt", 0, K.ptr(x), None, K.ptr(w), K.ptr(y), B, C, H, W, O, Ho, Wo, R, S, sh, sw,
           ph, pw, _ll4(x.stride()), _ll4(y.stride()), _ll4(w.stride()), 1.0, K.dtype_code(y), K.stream_of(x))
    return y


def conv2d_dgrad_simt(gy, w, stride, padding, in_hw, like=None):
    """Data gradient of the convolution above = conv_transpose2d(gy, w) cropped to in_hw."""
    K.require_cuda(gy, w)
    dt = _simt_dtype(gy, w)
    gy, w = gy.detach().to(dt), w.detach().to(dt)
    B, O, Ho, Wo = gy.shape
    _, C, R, S = w.shape
    H, W = in_hw
    gx = _empty_like_layout((B, C, H, W), gy if like is None else like, dt)
    K.call("dusty_conv2d_simt", 1, None, K.ptr(gy), K.ptr(w), K.ptr(gx), B, C, H, W, O, Ho, Wo, R, S,
           stride[0], stride[1], padding[0], padding[1], _ll4(gx.stride()), _ll4(gy.stride()),
           _ll4(w.stride()), 1.0, K.dtype_code(gx), K.stream_of(gy))
    return gx


def conv2d_wgrad_simt(gy, x, stride, padding, w_shape):
    """Filter gradient, fp32 [O, C, R, S] contiguous."""
    K.require_cuda(gy, x)
    dt = _simt_dtype(gy, x)
    gy, x = gy.detach().to(dt), x.detach().to(dt)
    B, C, H, W = x.shape
    O, _, R, S = w_shape
    _, _, Ho, Wo = gy.shape
    gw = torch.empty((O, C, R, S), dtype=torch.float32, device=x.device)
    K.call("dusty_conv2d_simt", 2, K.ptr(x), K.ptr(gy), None, K.ptr(gw), B, C, H, W, O, Ho, Wo, R, S,
           stride[0], stride[1], padding[0], padding[1], _ll4(x.stride()), _ll4(gy.stride()),
           _ll4((C * R * S, R * S, S, 1)), 1.0, K.dtype_code(x), K.stream_of(x))
    return gw


# ---- a11: linears of the discriminator epilogue on own GEMMs (linear_tc.cu) -----------------------
def _major(t, mult):
    """(mn_major flag, leading dimension) of a 2-D operand [rows, K] for the GEMM kernels, or None
    when it has to be copied first (`mult`: leading dimensions are multiples of 16 bytes)."""
    if t.stride(1) == 1 and t.stride(0) % mult == 0 and t.stride(0) >= t.shape[1]:
        return 0, t.stride(0)
    if t.stride(0) == 1 and t.stride(1) % mult == 0 and t.stride(1) >= t.shape[0]:
        return 1, t.stride(1)
    return None


def matmul_nt(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """alpha * a @ b.T -> fp32 [M, N] for CUDA matrices a [M, K], b [N, K] (either may be a
    transposed view: both majors are read in place).  fp32 operands with K contiguous run on
    tcgen05 kind::tf32 (the forward of the 65536 -> 512 linear: the fp32 master weight is read
    once, uncast); any other combination on kind::f16 with bf16 operands (the gradients: the same
    tensors as MN-major operands); tiny outputs on the CUDA-core kernel."""
    K.require_cuda(a, b)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise RuntimeError(f"matmul_nt: shapes {tuple(a.shape)} x {tuple(b.shape)}^T")
    a, b = a.detach(), b.detach()
    M, Kd = a.shape
    N = b.shape[0]
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    st = K.stream_of(a)
    if min(M, N) < 16 or N % 4 or Kd < 16 or (M * N <= (1 << 16) and Kd < 1024):
        a32 = a if a.dtype == torch.float32 else a.float()
        b32 = b if b.dtype == torch.float32 else b.float()
        K.call("dusty_gemm_simt", K.ptr(a32), K.ptr(b32), K.ptr(c), M, N, Kd, a32.stride(0), a32.stride(1),
               b32.stride(0), b32.stride(1), N, 1, float(alpha), st)
        return c
    if a.dtype == torch.float32 and b.dtype == torch.float32:
        ma, mb = _major(a, 4), _major(b, 4)
        if ma is not None and mb is not None and ma[0] == 0 and mb[0] == 0 and a.data_ptr() % 16 == 0 \
                and b.data_ptr() % 16 == 0:
            K.call("dusty_gemm_tf32", K.ptr(a), K.ptr(b), K.ptr(c), M, N, Kd, ma[1], mb[1], N, float(alpha), 0, st)
            return c
    a16 = a if a.dtype == torch.bfloat16 else a.to(torch.bfloat16)      # strides are preserved
    b16 = b if b.dtype == torch.bfloat16 else b.to(torch.bfloat16)
    ma, mb = _major(a16, 8), _major(b16, 8)
    if ma is None or a16.data_ptr() % 16:
        a16 = a16.contiguous()
        ma = _major(a16, 8) or (0, a16.stride(0))
    if mb is None or b16.data_ptr() % 16:
        b16 = b16.contiguous()
        mb = _major(b16, 8) or (0, b16.stride(0))
    if ma[1] % 8 or mb[1] % 8:
        raise RuntimeError("matmul_nt: operand rows must be multiples of 8 elements for the bf16 GEMM")
    K.call("dusty_gemm_bf16", K.ptr(a16), K.ptr(b16), K.ptr(c), M, N, Kd, ma[0], mb[0], ma[1], mb[1], N,
           float(alpha), 0, st)
    return c


class _LinearNT(Function):
    """y = x @ w.T (bilinear: every derivative of every order is again a product of this form,
    so the R1 double backward runs on the same kernels)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        if x.dtype != w.dtype and w.dtype == torch.float32:
            return matmul_nt(x.float(), w)     # bf16 activations, fp32 master weight: TF32 on the weight as is
        return matmul_nt(x, w)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = gw = None
        if torch.is_grad_enabled():            # create_graph=True: keep it differentiable
            if ctx.needs_input_grad[0]:
                gx = _LinearNT.apply(gy, w.t()).to(x.dtype)
            if ctx.needs_input_grad[1]:
                gw = _LinearNT.apply(gy.t(), x.t()).to(w.dtype)
            return gx, gw
        if ctx.needs_input_grad[0]:
            gx = matmul_nt(gy, w.t()).to(x.dtype)
        if ctx.needs_input_grad[1]:
            gw = matmul_nt(gy.t(), x.t()).to(w.dtype)
        return gx, gw


def linear_nt(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x [B, K] @ w [N, K].T -> fp32 on this package's GEMM kernels, with autograd of any order."""
    return _LinearNT.apply(x, w)

