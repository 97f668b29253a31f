"""Drop-in for gans/models/ops/upfirdn2d/upfirdn2d.py (reference lines 20-164).

`upfirdn2d(input[N,C,H,W], kernel[kh,kw], up, down, pad)` with the reference's x-before-y
argument order; `upfirdn2d_op.upfirdn2d` mirrors the pybind entry
(upfirdn2d.cpp:17-31: input [major, in_h, in_w, minor]).  Backward and double backward
run through the adjoint / forward kernels (dusty_fir2d_adj / dusty_fir2d).
"""
from collections import abc

import torch

from ..... import _cabi as K
from ..... import functional as DF


def _cfg(kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    kh, kw = kernel.shape
    return DF.FirCfg(kh, kw, flip=1, up=(up_y, up_x), down=(down_y, down_x),
                     pad=(pad_y0, pad_y1, pad_x0, pad_x1), mode=(K.PAD_ZERO, K.PAD_ZERO))


class _NativeModule:
    @staticmethod
    def upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
        if not (input.is_cuda and kernel.is_cuda):
            raise RuntimeError("input and kernel must be CUDA tensors")
        if not (input.is_contiguous() and kernel.is_contiguous()):
            raise RuntimeError("input and kernel must be contiguous")
        major, in_h, in_w, minor = input.shape
        if minor != 1:
            raise RuntimeError("only minor == 1 is supported (the reference always reshapes to it)")
        cfg = _cfg(kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        out = DF._fir_raw(input.reshape(major, in_h, in_w), kernel.float(), cfg, False)
        return out.reshape(major, out.shape[1], out.shape[2], 1)


upfirdn2d_op = _NativeModule()


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    if kernel.ndim != 2:
        raise RuntimeError("kernel must be 2-D [kh, kw]")
    if DF.fir1d_supported(input, kernel, up, down, pad):     # single-axis fast path (ADA)
        return DF.fir1d(input, kernel.to(input.device), up, down, pad)
    cfg = _cfg(kernel, up[0], up[1], down[0], down[1], *pad)
    taps = kernel.detach().to(device=input.device, dtype=torch.float32).contiguous()
    return DF.fir2d(input, taps, cfg)
