"""Drop-in for gans/models/ops/fused_act/fused_act.py (reference lines 20-129).

Same public names and call contracts: `fused` (the native module stand-in with
`fused_bias_act`), `FusedLeakyReLUFunction`, `FusedLeakyReLUFunctionBackward`,
`FusedLeakyReLU`, `fused_leaky_relu`.  The CUDA work is dusty_bias_act /
dusty_bias_act_bwd from libdusty_b200.so; CPU tensors raise (no fallback).
"""
import torch
from torch import nn

from ..... import functional as DF


class _NativeModule:
    """Stands in for the JIT-built pybind module `fused` (fused_act.py:10-17)."""

    fused_bias_act = staticmethod(DF.fused_bias_act)


fused = _NativeModule()

# autograd pair, names kept for callers that reference them directly
FusedLeakyReLUFunction = DF._BiasAct
FusedLeakyReLUFunctionBackward = DF._BiasActBackward


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    if bias is not None and input.ndim >= 2 and bias.numel() != input.shape[1]:
        raise RuntimeError("bias must have one element per channel (dim 1)")
    return DF.bias_act(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)

    def extra_repr(self):
        return f"negative_slope={self.negative_slope}, scale={self.scale:.4f}"
