#!/usr/bin/env python
"""One launch each of the kernels added late in round 1 (fork backward, up2 + sum of squares, the
register-window FIR passes of ADA) at their training-step shapes -- a small target for
`ncu --set full -k regex:residual_fork|up2_fwd|fir1d` (DRAM bytes per launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import dusty_gan_v2_b200 as pkg  # noqa: E402
import dusty_gan_v2_b200.functional as DF  # noqa: E402
from dusty_gan_v2_b200 import _cabi as K  # noqa: E402
from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d  # noqa: E402

dev = torch.device("cuda", 0)
bf, CL = torch.bfloat16, torch.channels_last
B, H, W = 64, 64, 512
taps = (0.125, 0.375, 0.375, 0.125)
gp = torch.randn(B, 32, H + 2, W + 2, device=dev, dtype=bf).contiguous(memory_format=CL)
gd = torch.randn(B, 32, H // 2, W // 2, device=dev, dtype=bf).contiguous(memory_format=CL)
dx = torch.empty(B, 32, H, W, device=dev, dtype=bf).contiguous(memory_format=CL)
h = torch.randn(B, 64, 32, 256, device=dev, dtype=bf)
k = torch.randn(12, device=dev)
x1 = torch.randn(B, 1, 76, 524, device=dev)
x2 = torch.randn(B, 1, 76, 1048, device=dev)
x3 = torch.randn(B, 1, 140, 512, device=dev)
torch.cuda.synchronize()
K.call("dusty_residual_fork_bwd_cl", K.ptr(gp), K.ptr(gd), K.ptr(dx), *taps, B, H, W, 32, K.BF16,
       K.stream_of(dx))
DF.up2_with_sumsq(h, tuple(2 * t for t in taps))
upfirdn2d(x1, k.reshape(1, 12).contiguous(), up=(2, 1), pad=(6, 5, 0, 0))
upfirdn2d(x2, k.reshape(12, 1).contiguous(), up=(1, 2), pad=(0, 0, 6, 5))
upfirdn2d(x3, k.reshape(12, 1).contiguous(), down=(1, 2), pad=(0, 0, -1, -1))
torch.cuda.synchronize()
print("launched", pkg.launch_count())
