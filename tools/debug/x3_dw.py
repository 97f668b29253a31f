"""Which elements of the split-bf16 dW differ from fp64 (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import dusty_gan_v2_b200.functional as DF
K = DF.K
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(22)
for (B, Oc, C1, H, W) in [(2, 64, 64, 32, 256), (2, 64, 64, 16, 256), (2, 64, 64, 32, 128), (1, 32, 64, 64, 512)]:
    P = H * W
    gy = torch.randn(B, Oc, H, W, generator=g)
    x1 = torch.randn(B, C1, H, W, generator=g)
    ref = torch.bmm(gy.double().reshape(B, Oc, P), x1.double().reshape(B, C1, P).transpose(1, 2)).float()
    gyd, x1d = gy.to(dev), x1.to(dev)
    for trial in range(2):
        gp = DF._split_pixels(gyd, 0)
        xp = DF._split_pixels(x1d, 1)
        gw = torch.empty(B, Oc, C1, device=dev)
        K.call("dusty_modconv_bwd_dw", K.ptr(gp), K.ptr(xp), None, K.ptr(gw), B, Oc, C1, 0, 1, 3 * P, K.BF16, 2,
               0, K.stream_of(gw))
        torch.cuda.synchronize()
        d = (gw.cpu() - ref).abs()
        bad = (d > 2e-5 * ref.abs().max() + 1e-4 * ref.abs()).nonzero()
        print((B, Oc, C1, H, W), "trial", trial, "max err", float(d.max()), "bad", len(bad), bad[:12].tolist())
    # reference through torch on the split tensors themselves (is the split right?)
    gpf = gp.float().reshape(B, Oc, 3 * P)
    xpf = xp.float().reshape(B, C1, 3 * P)
    ref2 = torch.bmm(gpf.double(), xpf.double().transpose(1, 2)).float().cpu()
    print("   split-operand fp64 product vs fp64 ref: max err", float((ref2 - ref).abs().max()),
          " kernel vs split-operand product:", float((gw.cpu() - ref2).abs().max()))
