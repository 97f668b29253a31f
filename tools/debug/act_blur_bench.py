"""blur adjoint + bias_act backward: separate kernels vs the fused kernel (isolated, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import dusty_gan_v2_b200.functional as DF
K = DF.K
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=9):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        torch.cuda._sleep(300_000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts[2:])[n // 2]

taps = (0.125, 0.375, 0.375, 0.125)
for (B, C, H, W) in [(64, 32, 64, 512), (64, 64, 32, 256), (128, 32, 64, 512)]:
    cl = torch.channels_last
    gp = torch.randn(B, C, H + 2, W + 2, device=dev).bfloat16().contiguous(memory_format=cl)
    y = torch.randn(B, C, H, W, device=dev).bfloat16().contiguous(memory_format=cl)
    a = timeit(lambda: DF._BlurPadCL.apply(gp, taps, True))
    g = DF._BlurPadCL.apply(gp, taps, True)
    b = timeit(lambda: DF._BiasActBackward.apply(g, y, True, 0.2, 1.414))
    f = timeit(lambda: DF.blur_pad_adj_act(gp, y, taps, 0.2, 1.414))
    print((B, C, H, W), f"blur_adj {a:.1f} us + bias_act_bwd {b:.1f} us = {a + b:.1f}  | fused {f:.1f} us")
