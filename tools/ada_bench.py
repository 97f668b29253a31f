#!/usr/bin/env python
"""AdaptiveAugment at the training step's shapes: the fused device-side op (csrc/ada_fused.cu)
against the stage-by-stage composite path of the same module, forward and forward + backward,
CUDA events, median of 20."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dusty_gan_v2_b200.gans.augment.adaptive_augment import AdaptiveAugment  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda", 0)
    out = {}
    pol = dict(lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1, frac_trans=1, brightness=1, contrast=1,
               luma_flip=1, hue=1, saturation=1)
    for B in (64, 128):
        ada = AdaptiveAugment(p_init=0.6, **pol).to(dev)
        x = torch.randn(B, 1, 64, 512, device=dev, requires_grad=True)
        gy = torch.randn(B, 1, 64, 512, device=dev)
        row = {}
        for fused in (True, False):
            ada.fused = fused
            ada.generator = None if fused else torch.Generator().manual_seed(0)
            tag = "fused" if fused else "composite"
            with torch.no_grad():
                row[f"{tag}_fwd_us"] = timeit(lambda: ada(x))
            row[f"{tag}_fwd_bwd_us"] = timeit(lambda: torch.autograd.grad(ada(x), x, gy))
        out[f"B{B}"] = row
        print(B, {k: round(v, 1) for k, v in row.items()}, flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
