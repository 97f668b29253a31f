"""One launch of every own discriminator-convolution kernel (fprop / dgrad / wgrad per layer of
the B=64 training step) -- the command ncu profiles:
    ncu --set full --clock-control none --import-source on -k regex:conv_ -o gpurun_out/conv python tools/conv_once.py
Usage: python tools/conv_once.py [--batch 64] [--layers 0,3,6,9]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dusty_gan_v2_b200.functional as DF  # noqa: E402


def layers(B):
    out = []
    for i in range(4):
        C = 32 << i
        H, W = 64 >> i, 512 >> i
        out.append((f"RB{i}.conv1", C, C, H + 2, W + 2, 3, 1))
        out.append((f"RB{i}.conv2", C, 2 * C, H + 2, W + 2, 3, 2))
        out.append((f"RB{i}.skip", C, 2 * C, H // 2, W // 2, 1, 1))     # 1x1 on the decimated blur
    out.append(("EP.conv", 512, 512, 4 + 2, 32 + 2, 3, 1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--layers", default=None)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    dev, cl = "cuda", torch.channels_last
    sel = None if args.layers is None else {int(v) for v in args.layers.split(",")}
    for idx, (name, C, O, H, W, k, s) in enumerate(layers(args.batch)):
        if sel is not None and idx not in sel:
            continue
        x = torch.randn(args.batch, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=cl)
        w = (torch.randn(O, C, k, k, device=dev) / (C * k * k) ** 0.5).to(torch.bfloat16)
        st = (s, s)
        for _ in range(args.reps):
            torch.cuda.nvtx.range_push(name)
            y = DF.conv2d_fprop_tc(x, w, st)
            gy = torch.randn_like(y).contiguous(memory_format=cl)
            DF.conv2d_dgrad_tc(gy, w, st, (H, W))
            DF.conv2d_wgrad_tc(gy, x, st, w.shape, torch.float32)
            torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
        print(name, "ok", flush=True)


if __name__ == "__main__":
    main()
