"""Per-layer timing of the discriminator's dense convolutions: own tcgen05 implicit GEMM
(conv_tc.cu) vs the library (cuDNN, benchmark mode), CUDA events, L2 flushed between launches.
Usage: python tools/conv_bench.py [--batch 64] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dusty_gan_v2_b200.functional as DF  # noqa: E402


def timeit(fn, flush, iters=8):
    """Device time of one call: L2 flushed before EVERY timed launch, and a spin kernel between
    the flush and the start event so that the host has finished enqueueing `fn` (tensor-map
    encoding, Python) before the device reaches it -- otherwise the event interval of a 20 us
    kernel measures the host wrapper."""
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        torch.cuda._sleep(1_000_000)           # ~0.5 ms of device time
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    dev = "cuda"
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev).float()[: (64 << 20)]
    layers = []
    for i in range(4):
        C = 32 << i
        H, W = 64 >> i, 512 >> i
        layers.append((f"RB{i}.conv1 3x3 s1 {C}->{C} @{H}x{W}", C, C, H + 2, W + 2, 3, 1))
        layers.append((f"RB{i}.conv2 3x3 s2 {C}->{2*C} @{H}x{W}", C, 2 * C, H + 2, W + 2, 3, 2))
        layers.append((f"RB{i}.skip 1x1 s2 {C}->{2*C} @{H}x{W}", C, 2 * C, H, W, 1, 2))
    rows = []
    cl = torch.channels_last
    for name, C, O, H, W, k, s in layers:
        x = torch.randn(B, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=cl)
        w = (torch.randn(O, C, k, k, device=dev) / (C * k * k) ** 0.5).to(torch.bfloat16)
        wcl = w.contiguous(memory_format=cl)
        st = (s, s)
        y = torch.nn.functional.conv2d(x, wcl, None, st)
        gy = torch.randn_like(y).contiguous(memory_format=cl)
        flops = 2.0 * y.numel() * C * k * k
        t = {}
        t["fprop_own"] = timeit(lambda: DF.conv2d_fprop_tc(x, w, st), flush)
        t["fprop_lib"] = timeit(lambda: torch.nn.functional.conv2d(x, wcl, None, st), flush)
        t["dgrad_own"] = timeit(lambda: DF.conv2d_dgrad_tc(gy, w, st, (H, W)), flush)
        t["dgrad_lib"] = timeit(lambda: torch.ops.aten.convolution_backward(
            gy, x, wcl, None, st, (0, 0), (1, 1), False, (0, 0), 1, (True, False, False)), flush)
        t["wgrad_own"] = timeit(lambda: DF.conv2d_wgrad_tc(gy, x, st, w.shape, torch.float32), flush)
        t["wgrad_lib"] = timeit(lambda: torch.ops.aten.convolution_backward(
            gy, x, wcl, None, st, (0, 0), (1, 1), False, (0, 0), 1, (False, True, False)), flush)
        byts = 2.0 * (x.numel() + y.numel())
        row = {"layer": name, "gflop": flops / 1e9, "mbytes": byts / 1e6}
        row.update({k2: round(v * 1e3, 1) for k2, v in t.items()})
        rows.append(row)
        print(f"{name:38s} {flops/1e9:7.1f} GF {byts/1e6:7.1f} MB | fprop {t['fprop_own']*1e3:7.1f} / {t['fprop_lib']*1e3:7.1f}"
              f" | dgrad {t['dgrad_own']*1e3:7.1f} / {t['dgrad_lib']*1e3:7.1f}"
              f" | wgrad {t['wgrad_own']*1e3:7.1f} / {t['wgrad_lib']*1e3:7.1f}  us (own / cuDNN)", flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"batch": B, "unit": "us", "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
