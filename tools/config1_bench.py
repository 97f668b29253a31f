#!/usr/bin/env python
"""BASELINE config 1: dusty_v2 generator forward (eval, truncation psi = 1), batch 8, 64x512, random-init
weights -- the configuration the reference can run on CPU.  Times the device forward (CUDA events, after
warm-up, inputs resident; and end to end from pinned host z with the five output maps copied back), and,
with --cpu, the oracle port of the same forward on the host cores through bench.py (the one measurement
entry that executes oracle/).

    python tools/config1_bench.py --cpu --json gpurun_out/config1.json

(Written at the end of round 1 after the GPU budget was spent: not yet run on a B200.)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import dusty_gan_v2_b200 as pkg  # noqa: E402
from dusty_gan_v2_b200.gans.coords import CoordBridge  # noqa: E402
from dusty_gan_v2_b200.gans.models.builder import build_generator  # noqa: E402
from dusty_gan_v2_b200.presets import preset  # noqa: E402

KEYS = ("image", "image_orig", "raydrop_logit", "raydrop_mask")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "config1.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pkg.set_precision(args.precision)
    torch.manual_seed(0)
    G = build_generator(preset("dusty_v2").model.generator).eval().requires_grad_(False)
    sd = {k: v.clone() for k, v in G.state_dict().items()} if args.cpu else None
    G = G.to(dev)
    coord = CoordBridge(64, 512, 1.45, 80.0, os.path.join(ROOT, "data/coords/kitti_raw.npy")).to(dev)
    B = args.batch
    angle = coord.angle.expand(B, -1, -1, -1)
    z_host = torch.randn(B, 512, generator=torch.Generator().manual_seed(1)).pin_memory()
    z_dev = z_host.to(dev)
    outs_host = {k: torch.empty(B, 1, 64, 512).pin_memory() for k in KEYS}

    def device_step():
        with torch.no_grad():
            return G(z_dev, angle=angle)

    def e2e_step():
        with torch.no_grad():
            out = G(z_host.to(dev, non_blocking=True), angle=angle)
        for k in KEYS:
            outs_host[k].copy_(out[k].float(), non_blocking=True)

    res = {"workload": f"dusty_v2 generator forward (eval), batch {B}, 64x512, {args.precision}",
           "steps": args.steps, "warmup": args.warmup}
    for name, fn in (("device", device_step), ("e2e", e2e_step)):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[name] = {"ms_per_forward": ms, "images_per_s": B / (ms / 1e3)}
    res["e2e"].update(h2d_bytes_per_step=z_host.numel() * 4, d2h_bytes_per_step=len(KEYS) * B * 64 * 512 * 4)
    if args.cpu:
        res["cpu_oracle"] = bench.generator_forward_cpu_baseline(sd, z_host.clone(), coord.angle.cpu(), B)
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    json.dump(res, open(args.json, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
