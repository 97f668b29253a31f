#!/usr/bin/env python
"""Kernel-level time breakdown of the training step with torch.profiler (CUPTI, no replay).
Complements the ncu launch list: quick, un-serialised, warm caches.  Writes a table of the
top kernels (by total device time) over `--steps` iterations to gpurun_out/.

    python tools/profile_step.py --steps 4 --out gpurun_out/step_profile.txt
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import dusty_gan_v2_b200 as pkg  # noqa: E402
from dusty_gan_v2_b200.gans.trainer import Trainer  # noqa: E402
from dusty_gan_v2_b200.presets import preset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.txt"))
    ap.add_argument("--first-iter", type=int, default=16)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pkg.set_precision(args.precision)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = args.precision == "bf16"
    torch.manual_seed(0)
    cfg = preset("dusty_v2", batch_size=args.batch)
    pool = bench.synthetic_batches(2, args.batch, seed=2, device=dev)
    tr = Trainer(cfg, bench.cycle(pool), device=dev,
                 angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))
    tr.A.generator = torch.Generator().manual_seed(100)
    for i in range(3):
        tr.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        e0.record()
        for i in range(args.steps):
            tr.step(args.first_iter + i)
        e1.record()
        torch.cuda.synchronize()
    wall_ms = e0.elapsed_time(e1)
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0)
        if ev.device_type == torch.autograd.DeviceType.CUDA and t > 0:
            rows.append((t / 1e3, ev.count, ev.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(f"# {args.steps} steps (first iteration {args.first_iter}), batch {args.batch}, "
                f"{args.precision}; event time {wall_ms:.1f} ms, sum of kernel time {total:.1f} ms\n")
        f.write(f"{'ms_total':>10} {'share':>7} {'calls':>7}  kernel\n")
        for ms, n, name in rows[:70]:
            f.write(f"{ms:10.2f} {100 * ms / total:6.1f}% {n:7d}  {name[:150]}\n")
    json.dump({"event_ms": wall_ms, "kernel_ms": total, "steps": args.steps,
               "top": [{"ms": r[0], "calls": r[1], "name": r[2]} for r in rows[:200]]},
              open(args.out.replace(".txt", ".json"), "w"))
    print(open(args.out).read()[:6000])


if __name__ == "__main__":
    main()
