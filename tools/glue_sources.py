#!/usr/bin/env python
"""Where the library (ATen) glue kernels of a training iteration come from.

Runs a few EAGER iterations (no CUDA graphs, so every launch has a host-side op) under
torch.profiler with shapes and Python stacks, and lists the ATen ops by self device time
together with their input shapes and the innermost frames inside this repository.  Ops issued
by the autograd engine itself (gradient accumulation of a tensor with two consumers) have no
Python frames and are listed as "<autograd engine>".

    python tools/glue_sources.py --steps 2 --out gpurun_out/glue_sources.txt
"""
import argparse
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
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--first-iter", type=int, default=17)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "glue_sources.txt"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pkg.set_precision("bf16")
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0)
    cfg = preset("dusty_v2", batch_size=args.batch)
    pool = bench.synthetic_batches(2, args.batch, seed=2, device=dev)
    tr = Trainer(cfg, bench.cycle(pool), device=dev, cuda_graphs=False,
                 angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))
    tr.A.generator = torch.Generator().manual_seed(100)
    for i in range(3):
        tr.step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True,
                 with_stack=True) as prof:
        for i in range(args.steps):
            tr.step(args.first_iter + i)
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages(group_by_input_shape=True, group_by_stack_n=24):
        t = getattr(ev, "self_device_time_total", 0) or 0
        if t <= 0 or ev.device_type != torch.autograd.DeviceType.CPU:
            continue
        if not ev.key.startswith("aten::"):
            continue
        frames = [s for s in (ev.stack or []) if "dusty" in s or "bench.py" in s or "tools/" in s]
        rows.append((t / 1e3 / args.steps, ev.count / args.steps, ev.key, str(ev.input_shapes)[:110],
                     " <- ".join(f.replace(ROOT + "/", "")[:70] for f in frames[:4])
                     or "<autograd engine>"))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(f"# eager iterations {args.first_iter}..{args.first_iter + args.steps - 1}, batch "
                f"{args.batch}: ATen ops by self device time, per iteration; total {total:.2f} ms\n")
        for ms, n, key, shapes, where in rows[:90]:
            f.write(f"{ms:8.3f} ms {n:6.1f}x  {key:28s} {shapes}\n{'':22s}{where}\n")
    print(open(args.out).read()[:7000])


if __name__ == "__main__":
    main()
