#!/usr/bin/env python
"""Per-iteration device time of the training step (CUDA events between iterations) and the
host time of the same iterations: shows what the R1 iterations (every 16th) cost and whether
an iteration is launch-bound (host time ~ device time).

    python tools/step_times.py [--steps 34] [--batch 64]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import dusty_gan_v2_b200 as pkg  # noqa: E402
from dusty_gan_v2_b200.gans.trainer import Trainer  # noqa: E402
from dusty_gan_v2_b200.presets import preset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=34)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    # under torchrun: one process per GPU, per-GPU batch `--batch` (rank 0 reports)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    pkg.set_precision("bf16")
    torch.backends.cudnn.benchmark = True
    cfg = preset("dusty_v2", batch_size=args.batch * world)
    pool = bench.synthetic_batches(4, args.batch, seed=2 + rank, device=dev)
    tr = Trainer(cfg, bench.cycle(pool), device=dev, rank=rank, world_size=world,
                 angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))
    for i in range(4):
        tr.step(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    host = []
    ev[0].record()
    for i in range(args.steps):
        t0 = time.perf_counter()
        tr.step(16 + i)
        host.append((time.perf_counter() - t0) * 1e3)
        ev[i + 1].record()
    torch.cuda.synchronize()
    devms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    rows = [{"iteration": 16 + i, "r1": (16 + i) % 16 == 0, "device_ms": round(devms[i], 2),
             "host_ms": round(host[i], 2)} for i in range(args.steps)]
    plain = [r["device_ms"] for r in rows if not r["r1"]]
    r1 = [r["device_ms"] for r in rows if r["r1"]]
    out = {"plain_ms_median": sorted(plain)[len(plain) // 2], "r1_ms": r1,
           "host_plain_ms_median": sorted(r["host_ms"] for r in rows if not r["r1"])[len(plain) // 2],
           "host_r1_ms": [r["host_ms"] for r in rows if r["r1"]], "rows": rows}
    if rank == 0:
        print(json.dumps({k: v for k, v in out.items() if k != "rows"}))
        print("device ms per iteration:", [r["device_ms"] for r in rows])
    if args.json and rank == 0:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
