#!/usr/bin/env python
"""BASELINE config 5: GAN-inversion latent optimisation (demo_inversion.py stage 1) over 256
synthetic range images with the full-size dusty_v2 generator -- `LatentInversion.step_1st` per
chunk of `--batch` targets: G eval forward -> two multi-scale masked losses -> backward to the
latent -> Adam step.  Reports target-iterations / s (CUDA events, after warm-up), the final
integer counts of the projected clouds, and -- with --cpu -- the oracle's restatement of the
same step on the host cores for a bounded sample.

    python tools/inversion_bench.py --targets 256 --batch 64 --steps 10 --json gpurun_out/inv.json
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
from dusty_gan_v2_b200.gans.inversion import LatentInversion  # noqa: E402
from dusty_gan_v2_b200.gans.models.builder import build_generator  # noqa: E402
from dusty_gan_v2_b200.presets import preset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--latent-type", default="w+", choices=["w", "w+"])
    ap.add_argument("--schedule-steps", type=int, default=500,
                    help="length of the learning-rate schedule (demo default); the bench runs its first steps")
    ap.add_argument("--cpu", action="store_true", help="also time the oracle on the host (batch 2)")
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "inversion_bench.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pkg.set_precision(args.precision)
    torch.backends.cuda.matmul.allow_tf32 = args.precision == "bf16"
    torch.manual_seed(0)
    cfg = preset("dusty_v2", batch_size=args.batch)
    G = build_generator(cfg.model.generator).to(dev).eval()
    angle_file = os.path.join(ROOT, "data/coords/kitti_raw.npy")
    coord = CoordBridge(64, 512, 1.45, 80.0, angle_file).to(dev)
    n_chunks = (args.targets + args.batch - 1) // args.batch
    pool = bench.synthetic_batches(n_chunks, args.batch, seed=2, device=dev)
    total = args.steps + args.warmup
    jobs = [LatentInversion(G, coord, b["depth"] * b["mask"], b["mask"], latent_type=args.latent_type,
                            num_steps_1st=max(total, args.schedule_steps), num_steps_2nd=0,
                            num_z_samples=10_000)
            for b in pool]
    losses = []
    for s in range(args.warmup):
        for j in jobs:
            j.step_1st(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first = torch.stack([j.forward()[1].detach().mean() for j in jobs]).mean().item()
    e0.record()
    for s in range(args.warmup, total):
        for j in jobs:
            out, loss = j.step_1st(s)
        losses.append(loss.mean())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    last = torch.stack([j.forward()[1].detach().mean() for j in jobs]).mean().item()
    pts = jobs[-1].point_cloud(out, "inv_depth")
    res = {
        "workload": f"GAN inversion stage 1, {args.targets} synthetic 64x512 targets in chunks of "
                    f"{args.batch}, latent {args.latent_type}, dusty_v2 full size, {args.precision}",
        "ms_per_iteration_all_targets": ms,
        "target_iterations_per_s": n_chunks * args.batch / (ms / 1e3),
        "steps": args.steps, "warmup": args.warmup,
        "mean_loss_before_timed": first, "mean_loss_after_timed": last,
        "last_chunk_valid_points": int(jobs[-1].last_valid_count),
        "last_chunk_raydrop_kept": int(out["raydrop_mask"].sum().item()),
        "last_chunk_point_set_shape": list(pts.shape),
    }
    if args.cpu:
        sd = {k: v.detach().float().cpu() for k, v in G.state_dict().items()}
        b = {k: v[:2].cpu() for k, v in pool[0].items()}
        res["cpu_oracle"] = bench.inversion_cpu_baseline(sd, jobs[0].z.detach()[:2].float().cpu(),
                                                         coord.angle.cpu(), b["depth"] * b["mask"], b["mask"],
                                                         args.latent_type)
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    json.dump(res, open(args.json, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
