#!/usr/bin/env python
"""Cost of the data-parallel exchange steps of the graphed trainer, in isolation (torchrun, NCCL):
flat-bucket pack + all-reduce of the D and G gradients, the rank-0 buffer broadcast, the packed
scalar reduction.  CUDA events on rank 0, median of 20, every rank synchronised before each.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/comm_bench.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import dusty_gan_v2_b200 as pkg  # noqa: E402
from dusty_gan_v2_b200.gans.trainer import Trainer  # noqa: E402
from dusty_gan_v2_b200.presets import preset  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg.set_precision("bf16")
    cfg = preset("dusty_v2", batch_size=64 * world)
    tr = Trainer(cfg, bench.cycle(bench.synthetic_batches(2, 64, seed=2 + rank, device=dev)), device=dev, rank=rank,
                 world_size=world, angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"), cuda_graphs=True)
    for p in tr._D_params + tr._G_params:
        p.grad = torch.randn_like(p)

    def timeit(fn, n=20):
        ts = []
        for _ in range(n + 3):
            dist.barrier()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        return sorted(ts[3:])[n // 2]

    flatD = torch.empty(sum(p.numel() for p in tr._D_params), device=dev)
    out = {
        "world": world,
        "D_grads_MB": flatD.numel() * 4 / 1e6,
        "D_pack_allreduce_us": timeit(lambda: tr._reduced_grads("D", tr._D_params)),
        "G_pack_allreduce_us": timeit(lambda: tr._reduced_grads("G", tr._G_params)),
        "D_allreduce_only_us": timeit(lambda: dist.all_reduce(flatD)),
        "G_buffer_broadcast_us": timeit(tr._sync_G_buffers),
        "scalar_allreduce_us": timeit(lambda: dist.all_reduce(torch.zeros(8, device=dev))),
    }
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
