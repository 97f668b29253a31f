"""Host time of the trainer's phases per iteration (wrap methods with perf_counter); under torchrun."""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench, dusty_gan_v2_b200 as pkg
from dusty_gan_v2_b200.gans.trainer import Trainer
from dusty_gan_v2_b200.presets import preset

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pkg.set_precision("bf16")
cfg = preset("dusty_v2", batch_size=64 * world)
tr = Trainer(cfg, bench.cycle(bench.synthetic_batches(4, 64, seed=2 + rank, device=dev)), device=dev, rank=rank,
             world_size=world, angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))
acc = collections.defaultdict(float)

def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[label or name] += time.perf_counter() - t
    setattr(obj, name, w)

for i in range(4):
    tr.step(i)
torch.cuda.synchronize()
for n in ("_G_train_forward", "_D_forward", "_fake_images_nograd", "_reduced_grads", "_sync_G_buffers", "_update_D",
          "_copy_ema_buffers", "_next_batch", "fetch_reals", "sample_z", "warmup"):
    wrap(tr, n)
wrap(tr.optim_G, "step", "optim_G.step")
wrap(tr.optim_D, "step", "optim_D.step")
wrap(tr.A, "forward", "ADA.forward")
if world > 1:
    wrap(dist, "all_reduce", "dist.all_reduce")
    wrap(dist, "broadcast", "dist.broadcast")
N = 12
t0 = time.perf_counter()
for i in range(N):
    tr.step(17 + i)
host = time.perf_counter() - t0
torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}: host {host / N * 1e3:.2f} ms per plain iteration")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"   {k:24s} {v / N * 1e3:7.2f} ms")
