"""Stability of the graphed trainer with its side streams: N steps, every step's scalars and the
final parameters must be finite (a cross-stream race would show up as garbage / NaN sooner or later)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, dusty_gan_v2_b200 as pkg
from dusty_gan_v2_b200.gans.trainer import Trainer
from dusty_gan_v2_b200.presets import preset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dev = torch.device("cuda", 0)
pkg.set_precision("bf16")
cfg = preset("dusty_v2", batch_size=64)
cfg.training.augment.p_init = 0.5
tr = Trainer(cfg, bench.cycle(bench.synthetic_batches(4, 64, seed=2, device=dev)), device=dev,
             angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))
bad = 0
hist = []
for i in range(n):
    packed = tr.step(i)
    if i % 20 == 0 or i == n - 1:
        v = packed.float().cpu()
        hist.append((i, [round(float(x), 4) for x in v[:4]]))
        if not torch.isfinite(v).all():
            bad += 1
torch.cuda.synchronize()
finite = all(torch.isfinite(p).all().item() for p in list(tr.G.parameters()) + list(tr.D.parameters()))
print("steps", n, "non-finite scalar reads", bad, "parameters finite", finite)
print(hist[:3], "...", hist[-3:])
