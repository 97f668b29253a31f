#!/usr/bin/env python
"""Host-side (Python) profile of the training step: where does the launch-side time go?"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import dusty_gan_v2_b200 as pkg  # noqa: E402
from dusty_gan_v2_b200.gans.trainer import Trainer  # noqa: E402
from dusty_gan_v2_b200.presets import preset  # noqa: E402

dev = torch.device("cuda", 0)
pkg.set_precision("bf16")
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
cfg = preset("dusty_v2", batch_size=64)
pool = bench.synthetic_batches(2, 64, seed=2, device=dev)
tr = Trainer(cfg, bench.cycle(pool), device=dev, angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))
tr.A.generator = torch.Generator().manual_seed(100)
for i in range(3):
    tr.step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(8):
    tr.step(17 + i)
t_launch = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"8 steps (no R1): host launch time {t_launch*1e3/8:.1f} ms/step, wall {t_all*1e3/8:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(8):
    tr.step(33 + i)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
print(s.getvalue()[:7000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats("dusty_gan_v2_b200|optim|run_backward", 45)
print(s.getvalue()[:9000])
