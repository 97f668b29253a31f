"""How fast is the oracle PORT next to the reference it restates?  bench.py's `cpu_baseline` /
`--impl reference` legs time the port (the Python reference cannot travel to the GPU box), so the
port must not be a strawman: this script times both on the same host cores in the build container --
full-size dusty_v2 generator forward (BASELINE config 1, batch 8) and discriminator forward +
backward with and without the R1 double backward (batch 4).

    python tests/golden/calibrate_cpu_port.py    ->  prints a table (kept in profiles/)
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import dusty_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

ref_import.install()

from gans.models.builder import build_discriminator, build_generator  # noqa: E402

from dusty_gan_v2_b200.presets import preset  # noqa: E402


def best(fn, reps=3):
    ts = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts[1:])


def full_step_row(batch=2):
    """Full-size dusty_v2 training iteration (no R1: iteration 1) at `batch`: the reference's real
    Trainer.step (assembled for CPU by tests/ref_trainer_harness.py) vs the port exactly as
    bench.py's CPU baseline runs it."""
    import tempfile
    from types import SimpleNamespace

    import torch.distributed as dist

    import bench
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ref_trainer_harness import build_reference_trainer
    cfg = preset("dusty_v2").model
    pool = bench.synthetic_batches(3, batch, seed=2)
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("gloo", init_method=f"file://{tempfile.mkdtemp()}/pg", rank=0, world_size=1)
    try:
        T, _, _ = build_reference_trainer(dict(cfg.generator), dict(cfg.discriminator), batch, (64, 512), pool)
        T.step(1)                                    # warm-up (plain iteration)
        t0 = time.perf_counter()
        T.step(2)
        t_ref = time.perf_counter() - t0
    finally:
        if own:
            dist.destroy_process_group()
    args = SimpleNamespace(arch="dusty_v2", ada_p=None)
    _, spt, _, desc = bench.cpu_reference_run(args, 2, 1, batch)      # iterations 0 (R1) and 1 (plain)
    # the description carries the plain / R1 split: "... ratio (P s / R s)"
    plain = float(desc.split("ratio (")[1].split(" s")[0])
    return (f"training iteration (G + D step), batch {batch}", t_ref, plain)


def main():
    O.set_fir_impl("library")          # what bench.py's CPU legs run
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = preset("dusty_v2").model
    G = build_generator(ref_import.to_attr(dict(cfg.generator))).eval()
    D = build_discriminator(ref_import.to_attr(dict(cfg.discriminator)))
    angle = O.angle_grid(np.load(os.path.join(ROOT, "data/coords/kitti_raw.npy")), 64, 512).repeat_interleave(8, 0)
    z, u = torch.randn(8, 512), torch.rand(8, 1, 64, 512)
    sdG = {k: v.clone() for k, v in G.state_dict().items()}
    rows = []
    with torch.no_grad():
        rows.append(("G forward, eval, batch 8", best(lambda: G(z, angle=angle)),
                     best(lambda: O.generator(sdG, z, angle, u))))
    x = torch.tanh(torch.randn(4, 1, 64, 512))
    sdD = {k: v.clone().requires_grad_("kernel" not in k) for k, v in D.state_dict().items()}
    params = [v for v in sdD.values() if v.requires_grad]

    def ref_d(r1):
        xx = x.clone().requires_grad_(r1)
        y = D(xx)
        if r1:
            (g,) = torch.autograd.grad(y.sum(), xx, create_graph=True)
            g.pow(2).sum([1, 2, 3]).mean().backward()
        else:
            torch.nn.functional.softplus(-y).mean().backward()

    def port_d(r1):
        xx = x.clone().requires_grad_(r1)
        y = O.discriminator(sdD, xx)
        if r1:
            (g,) = torch.autograd.grad(y.sum(), xx, create_graph=True)
            torch.autograd.grad(O.r1_penalty(g), params, allow_unused=True)
        else:
            torch.autograd.grad(torch.nn.functional.softplus(-y).mean(), params)

    rows.append(("D forward + backward, batch 4", best(lambda: ref_d(False)), best(lambda: port_d(False))))
    rows.append(("D forward + R1 double backward, batch 4", best(lambda: ref_d(True)), best(lambda: port_d(True))))
    rows.append(full_step_row())
    print(f"# host threads: {torch.get_num_threads()}; seconds, best of 3 after one warm-up "
          f"(training step: one plain iteration after one warm-up iteration)")
    print(f"{'case':42s} {'reference':>10s} {'port':>10s} {'port/ref':>9s}")
    for name, r, p in rows:
        print(f"{name:42s} {r:10.3f} {p:10.3f} {p / r:9.2f}")


if __name__ == "__main__":
    main()
