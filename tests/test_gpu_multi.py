"""Multi-GPU (`gpurun --gpus 2`) and optimiser tests: the flat NCCL gradient bucket + update of the graphed trainer against a single-GPU double batch, and the multi-tensor Adam / EMA
kernel against torch.optim.Adam."""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_multi_tensor_adam_matches_torch_adam_and_ema_lerp():
    """csrc/optim.cu against torch.optim.Adam (same state layout) over several steps, odd sizes,
    a gradient scale, and the EMA lerp of the shadow weights folded into the pass."""
    from dusty_gan_v2_b200.gans.optim import Adam, multi_copy
    g = torch.Generator().manual_seed(5)
    shapes = [(7,), (33, 5), (4, 3, 3, 3), (1,), (257, 129), (4096 * 3 + 5,)]
    ours = [torch.randn(s, generator=g).to(DEV).requires_grad_() for s in shapes]
    ref = [p.detach().clone().requires_grad_() for p in ours]
    ema = [p.detach().clone() for p in ours]
    ema_ref = [p.detach().clone() for p in ours]
    lazy = 16 / 17.0
    oa = Adam(ours, lr=0.002 * lazy, betas=(0.0, 0.99 ** lazy), ema_params=ema)
    ob = torch.optim.Adam(ref, lr=0.002 * lazy, betas=(0.0, 0.99 ** lazy))
    for step in range(4):
        scale = 0.5 if step % 2 else 1.0
        w = 0.25 if step > 0 else 1.0
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).to(DEV)
            a.grad = gr.clone()
            b.grad = gr * scale
        oa.step(grad_scale=scale, ema_weight=w)
        ob.step()
        torch._foreach_lerp_(ema_ref, [p.detach() for p in ref], w)
        for a, b, e, er in zip(ours, ref, ema, ema_ref):
            np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-6, atol=2e-7)
            np.testing.assert_allclose(e.cpu().numpy(), er.cpu().numpy(), rtol=2e-6, atol=2e-7)
    sa, sb = oa.state_dict(), ob.state_dict()
    assert set(sa["state"][0]) == set(sb["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sa["state"][0]["step"]) == float(sb["state"][0]["step"]) == 4.0
    ob.load_state_dict(sa)                                   # interchangeable checkpoints
    oa.load_state_dict(sb)
    # beta1 != 0 and a fresh optimiser (bias corrections)
    p1 = torch.randn(1000, generator=g).to(DEV).requires_grad_()
    p2 = p1.detach().clone().requires_grad_()
    o1, o2 = Adam([p1], lr=0.01, betas=(0.9, 0.999)), torch.optim.Adam([p2], lr=0.01, betas=(0.9, 0.999))
    for _ in range(3):
        gr = torch.randn(1000, generator=g).to(DEV)
        p1.grad, p2.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    np.testing.assert_allclose(p1.detach().cpu().numpy(), p2.detach().cpu().numpy(), rtol=2e-6, atol=2e-7)
    # multi-tensor copy / scale
    dst = [torch.empty_like(p) for p in ours]
    multi_copy(dst, [p.detach() for p in ours], 0.5)
    for d, p in zip(dst, ours):
        assert torch.equal(d, p.detach() * 0.5)
    with pytest.raises(RuntimeError):
        multi_copy([torch.zeros(3)], [torch.zeros(3)])


def _rank_main(rank, world, init_file, out_file):
    import torch.distributed as dist
    import dusty_gan_v2_b200 as pkg
    from dusty_gan_v2_b200.config import to_attr
    from dusty_gan_v2_b200.gans.trainer import Trainer, set_requires_grad
    from dusty_gan_v2_b200.presets import preset
    from small_cfgs import D_MID, G_MID
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    try:
        pkg.set_precision("fp32")
        torch.manual_seed(1000 + rank)                         # every rank seeds differently ...
        np.random.seed(1000 + rank)
        B = 4
        cfg = preset("dusty_v2", batch_size=B * world, resolution=(32, 128))
        cfg.model.generator, cfg.model.discriminator = to_attr(G_MID), to_attr(D_MID)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        tr = Trainer(cfg, iter([]), device=dev, rank=rank, world_size=world,
                     angle_file=os.path.join(root, "data/coords/kitti_raw.npy"), cuda_graphs=True)
        # ... and must still start from rank 0's weights (DDP-constructor semantics)
        flat = torch.cat([p.detach().reshape(-1) for p in list(tr.G_module.parameters()) + list(tr.D_module.parameters())])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1]), "ranks start from different weights"
        for a, b in zip(tr.G_ema.parameters(), tr.G_module.parameters()):
            assert torch.equal(a, b)
        # one discriminator update on this rank's half of a fixed batch
        gen = torch.Generator().manual_seed(7)
        x_all = torch.tanh(torch.randn(B * world, 1, 32, 128, generator=gen))
        w0 = [p.detach().clone() for p in tr.D_module.parameters()]
        set_requires_grad(tr._D_params, True)
        y = tr.D(x_all[rank * B:(rank + 1) * B].to(dev))
        torch.nn.functional.softplus(-y).mean().backward()
        tr._update_D()                                          # pack -> NCCL -> Adam on the bucket
        torch.cuda.synchronize()
        if rank == 0:
            # the same step on ONE GPU with the double batch and torch's Adam
            import copy
            D1 = copy.deepcopy(tr.D_module)
            with torch.no_grad():
                for p, w in zip(D1.parameters(), w0):
                    p.copy_(w)
            lazy = 16 / 17.0
            opt = torch.optim.Adam(D1.parameters(), lr=0.002 * lazy, betas=(0.0, 0.99 ** lazy))
            # MinibatchStdDev groups the STRIDED sets {m, m + B/4, ..} (reference common.py:237-250):
            # interleave the ranks' samples so that the double batch forms the same groups
            x_single = x_all.view(world, B, 1, 32, 128).transpose(0, 1).reshape(world * B, 1, 32, 128)
            torch.nn.functional.softplus(-D1(x_single.to(dev))).mean().backward()
            g_single = [p.grad.detach().clone() for p in D1.parameters()]
            opt.step()
            views = tr._flat["D"][1]
            worst = 0.0
            for v, gs in zip(views, g_single):                 # reduced bucket / world == double-batch gradient
                d = float((v / world - gs).abs().max() / gs.abs().max().clamp_min(1e-12))
                worst = max(worst, d)
            near = total = 0
            for p, q in zip(tr.D_module.parameters(), D1.parameters()):
                dlt = (p.detach() - q.detach()).abs()
                near += int((dlt < 2e-4).sum())
                total += dlt.numel()
            open(out_file, "w").write(f"{worst} {near / total}")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_flat_bucket_update_equals_single_gpu_double_batch():
    """gradients (and the Adam update) of a 2-rank data-parallel
    discriminator step through the trainer's own exchange path -- rank-0 weight broadcast, one
    packed fp32 bucket, ONE NCCL all-reduce, optimiser reading the bucket with 1 / world_size --
    equal those of one GPU fed the double batch."""
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "result")
        mp.spawn(_rank_main, args=(2, os.path.join(d, "rdzv"), out), nprocs=2, join=True)
        worst, frac = (float(v) for v in open(out).read().split())
    assert worst < 2e-3, worst              # fp32 gradients, different summation orders
    assert frac > 0.98, frac                # Adam-normalised first step (sign flips at ~0 gradients)
