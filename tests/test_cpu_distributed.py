"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: the flat gradient all-reduce of
the graphed trainer path, the rank-0 buffer broadcast (DDP's broadcast_buffers semantics), the
packed per-step scalar reduction, and the reference arm's rank gating in bench.py."""
import os
import subprocess
import sys
import tempfile
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, out_dir):
    from dusty_gan_v2_b200.gans.trainer import Trainer
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # different weights / grads per rank
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
        for p in net.parameters():
            p.grad = torch.full_like(p, float(rank + 1)) + torch.arange(p.numel()).reshape(p.shape) * 0.5
        net[1].running_mean.fill_(float(10 + rank))
        params = list(net.parameters())
        ns = types.SimpleNamespace(world_size=world, cuda_graphs=True, G=net, G_module=net, _G_params=params)
        Trainer._allreduce_grads(ns, params)               # D-style: flat all-reduce (avg)
        for p in params:
            want = torch.full_like(p, 1.5) + torch.arange(p.numel()).reshape(p.shape) * 0.5
            assert torch.allclose(p.grad, want), (rank, p.grad, want)
        for p in params:                                    # G-style helper, second round
            p.grad = torch.full_like(p, float(2 * rank))
        Trainer._allreduce_G_grads(ns)
        assert all(torch.allclose(p.grad, torch.full_like(p, 1.0)) for p in params)
        Trainer._sync_G_buffers(ns)                         # rank 0's buffers win
        assert torch.allclose(net[1].running_mean, torch.full((5,), 10.0)), net[1].running_mean
        packed = torch.tensor([1.0 + rank, 4.0 * (rank + 1)])   # per-step scalars: one all_reduce + avg
        dist.all_reduce(packed)
        packed /= world
        assert torch.allclose(packed, torch.tensor([1.5, 6.0]))
        # no-op guards: a single-process trainer must not touch the process group
        solo = types.SimpleNamespace(world_size=1, cuda_graphs=True, G=net, G_module=net, _G_params=params)
        before = [p.grad.clone() for p in params]
        Trainer._allreduce_grads(solo, params)
        Trainer._allreduce_G_grads(solo)
        assert all(torch.equal(a, p.grad) for a, p in zip(before, params))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_and_buffer_broadcast_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        mp.spawn(_worker, args=(2, init_file, d), nprocs=2, join=True)
        assert os.path.exists(os.path.join(d, "ok0")) and os.path.exists(os.path.join(d, "ok1"))


def test_reference_arm_runs_on_rank0_only():
    """Under torchrun (N > 1) only rank 0 times the CPU reference; the other ranks exit 0 at
    once and print nothing."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
               MASTER_PORT="29533")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-400:]
    assert r.stdout.strip() == ""
