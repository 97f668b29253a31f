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
        # (1) what DDP's constructor does and the graphed trainer must do itself: every rank
        # starts from rank 0's parameters AND buffers (ranks seed differently)
        net[1].running_mean.fill_(float(10 + rank))
        mine = [p.detach().clone() for p in net.parameters()]
        Trainer._broadcast_module_states(net)
        gathered = [torch.zeros_like(torch.cat([p.reshape(-1) for p in net.parameters()])) for _ in range(world)]
        dist.all_gather(gathered, torch.cat([p.detach().reshape(-1) for p in net.parameters()]))
        assert torch.equal(gathered[0], gathered[1]), "parameters differ across ranks after the broadcast"
        if rank == 0:
            assert all(torch.equal(a, p) for a, p in zip(mine, net.parameters()))
        assert torch.allclose(net[1].running_mean, torch.full((5,), 10.0)), net[1].running_mean
        # (2) flat gradient bucket: packed, summed by ONE all-reduce, handed to the optimiser as
        # views + a 1 / world_size scale (the packing kernel is CUDA-only: torch stands in here)
        for p in net.parameters():
            p.grad = torch.full_like(p, float(rank + 1)) + torch.arange(p.numel()).reshape(p.shape) * 0.5
        params = list(net.parameters())
        ns = types.SimpleNamespace(world_size=world, cuda_graphs=True, device=torch.device("cpu"), _flat={},
                                   _pack=lambda dst, src: torch._foreach_copy_(dst, src))
        views, scale = Trainer._reduced_grads(ns, "D", params)
        assert scale == 1.0 / world and len(views) == len(params)
        for p, v in zip(params, views):
            want = torch.full_like(p, 1.5) + torch.arange(p.numel()).reshape(p.shape) * 0.5
            assert v.shape == p.shape and torch.allclose(v * scale, want), (rank, v, want)
        flat0 = ns._flat["D"][0]
        views2, _ = Trainer._reduced_grads(ns, "D", params)          # bucket is reused, not re-allocated
        assert ns._flat["D"][0] is flat0 and views2[0].data_ptr() == views[0].data_ptr()
        # (3) rank 0's generator buffers win before a graphed forward (DDP broadcast_buffers)
        net[1].running_mean.fill_(float(20 + rank))
        nsg = types.SimpleNamespace(world_size=world, G_module=net)
        Trainer._sync_G_buffers(nsg)
        assert torch.allclose(net[1].running_mean, torch.full((5,), 20.0)), net[1].running_mean
        packed = torch.tensor([1.0 + rank, 4.0 * (rank + 1)])   # per-step scalars: one all_reduce + avg
        dist.all_reduce(packed)
        packed /= world
        assert torch.allclose(packed, torch.tensor([1.5, 6.0]))
        # no-op guards: a single-process trainer must not touch the process group
        solo = types.SimpleNamespace(world_size=1, cuda_graphs=True, device=torch.device("cpu"), _flat={})
        assert Trainer._reduced_grads(solo, "D", params) == (None, 1.0)
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
