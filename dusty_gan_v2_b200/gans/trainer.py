"""Host-side training step: mirror of gans/trainer.py `Trainer.step` (reference 247-482)
with the same phase order, loss definitions, lazy regularisation and EMA schedule --

    G step  : z -> G -> warmup -> ADA -> D -> softplus(-y) -> backward -> Adam(G)
    D step  : G (no grad), real/fake -> warmup -> ADA -> D -> nsgan -> backward -> Adam(D)
    R1      : every lazy.gp iterations, (gp*lazy/2) * mean ||grad_x D(ADA(warmup(x)))||^2
    exit    : EMA of G, ADA p update every lazy.ada iterations, scalar reduction

-- but without the host syncs of the reference: per-step scalars stay on the device in one
packed tensor (one all_reduce instead of 6-9, no .item()), ADA transforms are sampled on
the host, and the dataset (KITTI, out of scope) is replaced by any iterator of
{"depth", "mask"} batches.  Multi-GPU: torch DDP (NCCL) on G and D gradient buckets only,
exactly the collectives of the reference.
"""
import copy
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

from .. import functional as DF
from .augment.adaptive_augment import AdaptiveAugment
from .coords import CoordBridge
from .models.builder import build_discriminator, build_generator
from .models.loss import GANLoss
from .models.ops.common import filter2d
from .optim import Adam, multi_copy


def set_requires_grad(net, requires_grad: bool = True):
    params = net if isinstance(net, (list, tuple)) else net.parameters()
    for p in params:
        p.requires_grad = requires_grad


def sigmoid_to_tanh(x):
    return x * 2.0 - 1.0


def tanh_to_sigmoid(x):
    return (x + 1.0) / 2.0


_EMA_LISTS = {}


@torch.no_grad()
def ema_inplace(ema_model, new_model, decay):
    """reference trainer.py:30-41, as multi-tensor launches instead of ~240 tiny ones:
    params lerp towards the new weights, buffers are copied."""
    key = (id(ema_model), id(new_model))
    lists = _EMA_LISTS.get(key)
    if lists is None:
        lists = ([p.data for p in ema_model.parameters()], [p.data for p in new_model.parameters()],
                 list(ema_model.buffers()), list(new_model.buffers()))
        _EMA_LISTS[key] = lists
    ema_p, new_p, ema_b, new_b = lists
    torch._foreach_lerp_(ema_p, new_p, 1 - decay)
    torch._foreach_copy_(ema_b, new_b)


class _GImage(torch.nn.Module):
    """z -> G(z, angle)["image"]: the tensor-in / tensor-out view of the generator that CUDA
    graph capture needs (the dict of outputs stays internal)."""

    def __init__(self, G, auxin):
        super().__init__()
        self.G = G
        self._auxin = auxin

    def forward(self, z):
        return self.G(z, **self._auxin)["image"]


class _DLogits(torch.nn.Module):
    """x -> D(x): separate wrapper instances are graph-captured for the two ways the
    discriminator is used (frozen with a differentiable input in the G step; trainable with
    a detached, real+fake stacked input in the D step)."""

    def __init__(self, D, sub_batches=1):
        super().__init__()
        self.D = D
        self._sub = sub_batches
        self._mb = [m for m in D.modules() if hasattr(m, "sub_batches")]

    def forward(self, x):
        for m in self._mb:
            m.sub_batches = self._sub
        try:
            return self.D(x)
        finally:
            for m in self._mb:
                m.sub_batches = 1


class Trainer:
    def __init__(self, cfg, batch_iter, device=None, rank=0, world_size=1,
                 angle_file="data/coords/kitti_raw.npy", precision=None, cuda_graphs=True):
        self.cfg = cfg
        self.rank, self.world_size = rank, world_size
        self.device = torch.device(device if device is not None else f"cuda:{rank}")
        if precision is not None:
            DF.set_precision(precision)
        tr = cfg.training
        self.resolution = cfg.model.generator.synthesis_kwargs.resolution
        self.B = tr.batch_size // world_size
        self.batch_iter = batch_iter
        self._prefetched = None          # (iterator, device batch, event): next step's H2D copy in flight
        self._copy_stream = None

        self.G = build_generator(cfg.model.generator).to(self.device)
        self.G_ema = copy.deepcopy(self.G).eval()
        self.D = build_discriminator(cfg.model.discriminator).to(self.device)
        self.A = AdaptiveAugment(p_init=tr.augment.p_init, p_target=tr.augment.p_target,
                                 kimg=tr.augment.kimg, **tr.augment.policy).to(self.device)
        self.coord = CoordBridge(num_ring=self.resolution[0], num_points=self.resolution[1],
                                 min_depth=cfg.dataset.min_depth, max_depth=cfg.dataset.max_depth,
                                 angle_file=angle_file).eval().to(self.device)
        self.G_module, self.D_module = self.G, self.D
        self.cuda_graphs = bool(cuda_graphs) and self.device.type == "cuda"
        if world_size > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            kw = dict(device_ids=[self.device.index])
            if not self.cuda_graphs:
                self.G = DDP(self.G, broadcast_buffers=True, **kw)
                self.D = DDP(self.D, broadcast_buffers=False, **kw)
            else:
                # graphed path: no DDP wrapper (graph replays bypass its hooks), so do what its
                # constructor does -- every rank starts from rank 0's parameters and buffers
                # (ranks seed differently: reference utils.init_random_seed(seed, rank)) -- and
                # re-derive G_ema from the synchronised G.  Gradients of G (17.5 MB) and D
                # (154 MB) are all-reduced as flat NCCL buckets after each backward, G's
                # buffers broadcast before its forward: the collectives DDP would issue.
                self._broadcast_module_states(self.G, self.D)
                self.G_ema.load_state_dict(self.G.state_dict())
        for m in (self.G, self.G_ema, self.D, self.A, self.coord):
            m.requires_grad_(False)

        self.auxin = {}
        if "dusty_v2" in cfg.model.generator.arch:
            # a stride-0 view: the synthesis network sees at once that the grid is batch-shared
            self.auxin["angle"] = self.coord.angle.expand(self.B, -1, -1, -1)

        self.adversarial_loss = GANLoss(tr.gan_objective).to(self.device)
        self.gp_weight, self.gp_every = 0.0, 0
        lazy_D = 1.0
        if tr.loss.get("gp", 0) > 0:
            self.gp_every = tr.lazy.gp
            self.gp_weight = tr.loss.gp * tr.lazy.gp          # trainer.py:146
            lazy_D = tr.lazy.gp / (tr.lazy.gp + 1.0)
        if tr.loss.get("pl", 0) > 0:
            raise NotImplementedError("path-length regularisation is dead code in the reference "
                                      "(loss.pl = 0 in every config; the branch is broken)")
        lr_g, lr_d = tr.lr.generator, tr.lr.discriminator
        # own multi-tensor Adam (csrc/optim.cu); the generator's EMA lerp rides in its pass
        self.optim_G = Adam(self.G.parameters(), lr=lr_g.alpha, betas=(float(lr_g.beta1), float(lr_g.beta2)),
                            ema_params=self.G_ema.parameters())
        self.optim_D = Adam(self.D.parameters(), lr=lr_d.alpha * lazy_D,
                            betas=(float(lr_d.beta1 ** lazy_D), float(lr_d.beta2 ** lazy_D)))
        self._G_params = list(self.G.parameters())     # cached: no module-tree walk per step
        self._D_params = list(self.D.parameters())
        # flat fp32 gradient buckets of the graphed multi-GPU path: packed by one kernel, reduced
        # by one NCCL call, consumed in place by the optimiser (no cat / split / copy-back)
        self._flat = {}
        self._ema_bufs = None
        self.z_dim = cfg.model.generator.mapping_kwargs.in_ch
        # CUDA graphs for the static-shape segments (the step is launch-bound at B=64):
        #   * the no-grad generator forward of the D step (z drawn inside the graph)
        # ADA (data-dependent padding) and the discriminator stay eager.
        self._g_graph = None
        self._G_train_callable = None         # graphed forward+backward of the G step
        self._D_callables = {}                # graphed D: "frozen" (G step) / "train" (D step)
        self._D_launches = {}
        self._g_graph_out = None
        self._g_graph_launches = 0
        self.graph_replayed_launches = 0
        self.warmup_fade_imgs = tr.warmup.fade_kimg * 1e3
        self.blur_sigma = 0.0
        self.dropout_ratio = 0.0
        self.scalar_names = []

    @staticmethod
    @torch.no_grad()
    def _broadcast_module_states(*modules, src=0):
        """Rank `src`'s parameters and buffers to every rank (what DDP's constructor does),
        as one flat broadcast per dtype."""
        by_dtype = {}
        for m in modules:
            for t in list(m.parameters()) + list(m.buffers()):
                by_dtype.setdefault(t.dtype, []).append(t.data)
        for tensors in by_dtype.values():
            flat = torch.cat([t.reshape(-1) for t in tensors])
            dist.broadcast(flat, src)
            torch._foreach_copy_(tensors, [c.reshape(t.shape) for t, c in
                                           zip(tensors, flat.split([t.numel() for t in tensors]))])

    # ------------------------------------------------------------------ inputs
    def sample_z(self, batch_size):
        return torch.randn(batch_size, self.z_dim, device=self.device)

    # -- input pipeline: the NEXT step's host -> device copy runs on a copy stream under this step
    def _issue_copy(self, raw):
        if self.device.type != "cuda" or all((not torch.is_tensor(v)) or v.is_cuda for v in raw.values()):
            return raw, None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._copy_stream):
            dev = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in raw.items()}
            event = torch.cuda.Event()
            event.record(self._copy_stream)
        return dev, event

    def _next_batch(self):
        """next(self.batch_iter) with the following batch's pinned-memory upload already issued
        (a batch prefetched from an iterator that has since been replaced is dropped)."""
        pre = self._prefetched
        if pre is not None and pre[0] is self.batch_iter:
            cur, event = pre[1], pre[2]
        else:
            cur, event = self._issue_copy(next(self.batch_iter))
        try:
            self._prefetched = (self.batch_iter,) + self._issue_copy(next(self.batch_iter))
        except StopIteration:
            self._prefetched = None
        if event is not None:
            main = torch.cuda.current_stream()
            main.wait_event(event)
            for v in cur.values():
                if torch.is_tensor(v):
                    v.record_stream(main)
        return cur

    def fetch_reals(self, raw_batch):
        """trainer.py:211-217: depth -> inverse depth in [-1, 1], dropped rays -> raydrop_const."""
        depth = raw_batch["depth"].to(self.device, non_blocking=True)
        mask = raw_batch["mask"].to(self.device, non_blocking=True)
        x = sigmoid_to_tanh(self.coord.convert(depth, "depth", "inv_depth_norm"))
        x = mask * x + (1 - mask) * self.cfg.dataset.raydrop_const
        return {"image": x, "raydrop_mask": mask}

    def set_warmup_params(self, iteration):
        imgs = int(iteration * self.cfg.training.batch_size)
        w = self.cfg.training.warmup
        fade = max(1 - imgs / self.warmup_fade_imgs, 0) if self.warmup_fade_imgs > 0 else 0
        self.blur_sigma = fade * w.blur_init_sigma
        self.dropout_ratio = fade * w.dropout_init_ratio

    def warmup(self, x):
        """trainer.py:234-245: optional gaussian blur + bernoulli ray dropout, fading out."""
        blur_size = np.floor(self.blur_sigma * 3)
        if blur_size > 0:
            k = torch.arange(-blur_size, blur_size + 1, device=x.device)
            x = filter2d(x, k.div(self.blur_sigma).square().neg().exp2())
        if self.dropout_ratio > 0:
            keep = torch.bernoulli(torch.full_like(x, 1 - self.dropout_ratio))
            x = keep * x + (1 - keep) * self.cfg.dataset.raydrop_const
        return x

    # ------------------------------------------------------------------ graphed G forward
    def _sync_G_buffers(self):
        """DDP(broadcast_buffers=True) semantics for the graphed path: rank 0's buffers win.
        The generator's fp32 buffers are re-homed ONCE as views of one flat tensor (before any
        graph is captured), so the per-forward sync is a single broadcast of that tensor: the
        cat / broadcast / split / copy form cost 0.35 ms per call, twice per iteration -- all of
        the step's multi-GPU overhead beside the two gradient all-reduces."""
        if self.world_size > 1:
            flat = getattr(self, "_G_sync_flat", None)
            if flat is None:
                bufs = [b for b in self.G_module.buffers()
                        if b.dtype == torch.float32 and b.is_contiguous() and 0 < b.numel() <= (1 << 20)]
                flat = torch.empty(sum(b.numel() for b in bufs), dtype=torch.float32, device=bufs[0].device)
                off = 0
                with torch.no_grad():
                    for b in bufs:
                        n = b.numel()
                        flat[off:off + n].copy_(b.reshape(-1))
                        b.data = flat[off:off + n].view(b.shape)
                        off += n
                self._G_sync_flat = flat
            dist.broadcast(flat, 0)

    class _KeepState:
        """Graph warm-up and capture run extra real forwards of the live train-mode generator:
        each one moves `ema_var` / `w_avg` and consumes device RNG.  Snapshot the buffers and
        the RNG state, restore them afterwards, so that a graphed trainer starts from exactly
        the state an eager one (or the reference) has for the same seed."""

        def __init__(self, module, device):
            self.module, self.device = module, device

        def __enter__(self):
            self.bufs = [b.detach().clone() for b in self.module.buffers()]
            self.rng = torch.cuda.get_rng_state(self.device)

        def __exit__(self, *exc):
            torch.cuda.synchronize(self.device)
            with torch.no_grad():
                for b, c in zip(self.module.buffers(), self.bufs):
                    b.copy_(c)
            torch.cuda.set_rng_state(self.rng, self.device)
            return False

    def _G_train_forward(self, z):
        """x_fake for the G step, with autograd.  With CUDA graphs: forward and backward of
        the whole generator are two graph launches (torch.cuda.make_graphed_callables)."""
        if not self.cuda_graphs:
            return self.G(z, **self.auxin)["image"]
        self._sync_G_buffers()
        if self._G_train_callable is None:
            from .. import _cabi
            wrapper = _GImage(self.G_module, self.auxin)
            n0 = _cabi.launch_count()
            try:
                with self._KeepState(self.G_module, self.device):
                    self._G_train_callable = torch.cuda.make_graphed_callables(wrapper, (z.clone(),))
                # 3 warm-up passes + 1 capture, each forward + backward
                self._G_train_launches = (_cabi.launch_count() - n0) // 4
            except Exception as exc:           # capture not possible: stay eager, loudly
                import warnings
                warnings.warn(f"CUDA-graph capture of the generator step failed ({exc!r}); "
                              "running it eagerly")
                self._G_train_callable = wrapper
                self._G_train_launches = 0
        self.graph_replayed_launches += self._G_train_launches
        return self._G_train_callable(z)

    def _D_forward(self, x, mode):
        """mode 'frozen': D frozen, x differentiable (G step); 'train': D trainable, x is the
        detached real+fake stack (D step).  Graph-captured per mode; the R1 step (double
        backward) always runs eagerly."""
        sub = 2 if mode == "train" else 1
        if not self.cuda_graphs:
            fn = self._D_callables.get(mode)
            if fn is None:
                fn = self._D_callables[mode] = _DLogits(self.D, sub)
            return fn(x)
        fn = self._D_callables.get(mode)
        if fn is None:
            from .. import _cabi
            wrapper = _DLogits(self.D_module, sub)
            n0 = _cabi.launch_count()
            try:
                sample = x.detach().clone().requires_grad_(x.requires_grad)
                fn = torch.cuda.make_graphed_callables(wrapper, (sample,))
                self._D_launches[mode] = (_cabi.launch_count() - n0) // 4
            except Exception as exc:
                import warnings
                warnings.warn(f"CUDA-graph capture of the discriminator ({mode}) failed ({exc!r}); "
                              "running it eagerly")
                fn = wrapper
                self._D_launches[mode] = 0
            self._D_callables[mode] = fn
        self.graph_replayed_launches += self._D_launches[mode]
        return fn(x)

    def _reduced_grads(self, name, params):
        """Graphed multi-GPU path: (views of the summed flat bucket, 1 / world_size) for the
        parameters that have a gradient; (None, 1.0) when there is nothing to exchange (one GPU,
        or DDP already averaged the gradients in its hooks)."""
        if not (self.world_size > 1 and self.cuda_graphs):
            return None, 1.0
        grads = [p.grad for p in params if p.grad is not None]
        sizes = [g.numel() for g in grads]
        entry = self._flat.get(name)
        if entry is None or entry[2] != sizes:
            flat = torch.empty(sum(sizes), dtype=torch.float32, device=self.device)
            entry = (flat, [v.view_as(g) for v, g in zip(flat.split(sizes), grads)], sizes)
            self._flat[name] = entry
        flat, views, _ = entry
        self._pack(views, [g if g.dtype == torch.float32 and g.is_contiguous() else g.float().contiguous()
                           for g in grads])
        dist.all_reduce(flat)
        return views, 1.0 / self.world_size

    _pack = staticmethod(multi_copy)        # the packing kernel (host-logic tests on gloo swap it)

    def _update_D(self):
        """Gradient exchange + Adam step of the discriminator.  (Running the exchange and the
        update on a side stream under the next generator forward was measured at 2 GPUs: 18.25
        vs 17.97 ms per step -- the NCCL kernel and the forward contend for the same SMs -- so
        it stays on the main stream.)"""
        views, scale = self._reduced_grads("D", self._D_params)
        self.optim_D.step(grads=views, grad_scale=scale)

    def _fake_images_nograd(self, B):
        """x_fake for the D step (no graph of G is needed: trainer.py:380-383)."""
        if not self.cuda_graphs:
            with torch.no_grad():
                return self.G(self.sample_z(B), **self.auxin)["image"]
        from .. import _cabi
        self._sync_G_buffers()
        if self._g_graph is None:
            with self._KeepState(self.G_module, self.device):
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side), torch.no_grad():
                    for _ in range(2):          # warm-up outside capture (lazy inits, autotune)
                        self.G_module(self.sample_z(B), **self.auxin)
                torch.cuda.current_stream(self.device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                n0 = _cabi.launch_count()
                with torch.no_grad(), torch.cuda.graph(graph):
                    out = self.G_module(self.sample_z(B), **self.auxin)["image"]
                self._g_graph_launches = _cabi.launch_count() - n0
            self._g_graph, self._g_graph_out = graph, out
        self._g_graph.replay()
        self.graph_replayed_launches += self._g_graph_launches
        return self._g_graph_out

    # ------------------------------------------------------------------ one iteration
    def step(self, iteration):
        tr = self.cfg.training
        if not self.G.training:
            self.G.train()
        self.set_warmup_params(iteration)
        B = self.B
        scalars = OrderedDict()
        x_real = self.fetch_reals(self._next_batch())["image"]
        ema_imgs = int(tr.ema_kimg * 1e3)                       # trainer.py:459-466
        if tr.ema_rampup is not None:
            ema_imgs = min(ema_imgs, iteration * tr.batch_size * tr.ema_rampup)
        ema_decay = 0.5 ** (tr.batch_size / max(ema_imgs, 1e-8))

        # ---- G step (trainer.py:262-301)
        set_requires_grad(self._G_params, True)
        self.optim_G.zero_grad(set_to_none=True)
        x_fake = self._G_train_forward(self.sample_z(B))
        y_fake = self._D_forward(self.A(self.warmup(x_fake)), "frozen")
        y_real = None
        if tr.gan_objective in ("ragan", "rahinge", "ralsgan"):      # trainer.py:263,281-286
            y_real = _DLogits(self.D, 1)(self.A(self.warmup(x_real)).detach())
        loss_gan = self.adversarial_loss(y_real, y_fake, "G")
        (tr.loss.gan * loss_gan).backward()
        views, scale = self._reduced_grads("G", self._G_params)
        # Adam + the EMA lerp of G_ema's parameters towards the new weights in one pass
        self.optim_G.step(grads=views, grad_scale=scale, ema_weight=1.0 - ema_decay)
        scalars["loss/G/adversarial"] = loss_gan.detach()
        set_requires_grad(self._G_params, False)

        # ---- D step (trainer.py:373-412)
        set_requires_grad(self._D_params, True)
        self.optim_D.zero_grad(set_to_none=True)
        x_fake = self._fake_images_nograd(B)
        # real and fake go through warm-up, ADA and D as ONE stacked batch (per-sample
        # transforms, per-half minibatch statistics): same math, half the launches
        x_both = self.A(self.warmup(torch.cat([x_real, x_fake], dim=0))).detach()
        y_real, y_fake = self._D_forward(x_both, "train").chunk(2, dim=0)
        self.A.cumulate(y_real)
        loss_gan = self.adversarial_loss(y_real, y_fake, "D")
        (tr.loss.gan * loss_gan).backward()
        self._update_D()
        scalars["loss/D/output/real"] = y_real.mean().detach()
        scalars["loss/D/output/fake"] = y_fake.mean().detach()
        scalars["loss/D/adversarial"] = loss_gan.detach()

        # ---- lazy R1 (trainer.py:419-451)
        if self.gp_weight > 0 and iteration % self.gp_every == 0:
            self.optim_D.zero_grad(set_to_none=True)
            x_gp = x_real.detach().requires_grad_()
            y_real = self.D(self.A(self.warmup(x_gp)))
            (grads,) = torch.autograd.grad(outputs=[y_real.sum()], inputs=[x_gp], create_graph=True)
            r1 = DF.sumsq_rows(grads).mean()
            loss = (self.gp_weight / 2) * r1 + 0.0 * y_real.squeeze()[0]
            loss.backward()
            self._update_D()
            scalars["loss/D/gradient_penalty"] = r1.detach()
        set_requires_grad(self._D_params, False)

        # ---- exit (trainer.py:459-476)
        self._copy_ema_buffers()          # parameters were lerped in the G step's Adam pass
        if iteration % tr.lazy.ada == 0:
            scalars["stats/ada_rt"] = self.A.update_p().detach().reshape(())
            scalars["stats/ada_p"] = self.A.p.detach().reshape(())

        packed = torch.stack([v.float().reshape(()) for v in scalars.values()])
        if self.world_size > 1:
            dist.all_reduce(packed)
            packed /= self.world_size
        self.scalar_names = list(scalars.keys())
        self.last_stats = dict(ema_decay=ema_decay, blur_sigma=self.blur_sigma,
                               dropout_ratio=self.dropout_ratio)
        return packed            # device tensor; names in self.scalar_names

    @torch.no_grad()
    def _copy_ema_buffers(self):
        """The buffer half of ema_inplace (trainer.py:38-41): G_ema's buffers <- G's."""
        if self._ema_bufs is None:
            pairs = list(zip(self.G_ema.buffers(), self.G_module.buffers()))
            own = [(a, b) for a, b in pairs if a.dtype == torch.float32 and b.dtype == torch.float32
                   and a.is_contiguous() and b.is_contiguous() and a.numel() > 0]
            rest = [(a, b) for a, b in pairs if not any(a is x for x, _ in own)]
            self._ema_bufs = ([a for a, _ in own], [b for _, b in own], [a for a, _ in rest], [b for _, b in rest])
        da, sa, db, sb = self._ema_bufs
        multi_copy(da, sa)
        if db:
            torch._foreach_copy_(db, sb)

    def scalars_to_host(self, packed):
        vals = packed.detach().cpu().tolist()
        out = dict(zip(self.scalar_names, vals))
        out.update({f"stats/{k}": v for k, v in self.last_stats.items()})
        return out

    @torch.no_grad()
    def sample(self, z, ema=True):
        net = self.G_ema if ema else self.G_module
        was = net.training
        net.eval()
        out = net(z, **{k: v[: z.shape[0]] for k, v in self.auxin.items()})
        net.train(was)
        return out

    def state_dict(self, step):
        """Checkpoint payload with the reference's keys (trainer.py:551-567)."""
        return {"cfg": self.cfg, "step": step, "angle": self.coord.angle.detach().cpu(),
                "G": self.G_module.state_dict(), "D": self.D_module.state_dict(),
                "G_ema": self.G_ema.state_dict(), "A": self.A.state_dict(),
                "optim_G": self.optim_G.state_dict(), "optim_D": self.optim_D.state_dict()}

    def load_state_dict(self, state, strict=True):
        """Resume from a checkpoint payload with the reference's keys -- one written by
        `state_dict(step)` here or by the reference's `Trainer.save_checkpoint`
        (trainer.py:184-196, 551-567).  Weights are copied into the existing tensors, so captured
        CUDA graphs stay valid.  Returns the iteration to continue from (`step // batch_size`)."""
        self.G_module.load_state_dict(state["G"], strict=strict)
        self.D_module.load_state_dict(state["D"], strict=strict)
        self.G_ema.load_state_dict(state["G_ema"], strict=strict)
        self.A.load_state_dict(state["A"])
        self.optim_G.load_state_dict(state["optim_G"])
        self.optim_D.load_state_dict(state["optim_D"])
        return int(state["step"]) // int(self.cfg.training.batch_size)
