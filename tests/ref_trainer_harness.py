"""Assembles the REFERENCE'S `gans.trainer.Trainer` for CPU execution without running its
`__init__` (which needs a CUDA rank, NCCL and the KITTI files): same attributes, same optimiser
settings, DDP over a one-rank gloo group the caller has initialised.  Used by the live
cross-check of `Trainer.step` (tests/test_reference_live.py) and by the CPU-port calibration
(tests/golden/calibrate_cpu_port.py).  Build container only (imports /root/reference)."""
import copy
import os

import torch

from oracle import ref_import


def build_reference_trainer(g_cfg, d_cfg, batch_size, resolution, batches, p_init=0.0):
    """Returns (trainer, G, D); `batches` is an iterable of {"depth", "mask"} dicts."""
    ref_import.install()
    from torch.nn.parallel import DistributedDataParallel as DDP
    from gans import trainer as rtr
    from gans.augment.adaptive_augment import AdaptiveAugment
    from gans.coords import CoordBridge
    from gans.models.builder import build_discriminator, build_generator
    from gans.models.loss import GANLoss

    B, (H, W) = batch_size, resolution
    g_cfg, d_cfg = dict(g_cfg), dict(d_cfg)
    cfg = ref_import.to_attr(dict(
        training=dict(batch_size_per_gpu=B, batch_size=B, num_gpus=1, gan_objective="nsgan",
                      amp=dict(main=False, reg=False), loss=dict(gan=1.0, gp=16.0, pl=0.0),
                      lazy=dict(gp=16, pl=4, ada=4), ema_kimg=10, ema_rampup=0.05,
                      warmup=dict(fade_kimg=200, blur_init_sigma=0, dropout_init_ratio=0.5)),
        dataset=dict(raydrop_const=-1, min_depth=1.45, max_depth=80.0),
        model=dict(generator=dict(arch=g_cfg["arch"], mapping_kwargs=dict(
            in_ch=g_cfg.get("mapping_kwargs", g_cfg["synthesis_kwargs"])["in_ch"])))))
    G = build_generator(ref_import.to_attr(g_cfg))
    D = build_discriminator(ref_import.to_attr(d_cfg))
    T = object.__new__(rtr.Trainer)
    T.cfg, T.device = cfg, torch.device("cpu")
    T.G_ema = copy.deepcopy(G).eval()
    T.A = AdaptiveAugment(p_init=p_init, p_target=0.6, kimg=500, lr_flip=1, ud_flip=1, int_trans=1,
                          iso_scale=1, frac_trans=1, brightness=1, contrast=1, luma_flip=1, hue=1,
                          saturation=1)
    T.coord = CoordBridge(H, W, 1.45, 80.0,
                          os.path.join(ref_import.REFERENCE_ROOT, "data/coords/kitti_raw.npy")).eval()
    T.G, T.D = DDP(G, broadcast_buffers=True), DDP(D, broadcast_buffers=False)
    T.ddp_models = (T.G, T.D)
    for m in (T.G, T.G_ema, T.D, T.A, T.coord):
        m.requires_grad_(False)
    T.auxin = {}
    if "dusty_v2" in g_cfg["arch"]:
        T.auxin["angle"] = T.coord.angle.repeat_interleave(B, dim=0)
    T.iter_train_loader = iter(batches)
    T.adversarial_loss = GANLoss("nsgan")
    lazy = 16 / 17.0
    T.optim_G = torch.optim.Adam(T.G.parameters(), lr=0.002, betas=(0.0, 0.99))
    T.optim_D = torch.optim.Adam(T.D.parameters(), lr=0.002 * lazy, betas=(0.0, 0.99 ** lazy))
    for name in ("scaler_D", "scaler_G", "scaler_r1", "scaler_pl"):
        setattr(T, name, rtr.GradScaler(enabled=False))
    T.warmup_fade_kimg, T.blur_sigma, T.dropout_ratio = 200e3, 0, 0
    T.iters_to_imgs = lambda i: int(i * B)
    return T, G, D


def record_step(T, G, D, iteration):
    """Runs T.step(iteration) with every random draw logged; returns (scalars, log, g_grads, d_grads)."""
    log = {"randn": [], "rand": [], "uniform_": [], "bernoulli": [], "affine": [], "color": []}
    quiet = [0]
    real = dict(randn=torch.randn, rand=torch.rand, bernoulli=torch.bernoulli, uniform_=torch.Tensor.uniform_)

    def recorder(name):
        def fn(*a, **k):
            out = real[name](*a, **k)
            if not quiet[0]:
                log[name].append(out.detach().clone())
            return out
        return fn

    def sampler(name, orig):
        def fn(*a, **k):
            quiet[0] += 1
            try:
                out = orig(*a, **k)
            finally:
                quiet[0] -= 1
            log[name].append(out.detach().clone())
            return out
        return fn

    T.A.sample_affine = sampler("affine", T.A.sample_affine)
    T.A.sample_color = sampler("color", T.A.sample_color)
    d_grads = []
    d_step = T.optim_D.step

    def recording_step(*a, **k):
        d_grads.append({n: p.grad.detach().clone() for n, p in D.named_parameters() if p.grad is not None})
        return d_step(*a, **k)

    T.optim_D.step = recording_step
    torch.randn, torch.rand, torch.bernoulli = recorder("randn"), recorder("rand"), recorder("bernoulli")
    torch.Tensor.uniform_ = recorder("uniform_")
    try:
        scalars = T.step(iteration)
    finally:
        torch.randn, torch.rand, torch.bernoulli = real["randn"], real["rand"], real["bernoulli"]
        torch.Tensor.uniform_ = real["uniform_"]
    g_grads = {n: p.grad.detach().clone() for n, p in G.named_parameters() if p.grad is not None}
    return scalars, log, g_grads, d_grads
