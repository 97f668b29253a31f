"""Live cross-checks of the oracle against the UNMODIFIED reference imported from /root/reference
(CPU tensors), on fresh random inputs and random module configurations -- beyond the committed
golden fixtures.  Runs only in the build container: skipped where the reference tree is absent
(the GPU box) or where its `fused` extension has not been built yet (importing the reference's
ops compiles it; tests/golden/make_golden.py does that once)."""
import os

import numpy as np
import pytest
import torch

from oracle import dusty_oracle as O
from oracle import ref_import

_EXT = os.path.join(os.environ.get("TORCH_EXTENSIONS_DIR", "/tmp/torch_ext"), "fused", "fused.so")
pytestmark = pytest.mark.skipif(
    not (ref_import.available() and os.path.exists(_EXT)),
    reason="needs /root/reference and its prebuilt `fused` extension (build container only)")


@pytest.fixture(scope="module")
def rops():
    ref_import.install()
    from gans.models import ops
    return ops


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("seed", range(4))
def test_resample_family_random(rops, seed):
    g = torch.Generator().manual_seed(100 + seed)
    C, H, W = int(torch.randint(1, 5, (1,), generator=g)), 2 * int(torch.randint(2, 9, (1,), generator=g)), \
        4 * int(torch.randint(2, 9, (1,), generator=g))
    x = torch.randn(2, C, H, W, generator=g)
    for kw in (dict(up=2), dict(down=2), dict(), dict(window=[1, 2, 1]), dict(window=[1, 2, 1], direction="h"),
               dict(window=[1, 2, 1], direction="w")):
        ref = rops.Resample(**kw)(x)
        got = O.resample(x, up=kw.get("up", 1), down=kw.get("down", 1), window=kw.get("window", (1, 3, 3, 1)),
                         direction=kw.get("direction", "hw"))
        assert got.shape == ref.shape, kw
        close(got, ref, rtol=1e-5, atol=1e-6)
    close(O.blur_vh(x[:, :1]), rops.BlurVH()(x[:, :1]))
    for pad, ring, mode in ((1, True, "replicate"), (2, True, "reflect"), ((1, 2, 0, 1), False, "replicate")):
        close(O.pad2d(x, pad, ring=ring, mode=mode), rops.Pad(pad, ring=ring, mode=mode)(x), rtol=0, atol=0)


@pytest.mark.parametrize("seed", range(4))
def test_fused_leaky_relu_and_modconv_random(rops, seed):
    g = torch.Generator().manual_seed(200 + seed)
    C = int(torch.randint(2, 9, (1,), generator=g))
    x = torch.randn(3, C, 4, 8, generator=g, requires_grad=True)
    b = torch.randn(C, generator=g)
    act = rops.FusedLeakyReLU(C)
    with torch.no_grad():
        act.bias.copy_(b)
    ref = act(x)
    close(O.bias_act(x, b), ref)
    (gr,) = torch.autograd.grad(ref.square().sum(), x)
    (go,) = torch.autograd.grad(O.bias_act(x, b).square().sum(), x)
    close(go, gr)
    # modulated 1x1 convolution, eval and training (EMA side effect)
    O_ch, M = int(torch.randint(1, 7, (1,), generator=g)), 8
    torch.manual_seed(300 + seed)
    for demod, bias, ema in ((True, False, True), (False, True, True), (True, False, False)):
        m = rops.ModConv2d(in_ch=C, out_ch=O_ch, mod_ch=M, ksize=1, stride=1, padding=0, demod=demod,
                           bias=bias, ema=ema)
        with torch.no_grad():
            m.ema_var.fill_(0.7)
            if bias:
                m.bias.normal_()
        style = torch.randn(3, M, generator=g)
        xin = torch.randn(3, C, 4, 8, generator=g)
        ev = m.ema_var.clone() if ema else torch.tensor(1.0)
        m.eval()
        ref = m(xin, style)
        got, _ = O.modconv(xin, style, m.weight, m.mod.module.weight, m.mod.module.bias, ev, demod=demod,
                           bias=m.bias if bias else None)
        close(got, ref, rtol=1e-4, atol=1e-5)
        if ema:                     # training mode: the EMA is updated before it is used
            m.train()
            ref_t = m(xin, style)
            got_t, new_ev = O.modconv(xin, style, m.weight, m.mod.module.weight, m.mod.module.bias, ev,
                                      demod=demod, bias=m.bias if bias else None, training=True,
                                      ema_decay=m.ema_decay)
            close(got_t, ref_t, rtol=1e-4, atol=1e-5)
            close(new_ev, m.ema_var, rtol=1e-6, atol=0)


@pytest.mark.parametrize("seed", range(2))
def test_small_generator_eval_random_weights(rops, seed):
    """A freshly initialised small dusty_v2 generator (new seed, new Fourier basis) in eval mode
    against the oracle driven by its state_dict."""
    from gans.models.builder import build_generator
    from small_cfgs import G_SMALL
    torch.manual_seed(400 + seed)
    np.random.seed(400 + seed)
    G = build_generator(ref_import.to_attr(G_SMALL)).eval()
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    B = 2
    z = torch.randn(B, 16)
    el = torch.linspace(0.05, -0.41, 16)[:, None].expand(16, 64)
    az = -((torch.arange(64) + 0.5) / 64 * 2 * np.pi - np.pi)[None].expand(16, 64)
    angle = torch.stack([el, az], 0)[None].repeat(B, 1, 1, 1).contiguous()
    torch.manual_seed(500 + seed)
    with torch.no_grad():
        ref = G(z, angle=angle)
    torch.manual_seed(500 + seed)
    u = torch.rand(B, 1, 16, 64)
    with torch.no_grad():
        got = O.generator(sd, z, angle, u)
    for k in ("image_orig", "raydrop_logit", "image"):
        close(got[k], ref[k], rtol=1e-4, atol=1e-5)
    assert torch.equal(got["raydrop_mask"], ref["raydrop_mask"])
    assert int(got["raydrop_mask"].sum()) == int(ref["raydrop_mask"].sum())


def test_inversion_losses_random(rops):
    from gans import inversion as rinv
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(600)
    for H, W, level in ((8, 32, None), (16, 64, 2), (32, 32, 3)):
        gen = (torch.rand(2, 1, H, W, generator=g) * 0.8 + 0.1).requires_grad_()
        ref = torch.rand(2, 1, H, W, generator=g) * 0.8 + 0.1
        mask = (torch.rand(2, 1, H, W, generator=g) < 0.6).float()
        for loss_fn, name, rel in ((F.l1_loss, "l1", True), (F.mse_loss, "l2", False)):
            r = rinv.MultiScaleMaskedLoss(loss_fn, level=level, relative=rel)(gen, ref, mask)
            o = O.multiscale_masked_loss(gen, ref, mask, level=level, loss=name, relative=rel)
            close(o, r, rtol=1e-5, atol=1e-6)
    lat = torch.randn(3, 10, 16, generator=g)
    close(O.geocross_loss(lat), rinv.geocross_loss(lat), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("seed", range(2))
def test_small_discriminator_and_r1_random_weights(rops, seed):
    from gans.models.builder import build_discriminator
    from gans.models.loss import GANLoss
    from small_cfgs import D_SMALL
    torch.manual_seed(700 + seed)
    D = build_discriminator(ref_import.to_attr(D_SMALL))
    with torch.no_grad():
        for n, p in D.named_parameters():
            if "bias" in n:
                p.normal_(0, 0.3)
    sd = {k: v.clone() for k, v in D.state_dict().items()}
    x = torch.tanh(torch.randn(4, 1, 16, 64)).requires_grad_()
    y_ref = D(x)
    xo = x.detach().clone().requires_grad_()
    y_got = O.discriminator(sd, xo)
    close(y_got, y_ref, rtol=1e-4, atol=1e-5)
    # R1 penalty (trainer.py:426-447) and the non-saturating losses (loss.py:39-41,68-69)
    (g_ref,) = torch.autograd.grad(y_ref.sum(), x, create_graph=True)
    (g_got,) = torch.autograd.grad(y_got.sum(), xo, create_graph=True)
    r1_ref = g_ref.pow(2).sum(dim=[1, 2, 3]).mean()
    close(O.r1_penalty(g_got), r1_ref, rtol=1e-4, atol=1e-7)
    crit = GANLoss("nsgan")
    y_fake = torch.randn(4, 1)
    close(O.nsgan_g(y_fake), crit(None, y_fake, "G"), rtol=1e-6, atol=1e-7)
    close(O.nsgan_d(y_ref.detach(), y_fake), crit(y_ref.detach(), y_fake, "D"), rtol=1e-6, atol=1e-7)


def test_coord_bridge_random(rops):
    from gans.coords import CoordBridge
    angle_file = os.path.join(ref_import.REFERENCE_ROOT, "data/coords/kitti_raw.npy")
    for H, W in ((16, 64), (64, 512)):
        cb = CoordBridge(H, W, 1.45, 80.0, angle_file)
        assert torch.equal(O.angle_grid(np.load(angle_file), H, W), cb.angle)
        g = torch.Generator().manual_seed(800 + H)
        depth = 80.0 * torch.rand(2, 1, H, W, generator=g) ** 2
        x = cb.convert(depth.clone(), "depth", "inv_depth_norm")
        close(O.depth_to_inv_depth_norm(depth, 1.45, 80.0), x, rtol=0, atol=0)
        pm = cb.convert(x.clone(), "inv_depth_norm", "point_map")
        ps = cb.convert(x.clone(), "inv_depth_norm", "point_set")
        opm, ops_, cnt = O.inv_depth_norm_to_points(x.clone(), cb.angle, 1.45, 80.0)
        close(opm, pm, rtol=0, atol=0)
        close(ops_, ps, rtol=0, atol=0)
        valid = (x > 1e-11).float() * cb.get_mask(x / 1.45, "inv_depth").float()
        assert cnt == int(valid.sum().item())
        close(O.inv_depth_norm_to_depth_norm(x.clone(), 1.45, 80.0), cb.convert(x.clone(), "inv_depth_norm", "depth_norm"),
              rtol=0, atol=0)


def test_utils_sampler_and_seeding_match_reference(rops):
    """gans/utils.py host helpers: the infinite windowed-shuffle sampler yields the reference's
    index stream for every (seed, rank, replicas, window); init_random_seed leaves every RNG in the
    same state."""
    import random
    import gans.utils as rutils
    from dusty_gan_v2_b200.gans import utils as mine
    data = list(range(37))
    for seed, rank, reps, win, shuffle in ((0, 0, 1, 0.5, True), (3, 1, 4, 0.5, True), (7, 2, 3, 0.1, True),
                                           (1, 0, 2, 0.0, True), (5, 1, 2, 0.5, False)):
        # the reference's constructor calls Sampler.__init__(dataset), which torch 2.11 rejects:
        # fill the attributes it would set and run the reference's own __iter__
        ref_s = object.__new__(rutils.InfiniteSampler)
        ref_s.__dict__.update(dataset=data, rank=rank, num_replicas=reps, shuffle=shuffle, seed=seed,
                              window_size=win)
        a = iter(ref_s)
        b = iter(mine.InfiniteSampler(data, rank=rank, num_replicas=reps, shuffle=shuffle, seed=seed,
                                      window_size=win))
        assert [int(next(a)) for _ in range(150)] == [int(next(b)) for _ in range(150)]
    draws = []
    for mod in (rutils, mine):
        mod.init_random_seed(11, rank=2)
        draws.append((random.random(), float(np.random.rand()), float(torch.rand(1))))
    assert draws[0] == draws[1]
    x = torch.linspace(-1, 1, 9)
    assert torch.equal(mine.tanh_to_sigmoid(x), rutils.tanh_to_sigmoid(x))
    assert torch.equal(mine.sigmoid_to_tanh(x), rutils.sigmoid_to_tanh(x))
    g = torch.Generator().manual_seed(1)
    a, b, m = torch.rand(2, 1, 4, 4, generator=g), torch.rand(2, 1, 4, 4, generator=g), torch.ones(2, 1, 4, 4)
    for d in ("l1", "l2"):
        assert torch.allclose(mine.masked_loss(a, b, m, d), rutils.masked_loss(a, b, m, d))


_EXT_UFD = os.path.join(os.environ.get("TORCH_EXTENSIONS_DIR", "/tmp/torch_ext"), "upfirdn2d", "upfirdn2d.so")


@pytest.mark.skipif(not os.path.exists(_EXT_UFD), reason="reference `upfirdn2d` extension not prebuilt")
@pytest.mark.parametrize("seed,hw", [(0, (16, 64)), (1, (16, 64)), (2, (32, 128))])
def test_ada_pipeline_random_transforms(rops, seed, hw):
    """AdaptiveAugment.forward (adaptive_augment.py:471-545) with freshly sampled affine / colour
    transforms pinned on both sides: value, gradient, and the gradient of a gradient-norm penalty
    (the path the R1 step takes through ADA), oracle vs the live reference."""
    from gans.augment import adaptive_augment as ra
    torch.manual_seed(900 + seed)
    np.random.seed(900 + seed)
    ada = ra.AdaptiveAugment(p_init=0.8, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1, frac_trans=1,
                             brightness=1, contrast=1, luma_flip=1, hue=1, saturation=1)
    B, (H, W) = 3, hw
    G = ada.sample_affine(B, H, W)
    C = ada.sample_color(B)
    ada.sample_affine = lambda *a, **k: G.clone()
    ada.sample_color = lambda *a, **k: C.clone()
    x = torch.tanh(torch.randn(B, 1, H, W))
    xr = x.clone().requires_grad_()
    yr = ada(xr)
    xo = x.clone().requires_grad_()
    yo = O.ada_apply(xo, torch.inverse(G), C)
    close(yo, yr, rtol=1e-4, atol=2e-5)
    gy = torch.randn_like(yr)
    (gr,) = torch.autograd.grad(yr, xr, gy, create_graph=True)
    (go,) = torch.autograd.grad(yo, xo, gy, create_graph=True)
    close(go, gr, rtol=1e-4, atol=2e-5)
    v = torch.randn_like(x)
    # second order: d/d(gy-direction) is linear, so probe d<grad, v>/dx through a nonlinearity in y
    (g2r,) = torch.autograd.grad((torch.autograd.grad(yr.square().sum(), xr, create_graph=True)[0] * v).sum(), xr)
    (g2o,) = torch.autograd.grad((torch.autograd.grad(yo.square().sum(), xo, create_graph=True)[0] * v).sum(), xo)
    close(g2o, g2r, rtol=1e-3, atol=1e-4)


@pytest.mark.skipif(not os.path.exists(_EXT_UFD), reason="reference `upfirdn2d` extension not prebuilt")
def test_reference_trainer_step_vs_oracle(rops, tmp_path):
    """The reference's REAL `Trainer.step` (gans/trainer.py:247-482) run on CPU -- the object is
    assembled without `__init__` (which needs a CUDA rank and KITTI files), DDP over gloo with one
    rank, small G / D -- against `O.train_iteration`, the restatement bench.py times as the CPU
    baseline.  Every random draw of the step (z, azimuth shift, Gumbel uniforms, warm-up dropout,
    ADA transforms) is recorded on the way and replayed into the oracle; the oracle is advanced
    phase by phase with the reference's Adam settings (G step -> D step -> lazy R1)."""
    import torch.distributed as dist
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        _trainer_step_vs_oracle()
    finally:
        if own_group:
            dist.destroy_process_group()


def _trainer_step_vs_oracle():
    from small_cfgs import D_SMALL, G_SMALL
    B, H, W = 4, 16, 64
    torch.manual_seed(1000)
    np.random.seed(1000)
    g = torch.Generator().manual_seed(1001)
    batch = {"depth": 1.45 + 78.55 * torch.rand(B, 1, H, W, generator=g),
             "mask": (torch.rand(B, 1, H, W, generator=g) < 0.85).float()}
    from ref_trainer_harness import build_reference_trainer, record_step
    T, G, D = build_reference_trainer(G_SMALL, D_SMALL, B, (H, W), [batch], p_init=0.5)
    lazy = 16 / 17.0

    sdG0 = {k: v.clone() for k, v in G.state_dict().items()}
    sdD0 = {k: v.clone() for k, v in D.state_dict().items()}

    # ---- run the reference's step with every random draw recorded
    scalars, log, g_grads, d_grads = record_step(T, G, D, 0)
    assert [len(log[k]) for k in ("randn", "uniform_", "rand", "bernoulli", "affine", "color")] == [2, 2, 2, 4, 4, 4]
    rnd = dict(z_g=log["randn"][0], z_d=log["randn"][1], shift_g=log["uniform_"][0], shift_d=log["uniform_"][1],
               u_g=log["rand"][0], u_d=log["rand"][1])
    for i, tag in enumerate(("g_fake", "d_real", "d_fake", "r1")):
        rnd[f"keep_{tag}"] = log["bernoulli"][i]
        rnd[f"Ginv_{tag}"] = torch.inverse(log["affine"][i])
        rnd[f"C_{tag}"] = log["color"][i]

    # ---- the oracle, phase by phase
    nograd = ("ema_var", "w_avg", "kernel", "pe.", "raydrop_const")
    sdG = {k: v.clone().requires_grad_(not any(t in k for t in nograd)) for k, v in sdG0.items()}
    sdD = {k: v.clone().requires_grad_("kernel" not in k) for k, v in sdD0.items()}
    angle = T.auxin["angle"]
    x_real = O.fetch_reals(batch["depth"], batch["mask"], 1.45, 80.0)
    optG = torch.optim.Adam([v for v in sdG.values() if v.requires_grad], lr=0.002, betas=(0.0, 0.99))
    optD = torch.optim.Adam([v for v in sdD.values() if v.requires_grad], lr=0.002 * lazy, betas=(0.0, 0.99 ** lazy))

    def apply(opt, sd, grads):
        for k, gr in grads.items():
            sd[k].grad = gr
        opt.step()
        opt.zero_grad(set_to_none=True)

    def check_grads(got, ref, min_n):
        n = 0
        for k, gr in got.items():
            if gr is None or k not in ref:
                continue
            r = ref[k]
            np.testing.assert_allclose(gr.numpy(), r.numpy(), rtol=5e-3, atol=2e-3 * max(float(r.abs().max()), 1e-7))
            n += 1
        assert n >= min_n, n

    # G step (pre-step weights on both sides)
    r = O.train_iteration(sdG, sdD, x_real, angle, rnd, with_r1=False)
    assert float(r["loss_G"]) == pytest.approx(scalars["loss/G/adversarial"], rel=1e-4, abs=1e-6)
    check_grads(r["grads_G"], g_grads, 40)
    # the G-step forward also moved the EMA buffers; the D step sees them and the updated weights
    nb = {}
    with torch.no_grad():
        O.generator(sdG, rnd["z_g"], angle, rnd["u_g"], training=True,
                    shifts_rad=rnd["shift_g"] * (2 * np.pi), new_buffers=nb)
    apply(optG, sdG, r["grads_G"])
    for k, v in nb.items():
        sdG[k] = v.detach().clone()
    # D step
    r = O.train_iteration(sdG, sdD, x_real, angle, rnd, with_r1=False)
    assert float(r["loss_D"]) == pytest.approx(scalars["loss/D/adversarial"], rel=2e-3, abs=1e-5)
    check_grads(r["grads_D"], d_grads[0], 10)
    apply(optD, sdD, r["grads_D"])
    # lazy R1 step on the updated discriminator
    r = O.train_iteration(sdG, sdD, x_real, angle, rnd, with_r1=True)
    assert float(r["r1"]) == pytest.approx(scalars["loss/D/gradient_penalty"], rel=5e-3, abs=1e-7)
    check_grads(r["grads_R1"], d_grads[1], 10)


@pytest.mark.skipif(not os.path.exists(_EXT_UFD), reason="reference `upfirdn2d` extension not prebuilt")
@pytest.mark.parametrize("p", [1.0, 0.3])
def test_ada_host_samplers_match_reference_distribution(rops, p):
    """The mirror samples ADA's transforms on the host (no device sync); their distribution must be
    the reference's (adaptive_augment.py:386-469): per-entry mean and standard deviation of the
    affine (3x3) and colour (4x4) matrices over 40 000 draws, within 5 standard errors."""
    from gans.augment import adaptive_augment as ra
    from dusty_gan_v2_b200.gans.augment.adaptive_augment import AdaptiveAugment as Mine
    kw = dict(p_init=p, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1, frac_trans=1, brightness=1, contrast=1,
              luma_flip=1, hue=1, saturation=1)
    N, H, W = 40000, 64, 512
    torch.manual_seed(1234)
    ref = ra.AdaptiveAugment(**kw)
    Gr, Cr = ref.sample_affine(N, H, W), ref.sample_color(N)
    mine = Mine(**kw)
    mine.generator = torch.Generator().manual_seed(4321)
    Gm, Cm = mine.sample_affine(N, H, W), mine.sample_color(N)
    for a, b, name in ((Gr, Gm, "affine"), (Cr, Cm, "color")):
        assert a.shape == b.shape, name
        a, b = a.double(), b.double()
        se = (a.std(0) + b.std(0)) / np.sqrt(N) + 1e-9
        assert bool(((a.mean(0) - b.mean(0)).abs() <= 5 * se + 1e-6).all()), (name, a.mean(0), b.mean(0))
        # standard deviations: relative agreement (heavy-tailed entries: log-normal scales)
        assert bool(((a.std(0) - b.std(0)).abs() <= 0.05 * (a.std(0) + b.std(0)) / 2 + 1e-6).all()), (
            name, a.std(0), b.std(0))


@pytest.mark.skipif(not os.path.exists(_EXT_UFD), reason="reference `upfirdn2d` extension not prebuilt")
def test_ada_padding_and_filter_bank_match_reference(rops):
    """`padding_for` (host ints, no device sync) against the reference's `get_padding`
    (adaptive_augment.py:271-291) on sampled transforms, and the wavelet filter bank buffer."""
    from gans.augment import adaptive_augment as ra
    from dusty_gan_v2_b200.gans.augment import adaptive_augment as ma
    kw = dict(p_init=0.9, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1, frac_trans=1, brightness=1, contrast=1,
              luma_flip=1, hue=1, saturation=1)
    torch.manual_seed(77)
    ref = ra.AdaptiveAugment(**kw)
    for H, W, n in ((64, 512, 64), (16, 64, 8), (32, 64, 128)):
        G_inv = torch.inverse(ref.sample_affine(n, H, W))
        want = tuple(int(v) for v in ra.get_padding(G_inv, H, W, 12))
        assert tuple(int(v) for v in ma.padding_for(G_inv, H, W, 12)) == want
    mine = ma.AdaptiveAugment(**kw)
    assert torch.allclose(mine.Hz_fbank, ref.Hz_fbank, rtol=0, atol=1e-7)
    assert sorted(mine.state_dict().keys()) == sorted(ref.state_dict().keys())


@pytest.mark.skipif(not os.path.exists(_EXT_UFD), reason="reference `upfirdn2d` extension not prebuilt")
def test_trainer_host_methods_match_reference(rops, tmp_path):
    """Host-side pieces of the mirror Trainer against the reference's own methods executed on the
    CPU-assembled reference trainer: warm-up schedule (trainer.py:219-232), `fetch_reals`
    (211-217), the warm-up dropout (234-245, same Bernoulli draw on both sides), `sample_z`."""
    import torch.distributed as dist
    from ref_trainer_harness import build_reference_trainer
    from small_cfgs import D_SMALL, G_SMALL
    from dusty_gan_v2_b200.config import to_attr
    from dusty_gan_v2_b200.gans.trainer import Trainer
    from dusty_gan_v2_b200.presets import preset
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        B, H, W = 4, 16, 64
        R, _, _ = build_reference_trainer(G_SMALL, D_SMALL, B, (H, W), [], p_init=0.0)
    finally:
        if own_group:
            dist.destroy_process_group()
    cfg = preset("dusty_v2", batch_size=B)
    cfg.model.generator, cfg.model.discriminator = to_attr(G_SMALL), to_attr(D_SMALL)
    M = Trainer(cfg, iter([]), device="cpu", precision="fp32",
                angle_file=os.path.join(ref_import.REFERENCE_ROOT, "data/coords/kitti_raw.npy"))
    assert torch.equal(M.coord.angle, R.coord.angle)
    for it in (0, 1, 999, 12500, 25000, 49999, 50000, 80000):
        R.set_warmup_params(it)
        M.set_warmup_params(it)
        assert float(M.blur_sigma) == pytest.approx(float(R.blur_sigma), abs=0)
        assert float(M.dropout_ratio) == pytest.approx(float(R.dropout_ratio), abs=1e-15), it
    g = torch.Generator().manual_seed(5)
    batch = {"depth": 90.0 * torch.rand(B, 1, H, W, generator=g), "mask": (torch.rand(B, 1, H, W, generator=g) < 0.8).float()}
    r, m = R.fetch_reals(batch), M.fetch_reals(batch)
    assert torch.equal(m["image"], r["image"]) and torch.equal(m["raydrop_mask"], r["raydrop_mask"])
    for it in (0, 30000):
        R.set_warmup_params(it)
        M.set_warmup_params(it)
        torch.manual_seed(9)
        xr = R.warmup(r["image"].clone())
        torch.manual_seed(9)
        xm = M.warmup(m["image"].clone())
        assert torch.equal(xm, xr)
    torch.manual_seed(3)
    zr = R.sample_z(B)
    torch.manual_seed(3)
    assert torch.equal(M.sample_z(B), zr)
    # ema_inplace (trainer.py:28-41): parameters lerp, buffers copy
    from gans import trainer as rtr
    from dusty_gan_v2_b200.gans import trainer as mtr
    import copy
    src = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.BatchNorm1d(4))
    with torch.no_grad():
        src[1].running_mean.normal_()
    a, b = copy.deepcopy(src), copy.deepcopy(src)
    with torch.no_grad():
        for p_ in list(a.parameters()) + list(a.buffers()):
            if p_.is_floating_point():
                p_.add_(1.0)
        b.load_state_dict(a.state_dict())
    rtr.ema_inplace(a, src, 0.9)
    mtr.ema_inplace(b, src, 0.9)
    for (k, va), vb in zip(a.state_dict().items(), b.state_dict().values()):
        assert torch.allclose(va.float(), vb.float(), rtol=1e-6, atol=1e-7), k
    # optimiser settings the mirror derives from the config (trainer.py:137-171: lazy regularisation)
    lazy = 16 / 17.0
    assert M.optim_D.param_groups[0]["lr"] == pytest.approx(0.002 * lazy, rel=1e-12)
    assert M.optim_D.param_groups[0]["betas"] == pytest.approx((0.0, 0.99 ** lazy), rel=1e-12)
    assert M.optim_G.param_groups[0]["lr"] == pytest.approx(0.002) and M.gp_weight == 16 and M.gp_every == 16


def test_gan_loss_all_objectives_match_reference(rops):
    """gans/models/loss.py mirror: every objective the reference implements (loss.py:37-88), both
    modes, values and gradients."""
    from gans.models.loss import GANLoss as Ref
    from dusty_gan_v2_b200.gans.models.loss import GANLoss as Mine
    g = torch.Generator().manual_seed(42)
    for metric in Mine.METRICS:
        for smoothing in (1.0, 0.9):
            ref, mine = Ref(metric, smoothing), Mine(metric, smoothing)
            for mode in ("G", "D"):
                r0, f0 = torch.randn(6, 1, generator=g), torch.randn(6, 1, generator=g)
                outs = []
                for crit in (ref, mine):
                    r, f = r0.clone().requires_grad_(), f0.clone().requires_grad_()
                    loss = crit(r, f, mode)
                    grads = torch.autograd.grad(loss, [r, f], allow_unused=True)
                    outs.append((loss.detach(), grads))
                assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=1e-7), (metric, mode)
                for ga, gb in zip(outs[0][1], outs[1][1]):
                    assert (ga is None) == (gb is None), (metric, mode)
                    if ga is not None:
                        assert torch.allclose(ga, gb, rtol=1e-6, atol=1e-7), (metric, mode)


@pytest.mark.parametrize("arch", ["dusty_v2", "dusty_v1", "vanilla"])
def test_presets_equal_reference_yaml(arch):
    """`presets.preset(arch)` against the reference's own `configs/gans/<arch>.yaml` read by the
    mirror's `load_config`: every model / training / dataset value on the hot path is identical;
    only out-of-scope keys (dataset paths and splits, checkpoint cadence, validation) are absent."""
    from dusty_gan_v2_b200 import load_config
    from dusty_gan_v2_b200.presets import preset

    def plain(x):
        if isinstance(x, dict):
            return {k: plain(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [plain(v) for v in x]
        return x

    y = plain(load_config(os.path.join(ref_import.REFERENCE_ROOT, "configs", "gans", f"{arch}.yaml")))
    p = plain(preset(arch, batch_size=y["training"]["batch_size"],
                     resolution=tuple(y["model"]["generator"]["synthesis_kwargs"]["resolution"])))

    def diff(a, b, path=""):
        if isinstance(a, dict) and isinstance(b, dict):
            out = []
            for k in sorted(set(a) | set(b)):
                if k not in a or k not in b:
                    out.append(f"{path}/{k}")
                else:
                    out += diff(a[k], b[k], f"{path}/{k}")
            return out
        return [] if a == b else [path]

    allowed = {"/dataset/flip", "/dataset/root", "/dataset/test", "/dataset/train", "/dataset/val",
               "/training/checkpoint", "/training/pin_memory", "/validation"}
    assert set(diff(y, p)) <= allowed, sorted(set(diff(y, p)) - allowed)


@pytest.mark.skipif(not os.path.exists(_EXT_UFD), reason="reference `upfirdn2d` extension not prebuilt")
def test_checkpoint_written_by_reference_resumes_in_mirror(rops, tmp_path):
    """A checkpoint written by the reference's own `Trainer.save_checkpoint` (trainer.py:551-567)
    after one real training iteration resumes in the mirror Trainer: weights, EMA copy, ADA state
    and both Adam states land on the parameters of the same name (the optimiser state is indexed
    by parameter ORDER, so this also pins the module registration order); and the payload the
    mirror writes loads back into the reference's modules."""
    import pathlib
    import torch.distributed as dist
    from ref_trainer_harness import build_reference_trainer
    from small_cfgs import D_SMALL, G_SMALL
    from dusty_gan_v2_b200.config import to_attr
    from dusty_gan_v2_b200.gans.trainer import Trainer
    from dusty_gan_v2_b200.presets import preset
    B, H, W = 4, 16, 64
    g = torch.Generator().manual_seed(21)
    batch = {"depth": 1.45 + 78.55 * torch.rand(B, 1, H, W, generator=g),
             "mask": (torch.rand(B, 1, H, W, generator=g) < 0.85).float()}
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        torch.manual_seed(22)
        np.random.seed(22)
        R, RG, RD = build_reference_trainer(G_SMALL, D_SMALL, B, (H, W), [batch], p_init=0.3)
        R.step(0)
        path = pathlib.Path(tmp_path) / "ckpt" / "0000000004.pth"
        R.save_checkpoint(path, step=3 * B)
    finally:
        if own_group:
            dist.destroy_process_group()
    state = torch.load(path, map_location="cpu", weights_only=False)
    cfg = preset("dusty_v2", batch_size=B)
    cfg.model.generator, cfg.model.discriminator = to_attr(G_SMALL), to_attr(D_SMALL)
    M = Trainer(cfg, iter([]), device="cpu", precision="fp32",
                angle_file=os.path.join(ref_import.REFERENCE_ROOT, "data/coords/kitti_raw.npy"))
    assert [n for n, _ in M.G_module.named_parameters()] == [n for n, _ in RG.named_parameters()]
    assert [n for n, _ in M.D_module.named_parameters()] == [n for n, _ in RD.named_parameters()]
    assert M.load_state_dict(state) == 3
    for mine, ref in ((M.G_module, RG), (M.D_module, RD), (M.G_ema, R.G_ema), (M.A, R.A)):
        rs = ref.state_dict()
        ms = mine.state_dict()
        assert list(ms.keys()) == list(rs.keys())
        for k in rs:
            assert torch.equal(ms[k].float().reshape(-1), rs[k].float().reshape(-1)), k
    for m_opt, m_net, r_opt, r_net in ((M.optim_G, M.G_module, R.optim_G, RG), (M.optim_D, M.D_module, R.optim_D, RD)):
        r_by_name = dict(r_net.named_parameters())
        n = 0
        for name, p in m_net.named_parameters():
            rst = r_opt.state.get(r_by_name[name])
            if not rst:
                continue
            mst = m_opt.state[p]
            assert torch.equal(mst["exp_avg"], rst["exp_avg"]) and torch.equal(mst["exp_avg_sq"], rst["exp_avg_sq"]), name
            assert float(mst["step"]) == float(rst["step"])
            n += 1
        assert n >= 10
        assert m_opt.param_groups[0]["lr"] == r_opt.param_groups[0]["lr"]
        assert tuple(m_opt.param_groups[0]["betas"]) == tuple(r_opt.param_groups[0]["betas"])
    # and back: the mirror's payload loads into the reference's modules
    back = M.state_dict(step=3 * B)
    assert set(back.keys()) == set(state.keys())
    RG.load_state_dict(back["G"], strict=True)
    RD.load_state_dict(back["D"], strict=True)
    R.G_ema.load_state_dict(back["G_ema"], strict=True)
    # (the reference's own `p` buffer turns from 0-dim into shape [1] at its first update_p, so a
    # running reference trainer cannot even reload its own initial state; a fresh one takes ours)
    from gans.augment.adaptive_augment import AdaptiveAugment as RefADA
    RefADA(p_init=0.3, p_target=0.6, kimg=500, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1, frac_trans=1,
           brightness=1, contrast=1, luma_flip=1, hue=1, saturation=1).load_state_dict(back["A"])
    R.optim_G.load_state_dict(back["optim_G"])
    R.optim_D.load_state_dict(back["optim_D"])


def test_mapping_network_style_mixing_truncation_match_reference(rops):
    """base.Generator host-side paths that run without a kernel (base.py:26-114): mapping network,
    style mixing (same `random.randint` / `torch.randn_like` consumption), truncation trick and the
    w_avg moving average -- mirror against the reference with the same weights and seeds."""
    import random
    from gans.models.builder import build_generator as ref_build
    from small_cfgs import G_SMALL
    from dusty_gan_v2_b200.gans.models.builder import build_generator as my_build
    torch.manual_seed(7)
    np.random.seed(7)
    R = ref_build(ref_import.to_attr(G_SMALL))
    M = my_build(G_SMALL)
    M.load_state_dict(R.state_dict(), strict=True)
    z = torch.randn(5, 16)
    assert torch.allclose(M.mapping_network(z), R.mapping_network(z), rtol=1e-6, atol=1e-7)
    for seed in (0, 1, 2):
        outs = []
        for net in (R, M):
            random.seed(seed)
            torch.manual_seed(seed)
            outs.append(net.forward_mapping(z, style_mixing=True))
        assert outs[0].shape == outs[1].shape and torch.allclose(outs[0], outs[1], rtol=1e-6, atol=1e-7)
    w = R.forward_mapping(z, False)
    with torch.no_grad():
        R.w_avg.normal_()
        M.w_avg.copy_(R.w_avg)
    assert torch.allclose(M.truncation_trick(w, 0.6), R.truncation_trick(w, 0.6), rtol=1e-6, atol=1e-7)
    R.moving_average_w(w)
    M.moving_average_w(w)
    assert torch.allclose(M.w_avg, R.w_avg, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("demod,ema", [(True, True), (False, True), (True, False)])
def test_modconv_composite_weights_match_reference_forward(rops, demod, ema):
    """`ModConv2d._effective_weights_composite` (the plain-tensor statement of the fused
    dusty_modprep kernel, which the GPU tests hold that kernel to) against the reference module:
    bmm(wb, x) equals the reference's grouped-convolution forward (style.py:68-126)."""
    from dusty_gan_v2_b200.gans.models.ops.style import ModConv2d as Mine
    torch.manual_seed(11)
    C, Oc, Mch = 12, 7, 16
    ref = rops.ModConv2d(in_ch=C, out_ch=Oc, mod_ch=Mch, ksize=1, stride=1, padding=0, demod=demod, bias=False,
                         ema=ema).eval()
    mine = Mine(in_ch=C, out_ch=Oc, mod_ch=Mch, ksize=1, stride=1, padding=0, demod=demod, bias=False, ema=ema)
    mine.load_state_dict(ref.state_dict(), strict=True)
    with torch.no_grad():
        ref.ema_var.fill_(0.8)
        mine.ema_var.fill_(0.8)
    x, style = torch.randn(3, C, 4, 8), torch.randn(3, Mch)
    wb = mine._effective_weights_composite(mine.mod(style))
    y = torch.bmm(wb, x.flatten(2)).reshape(3, Oc, 4, 8)
    assert torch.allclose(y, ref(x, style), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("unfold", [True, False])
def test_kitti_scan_projection_matches_reference_live(rops, tmp_path, unfold):
    """f4: the oracle's scan -> range image (and the mirror's host-side cell computation) against
    the reference's own `KITTIRaw.load_pts_as_img` on a fresh synthetic scan."""
    from gans.datasets.kitti import KITTIRaw
    from dusty_gan_v2_b200.gans.datasets.kitti import scan_cells
    from small_cfgs import synthetic_scan
    pts = synthetic_scan(rings=70, per_ring=90, seed=11)
    f = tmp_path / "scan.bin"
    pts.tofile(str(f))
    ds = object.__new__(KITTIRaw)
    ds.min_depth, ds.max_depth = 1.45, 80.0
    ref = ds.load_pts_as_img(str(f), unfold, 64, 2048).transpose(2, 0, 1)
    ref = ref * ref[5:6]
    assert np.array_equal(O.scan_to_image(pts, 64, 2048, 2048, 1.45, 80.0, unfold), ref)
    a, b = scan_cells(pts, 64, 2048, unfold), O.scan_cells(pts, 64, 2048, unfold)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
