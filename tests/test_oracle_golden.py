"""Pins oracle/dusty_oracle.py against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import dusty_oracle as O

T = torch.from_numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else a
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_bias_act(g_ops):
    x, b = T(g_ops["ba_x"]), T(g_ops["ba_b"])
    y = O.bias_act(x, b)
    close(y, g_ops["ba_y"])
    dx, db = O.bias_act_grad(T(g_ops["ba_dy"]), y)
    close(dx, g_ops["ba_dx"])
    close(db, g_ops["ba_db"], rtol=1e-5, atol=1e-5)
    close(O.bias_act(T(g_ops["ba2_x"]), T(g_ops["ba2_b"])), g_ops["ba2_y"])


@pytest.mark.parametrize("name", ["ada_upx", "ada_upy", "ada_dnx", "ada_dny", "gen2d", "ident"])
def test_upfirdn2d(g_ops, name):
    cfg = g_ops[f"ufd_{name}_cfg"].tolist()
    y = O.upfirdn2d(T(g_ops[f"ufd_{name}_x"]), T(g_ops[f"ufd_{name}_k"]), up=tuple(cfg[0:2]),
                    down=tuple(cfg[2:4]), pad=tuple(cfg[4:8]))
    assert tuple(y.shape) == g_ops[f"ufd_{name}_y"].shape
    close(y, g_ops[f"ufd_{name}_y"], rtol=1e-5, atol=1e-6)


def test_resample_family(g_ops):
    x = T(g_ops["rs_x"])
    close(O.resample(x, up=2), g_ops["rs_up2"])
    close(O.resample(x, down=2), g_ops["rs_down2"])
    close(O.resample(x), g_ops["rs_blur4"])
    close(O.resample(x, window=(1, 2, 1), direction="h"), g_ops["rs_blur3_h"])
    close(O.resample(x, window=(1, 2, 1), direction="w"), g_ops["rs_blur3_w"])
    close(O.resample(x, up=2, ring=False), g_ops["rs_up2_noring"])
    close(O.blur_vh(x), g_ops["rs_blurvh"])
    close(O.pad2d(x, 1, ring=True), g_ops["rs_pad1"])
    close(O.pad2d(x, 1, ring=True, mode="reflect"), g_ops["rs_pad1_reflect"])
    close(O.filter2d(x, T(g_ops["rs_filter2d_k"])), g_ops["rs_filter2d"])


@pytest.mark.parametrize("nm,kw", [("up2", dict(up=2)), ("down2", dict(down=2)), ("blur4", {})])
def test_resample_grad(g_ops, nm, kw):
    x = T(g_ops["rs_x"]).clone().requires_grad_()
    y = O.resample(x, **kw)
    (gx,) = torch.autograd.grad(y, x, T(g_ops[f"rs_{nm}_gy"]))
    close(gx, g_ops[f"rs_{nm}_gx"], rtol=1e-5, atol=1e-6)


def test_fourier(g_ops):
    out = O.fourier_feature(T(g_ops["ff_angle"]), T(g_ops["ff_freqs"]), T(g_ops["ff_phase"]))
    # |arg| ~ 1e2: one fp32 ulp of the argument is ~1e-5
    close(out, g_ops["ff_out"], rtol=0, atol=3e-5)


@pytest.mark.parametrize("tag,demod", [("dm", True), ("hd", False)])
def test_modconv(g_ops, tag, demod):
    g = lambda k: T(g_ops[f"mc_{tag}_{k}"])
    sd = {k[len(f"mc_{tag}_sd_"):]: T(v) for k, v in g_ops.items() if k.startswith(f"mc_{tag}_sd_")}
    x = g("x").clone().requires_grad_()
    st = g("style").clone().requires_grad_()
    w = sd["weight"].clone().requires_grad_()
    mw = sd["mod.module.weight"].clone().requires_grad_()
    mb = sd["mod.module.bias"].clone().requires_grad_()
    bias = sd.get("bias")
    y, ev = O.modconv(x, st, w, mw, mb, sd["ema_var"], demod=demod, bias=bias, training=False)
    close(y, g_ops[f"mc_{tag}_y_eval"], rtol=1e-4, atol=1e-5)
    assert float(ev) == float(sd["ema_var"])
    grads = torch.autograd.grad(y, [x, st, w, mw, mb], g("gy"))
    for nm, gr in zip(("gx", "gstyle", "gw", "gmodw", "gmodb"), grads):
        ref = g_ops[f"mc_{tag}_{nm}"]
        close(gr, ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max())
    y2, ev2 = O.modconv(x, st, w, mw, mb, sd["ema_var"], demod=demod, bias=bias, training=True)
    close(y2, g_ops[f"mc_{tag}_y_train"], rtol=1e-4, atol=1e-5)
    close(ev2, g_ops[f"mc_{tag}_ema_after"], rtol=1e-6, atol=0)


def test_gumbel_raydrop(g_ops):
    logit = T(g_ops["gs_logit"]).clone().requires_grad_()
    img = T(g_ops["gs_img"]).clone().requires_grad_()
    o = O.raydrop(img, logit, T(g_ops["gs_u"]))
    # mask and count are integer-exact
    assert np.array_equal(o["raydrop_mask"].detach().numpy(), g_ops["gs_mask"])
    assert int(o["raydrop_mask"].sum().item()) == int(g_ops["gs_count"])
    assert np.array_equal(o["image"].detach().numpy(), g_ops["gs_image"])
    gl, gi = torch.autograd.grad(o["image"], [logit, img], T(g_ops["gs_gout"]))
    close(gl, g_ops["gs_glogit"], rtol=1e-5, atol=1e-7)
    close(gi, g_ops["gs_gimg"], rtol=1e-6, atol=0)


def test_minibatch_std_pixelnorm(g_ops):
    close(O.minibatch_stddev(T(g_ops["mb_x"])), g_ops["mb_y"], rtol=1e-5, atol=1e-6)
    close(O.minibatch_stddev(T(g_ops["mb2_x"])), g_ops["mb2_y"], rtol=1e-5, atol=1e-6)
    close(O.pixel_norm(T(g_ops["pn_x"])), g_ops["pn_y"], rtol=1e-6, atol=1e-7)


def test_coords(g_coords):
    raw = np.load(os.path.join(ROOT, "data", "coords", "kitti_raw.npy"))
    assert hashlib.sha256(raw.tobytes()).hexdigest() == str(g_coords["raw_sha256"])
    angle = O.angle_grid(raw, 64, 512)
    # bit-exact: same ATen CPU ops in the same order
    assert np.array_equal(angle.numpy(), g_coords["angle"])
    xin = T(g_coords["xin"])
    pm, ps, count = O.inv_depth_norm_to_points(xin, angle, 1.45, 80.0)
    assert count == int(g_coords["valid_count"])
    assert np.array_equal(pm.numpy()[:, :, ::4, ::8], g_coords["point_map"])
    # pixel indexing: point_set[b, h*W + w] == point_map[b, :, h, w]
    assert np.array_equal(ps.numpy()[:, :2048], g_coords["point_set_head"])
    depth, _ = O.inv_depth_norm_to_depth(xin, 1.45, 80.0)
    assert np.array_equal((depth / 80.0).numpy()[:, :, ::4, ::8], g_coords["depth_norm"])
    reals = O.fetch_reals(T(g_coords["depth"]), T(g_coords["mask"]), 1.45, 80.0)
    assert np.array_equal(reals.numpy(), g_coords["reals"])


def _sd(g, prefix="sd_"):
    return {k[len(prefix):]: T(v) for k, v in g.items() if k.startswith(prefix)}


def test_generator_eval(g_gen):
    sd = _sd(g_gen)
    z, angle = T(g_gen["z"]), T(g_gen["angle"])
    for tag, psi in (("eval", 1.0), ("psi", 0.7)):
        o = O.generator(sd, z, angle, T(g_gen[f"{tag}_u"]), training=False, truncation_psi=psi)
        close(o["w"][:, 0], g_gen[f"{tag}_w0"], rtol=1e-5, atol=1e-6)
        for k in ("image_orig", "raydrop_logit"):
            ref = g_gen[f"{tag}_{k}"]
            close(o[k], ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max())
        # the hard mask may only differ where logit+noise is within float noise of 0
        assert (o["raydrop_mask"].numpy() != g_gen[f"{tag}_raydrop_mask"]).mean() < 1e-3


def test_generator_train_and_grads(g_gen):
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and
                                       not (k.endswith("ema_var") or k == "w_avg"
                                            or "kernel" in k or "pe." in k or "raydrop_const" in k))
          for k, v in _sd(g_gen).items()}
    z, angle = T(g_gen["z"]), T(g_gen["angle"])
    newb = {}
    o = O.generator(sd, z, angle, T(g_gen["train_u"]), training=True,
                    shifts_rad=T(g_gen["train_shift01"]) * (2 * np.pi), new_buffers=newb)
    for k in ("image_orig", "raydrop_logit"):
        ref = g_gen[f"train_{k}"]
        close(o[k], ref, rtol=1e-3, atol=2e-4 * np.abs(ref).max())
    for k, v in newb.items():
        close(v, g_gen[f"after_{k}"], rtol=1e-5, atol=1e-7)
    loss = (o["image"] * T(g_gen["train_gi"])).sum() + (o["raydrop_logit"] * T(g_gen["train_gl"])).sum()
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    checked = 0
    for k, g in zip(names, grads):
        key = f"grad_{k}"
        if key in g_gen and g is not None:
            ref = g_gen[key]
            close(g, ref, rtol=2e-3, atol=2e-3 * max(np.abs(ref).max(), 1e-6))
            checked += 1
    assert checked > 40


def test_discriminator_first_and_second_order(g_disc):
    sd = {k: v.clone().requires_grad_("kernel" not in k) for k, v in _sd(g_disc).items()}
    x = T(g_disc["x"]).clone().requires_grad_()
    y = O.discriminator(sd, x)
    close(y, g_disc["y"], rtol=1e-4, atol=1e-5)
    names = [k for k, v in sd.items() if v.requires_grad]
    loss = O.nsgan_g(y)
    g1 = torch.autograd.grad(loss, [x] + [sd[k] for k in names], retain_graph=True)
    close(g1[0], g_disc["gx_loss"], rtol=1e-3, atol=1e-6)
    for k, g in zip(names, g1[1:]):
        ref = g_disc[f"g1_{k}"]
        close(g, ref, rtol=1e-3, atol=1e-4 * max(np.abs(ref).max(), 1e-8))
    (gx,) = torch.autograd.grad(y.sum(), x, create_graph=True)
    close(gx, g_disc["r1_gx"], rtol=1e-3, atol=1e-6)
    r1 = O.r1_penalty(gx)
    close(r1, g_disc["r1"], rtol=1e-4, atol=0)
    g2 = torch.autograd.grad(r1, [sd[k] for k in names], allow_unused=True)
    n = 0
    for k, g in zip(names, g2):
        if f"g2_{k}" in g_disc and g is not None:
            ref = g_disc[f"g2_{k}"]
            close(g, ref, rtol=2e-3, atol=2e-4 * max(np.abs(ref).max(), 1e-8))
            n += 1
    assert n >= 8


def test_ada_apply(g_ada):
    x = T(g_ada["x"]).clone().requires_grad_()
    y = O.ada_apply(x, T(g_ada["G_inv"]), T(g_ada["C"]))
    close(y, g_ada["y"], rtol=1e-4, atol=2e-5)
    (gx,) = torch.autograd.grad(y, x, T(g_ada["gy"]))
    close(gx, g_ada["gx"], rtol=1e-4, atol=2e-5)


def test_vanilla_and_dusty_v1(g_vanilla):
    sdG = {k[4:]: T(v).clone().requires_grad_(v.dtype.kind == "f" and "kernel" not in k and "raydrop_const" not in k and "w_avg" not in k)
           for k, v in g_vanilla.items() if k.startswith("sdG_")}
    sdD = {k[4:]: T(v).clone().requires_grad_("kernel" not in k) for k, v in g_vanilla.items()
           if k.startswith("sdD_")}
    z = T(g_vanilla["z"]).clone().requires_grad_()
    o = O.vanilla_generator(sdG, z, T(g_vanilla["u"]))
    for k in ("image_orig", "raydrop_logit"):
        close(o[k], g_vanilla[k], rtol=1e-4, atol=1e-5)
    assert np.array_equal(o["raydrop_mask"].detach().numpy(), g_vanilla["raydrop_mask"])
    y = O.vanilla_discriminator(sdD, o["image"])
    close(y, g_vanilla["y"], rtol=1e-4, atol=1e-5)
    loss = O.nsgan_g(y)
    namesG = [k for k, v in sdG.items() if v.requires_grad]
    namesD = [k for k, v in sdD.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [z] + [sdG[k] for k in namesG] + [sdD[k] for k in namesD],
                                allow_unused=True)
    close(grads[0], g_vanilla["gz"], rtol=1e-3, atol=1e-7)
    n = 0
    for k, g in zip(namesG, grads[1:1 + len(namesG)]):
        if f"gG_{k}" in g_vanilla and g is not None:
            ref = g_vanilla[f"gG_{k}"]
            close(g, ref, rtol=2e-3, atol=1e-4 * max(np.abs(ref).max(), 1e-8))
            n += 1
    for k, g in zip(namesD, grads[1 + len(namesG):]):
        if f"gD_{k}" in g_vanilla and g is not None:
            ref = g_vanilla[f"gD_{k}"]
            close(g, ref, rtol=2e-3, atol=1e-4 * max(np.abs(ref).max(), 1e-8))
            n += 1
    assert n >= 15


# ----------------------------------------------------------------------------- config 5
@pytest.mark.parametrize("tag,kw", [("l1_rel_full", dict(level=None, loss="l1", relative=True)),
                                    ("l1_rel_l2", dict(level=2, loss="l1", relative=True)),
                                    ("l2_abs_l3", dict(level=3, loss="l2", relative=False))])
def test_inversion_multiscale_masked_loss(g_inv, tag, kw):
    gen = T(g_inv["gen"]).requires_grad_()
    loss = O.multiscale_masked_loss(gen, T(g_inv["ref"]), T(g_inv["mask"]), **kw)
    close(loss, g_inv[f"{tag}_loss"], rtol=1e-5, atol=1e-6)
    (g,) = torch.autograd.grad(loss.sum(), gen)
    close(g, g_inv[f"{tag}_grad"], rtol=1e-4, atol=1e-6)


def test_inversion_geocross_and_spherical_projection(g_inv):
    lat = T(g_inv["lat"]).requires_grad_()
    gl = O.geocross_loss(lat)
    close(gl, g_inv["geocross"], rtol=1e-5, atol=1e-7)
    (g,) = torch.autograd.grad(gl.sum(), lat)
    close(g, g_inv["geocross_grad"], rtol=1e-4, atol=1e-7)
    # Adam (lr 0.1, betas 0.9 / 0.999) + projection to unit RMS, two steps with fixed gradients
    p = torch.nn.Parameter(T(g_inv["sph_p0"]).clone())
    opt = torch.optim.Adam([p], lr=0.1, betas=(0.9, 0.999))
    for i in range(2):
        p.grad = T(g_inv[f"sph_g{i}"]).clone()
        opt.step()
        O.spherical_project_(p.data)
        close(p, g_inv[f"sph_p{i + 1}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("latent_type", ["w", "w+"])
def test_inversion_loop_against_reference(g_gen, g_invloop, latent_type):
    """demo_inversion.py's stage-1 loop (targets, latent init, objective, Adam + LambdaLR):
    three steps of the oracle against the trajectory recorded from the reference's parts."""
    g, sd = g_invloop, _sd(g_gen)
    MIN_D, MAX_D, STEPS = 1.45, 80.0, 3
    depth, mask, angle = T(g["depth"]), T(g["mask"]), T(g["angle"])
    close(O.angle_grid(np.load(os.path.join(ROOT, "data/coords/kitti_raw.npy")), 16, 64), g["angle"],
          rtol=0, atol=0)
    t_depth, t_inv = O.inversion_targets(depth, mask, MIN_D, MAX_D)
    close(t_depth, g["t_depth"], rtol=0, atol=0)
    close(t_inv, g["t_inv_depth"], rtol=1e-6, atol=1e-8)
    # latent initialisation (demo_inversion.py:99-107)
    torch.manual_seed(0)
    ws = O.mapping_network(sd, torch.randn(256, 16))
    z_avg = ws.mean(dim=0, keepdim=True)
    close(z_avg, g["z_avg"], rtol=1e-5, atol=1e-6)
    close((((ws - z_avg) ** 2).sum() / 256).sqrt(), g["z_std"], rtol=1e-5)
    z = torch.nn.Parameter(T(g[f"{latent_type}_z0"]).clone())
    opt = torch.optim.Adam([z], lr=5e-2)
    for step in range(STEPS):
        lr = 5e-2 * O.inversion_lr_schedule(step, STEPS)
        assert abs(lr - float(g[f"{latent_type}_lr{step}"])) < 1e-12
        for grp in opt.param_groups:
            grp["lr"] = lr
        out, loss = O.inversion_forward(sd, z, angle, t_depth, t_inv, mask, MIN_D, MAX_D, latent_type)
        opt.zero_grad(set_to_none=True)
        loss.backward(gradient=torch.ones_like(loss))
        if step == 0:
            close(out["inv_depth_orig"], g[f"{latent_type}_inv_depth_orig0"], rtol=1e-4, atol=1e-6)
            close(out["g_depth"], g[f"{latent_type}_g_depth0"], rtol=1e-4, atol=1e-6)
            close(out["raydrop_logit"], g[f"{latent_type}_raydrop_logit0"], rtol=1e-4, atol=1e-5)
        ref_g = g[f"{latent_type}_grad{step}"]
        close(loss, g[f"{latent_type}_loss{step}"], rtol=1e-4, atol=1e-6)
        close(z.grad, ref_g, rtol=2e-3, atol=2e-4 * np.abs(ref_g).max())
        opt.step()
        close(z, g[f"{latent_type}_z{step + 1}"], rtol=1e-3, atol=2e-3)


def test_fir_axis_forms_agree(g_ops, g_gen):
    """The oracle's library-correlation FIR (what bench.py's CPU baseline legs run) against its
    literal tap-by-tap form (what the parity tests check against) over the whole Resample family
    (up / down / blur, both axes, ring and replicate boundaries), and the golden checks that
    lean on the FIR re-run under the library form."""
    prev = O.set_fir_impl("library")
    try:
        assert O._FIR_IMPL["mode"] == "library"
        test_resample_family(g_ops)
        test_generator_eval(g_gen)
    finally:
        O.set_fir_impl(prev)
    assert O._FIR_IMPL["mode"] == "gather"
    g = torch.Generator().manual_seed(5)
    for up, down, win, ring in ((2, 1, (1, 3, 3, 1), True), (1, 2, (1, 3, 3, 1), True), (1, 1, (1, 3, 3, 1), True),
                                (1, 1, (1, 2, 1), True), (2, 1, (1, 3, 3, 1), False), (2, 1, (1, 2, 1), True)):
        taps = torch.tensor(win, dtype=torch.float32)
        taps = taps / taps.sum()
        p0, p1 = O.resample_geometry(len(win), up, down)
        for H, W in ((4, 8), (6, 20), (16, 64)):
            x = torch.randn(2, 3, H, W, generator=g)
            for axis, circ in ((3, ring), (2, False)):
                O.set_fir_impl("library")
                a = O._fir_axis(x, taps, up, down, p0, p1, axis, circ)
                O.set_fir_impl("gather")
                b = O._fir_axis_gather(x, taps, up, down, p0, p1, axis, circ)
                assert a.shape == b.shape
                close(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("which", ["small", "mid"])
def test_train_iteration_against_reference_trainer_step(g_step, g_step_mid, which):
    """`O.train_iteration` (what bench.py times as the CPU baseline) against ONE FULL ITERATION of
    the reference's real `Trainer.step` (tests/golden/make_golden_trainer_step.py): the recorded
    random draws are replayed, the oracle advanced phase by phase with the reference's Adam
    settings -- G step, D step on the updated generator, lazy R1 step on the updated
    discriminator.  "mid": the fixture of the bf16 / CUDA-graph GPU twin (tcgen05-sized channel
    counts, gradients stored as samples + norms)."""
    from step_replay import (check, check_grads, check_updated_weights, fixture_draws, oracle_iteration,
                             oracle_rnd)
    g = g_step if which == "small" else g_step_mid
    sdG = {k[4:]: T(v).clone() for k, v in g.items() if k.startswith("sdG_")}
    sdD = {k[4:]: T(v).clone() for k, v in g.items() if k.startswith("sdD_")}
    x_real = O.fetch_reals(T(g["depth"]), T(g["mask"]), 1.45, 80.0)
    r = oracle_iteration(O, sdG, sdD, x_real, T(g["angle"]), oracle_rnd(fixture_draws(g)))
    close(r["loss_G"], g["loss_G"], rtol=1e-4, atol=1e-6)
    close(r["loss_D"], g["loss_D"], rtol=2e-3, atol=1e-5)
    close(r["r1"], g["r1"], rtol=5e-3, atol=1e-7)
    check_grads(r["grads_G"], g, "gG_", 5e-3, 2e-3, 20)
    check_grads(r["grads_D"], g, "gD_", 5e-3, 2e-3, 10)
    check_grads(r["grads_R1"], g, "gR1_", 5e-3, 2e-3, 10)
    check_updated_weights({k: v for k, v in r["sdG"].items() if v.requires_grad}, g, "afterG_", 0.002)
    check_updated_weights({k: v for k, v in r["sdD"].items() if v.requires_grad}, g, "afterD_", 0.004,
                          min_total=500)


@pytest.mark.parametrize("arch", ["dusty_v1", "vanilla"])
def test_train_iteration_config3_against_reference_trainer_step(arch):
    """BASELINE config 3: `O.train_iteration(arch=...)` against one full iteration of the reference's
    real `Trainer.step` with the dusty_v1 / vanilla generator and the vanilla discriminator."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"trainer_step_{arch}.npz")))
    sdG = {k[4:]: T(v).clone() for k, v in g.items() if k.startswith("sdG_")}
    sdD = {k[4:]: T(v).clone() for k, v in g.items() if k.startswith("sdD_")}
    sdG = {k: v.requires_grad_(v.dtype.is_floating_point and not any(t in k for t in ("kernel", "raydrop_const", "w_avg")))
           for k, v in sdG.items()}
    sdD = {k: v.requires_grad_("kernel" not in k) for k, v in sdD.items()}
    rnd = {k: T(g[k]) for k in ("z_g", "z_d", "u_g", "u_d") if k in g}
    rnd["shift_g"] = rnd["shift_d"] = torch.zeros(4)
    if arch == "vanilla":
        assert "u_g" not in g                        # no Gumbel draw: the vanilla generator has no raydrop head
        rnd["u_g"] = rnd["u_d"] = None
    for tag in ("g_fake", "d_real", "d_fake", "r1"):
        rnd[f"keep_{tag}"] = T(g[f"keep_{tag}"])
        rnd[f"Ginv_{tag}"] = torch.inverse(T(g[f"G_{tag}"]))
        rnd[f"C_{tag}"] = T(g[f"C_{tag}"])
    x_real = O.fetch_reals(T(g["depth"]), T(g["mask"]), 1.45, 80.0)
    lazy = 16 / 17.0
    optG = torch.optim.Adam([v for v in sdG.values() if v.requires_grad], lr=0.002, betas=(0.0, 0.99))
    optD = torch.optim.Adam([v for v in sdD.values() if v.requires_grad], lr=0.002 * lazy, betas=(0.0, 0.99 ** lazy))

    def apply(opt, sd, grads):
        for k, gr in grads.items():
            sd[k].grad = gr
        opt.step()
        opt.zero_grad(set_to_none=True)

    def check_grads(got, prefix, min_n):
        n = 0
        for k, gr in got.items():
            if gr is None or prefix + k not in g:
                continue
            ref = g[prefix + k]
            close(gr, ref, rtol=5e-3, atol=2e-3 * max(float(np.abs(ref).max()), 1e-7))
            n += 1
        assert n >= min_n, n

    r = O.train_iteration(sdG, sdD, x_real, None, rnd, with_r1=False, arch=arch)
    close(r["loss_G"], g["loss_G"], rtol=1e-4, atol=1e-6)
    check_grads(r["grads_G"], "gG_", 8)
    apply(optG, sdG, r["grads_G"])
    r = O.train_iteration(sdG, sdD, x_real, None, rnd, with_r1=False, arch=arch)
    close(r["loss_D"], g["loss_D"], rtol=2e-3, atol=1e-5)
    check_grads(r["grads_D"], "gD_", 8)
    apply(optD, sdD, r["grads_D"])
    r = O.train_iteration(sdG, sdD, x_real, None, rnd, with_r1=True, arch=arch)
    close(r["r1"], g["r1"], rtol=5e-3, atol=1e-7)
    check_grads(r["grads_R1"], "gR1_", 8)


@pytest.mark.parametrize("tag,unfold", [("unfold", True), ("elev", False)])
def test_kitti_scan_projection_golden(g_kitti, tag, unfold):
    """f4: the oracle's scan -> range image against the reference's `load_pts_as_img` + NEAREST
    resize + mask product (tests/golden/make_golden_kitti.py), bit for bit; and the mirror's
    loop-free cell computation (host side of `scan_to_image`) against the oracle's sequential
    ring walk, including the ring index -1 the reference assigns to the 65th ring from the top."""
    from dusty_gan_v2_b200.gans.datasets.kitti import scan_cells
    pts = g_kitti["points"]
    img = O.scan_to_image(pts, 64, 2048, 512, 1.45, 80.0, unfold)
    assert np.array_equal(img, g_kitti[f"{tag}_image"])
    a, b = scan_cells(pts, 64, 2048, unfold), O.scan_cells(pts, 64, 2048, unfold)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    if unfold:
        assert a[0].min() == -1 and a[0].max() == 63

