"""GPU parity tests: every kernel family, called through the host API (ctypes -> C ABI),
against the CPU oracle on identical seeded inputs and against the reference-generated
golden fixtures.  Tolerances: bit-exact for integer / index / mask-count results;
rtol 1e-3 (fp32) and 2e-2 (bf16) for floating point, as BASELINE.json states."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dusty_oracle as O  # noqa: E402

T = torch.from_numpy
DEV = "cuda"


def dev(a):
    return (T(a) if isinstance(a, np.ndarray) else a).to(DEV)


def close(a, b, rtol=1e-3, atol_rel=1e-4):
    a = a.detach().float().cpu().numpy()
    b = b.detach().float().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    atol = atol_rel * max(float(np.abs(b).max()), 1e-12)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def DF():
    import dusty_gan_v2_b200.functional as DF
    return DF


@pytest.fixture(scope="module")
def ops():
    from dusty_gan_v2_b200.gans.models import ops
    return ops


# ----------------------------------------------------------------------------- a3 bias_act
def test_bias_act_golden_and_grads(DF, g_ops):
    x = dev(g_ops["ba_x"]).requires_grad_()
    b = dev(g_ops["ba_b"]).requires_grad_()
    y = DF.bias_act(x, b)
    close(y, g_ops["ba_y"], rtol=1e-6, atol_rel=1e-7)
    dx, db = torch.autograd.grad(y, [x, b], dev(g_ops["ba_dy"]))
    close(dx, g_ops["ba_dx"], rtol=1e-6, atol_rel=1e-7)
    close(db, g_ops["ba_db"], rtol=1e-5, atol_rel=1e-6)
    close(DF.bias_act(dev(g_ops["ba2_x"]), dev(g_ops["ba2_b"])), g_ops["ba2_y"], 1e-6, 1e-7)


@pytest.mark.parametrize("shape", [(3, 5, 7, 9), (2, 32, 64, 512), (4, 6), (1, 3, 5), (2, 8, 4, 4)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bias_act_vs_oracle_double_backward(DF, shape, dtype):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    if dtype == torch.bfloat16:
        x, b = x.bfloat16().float(), b.bfloat16().float()
    xr, br = x.clone().requires_grad_(), b.clone().requires_grad_()
    yr = O.bias_act(xr, br)
    xg, bg = x.to(DEV, dtype).requires_grad_(), b.to(DEV).requires_grad_()
    yg = DF.bias_act(xg, bg)
    tol = dict(rtol=1e-3, atol_rel=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol_rel=1e-2)
    close(yg, yr, **tol)
    gy = torch.randn(shape, generator=g)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    (dxr, dbr) = torch.autograd.grad(yr, [xr, br], gy, create_graph=True)
    (dxg, dbg) = torch.autograd.grad(yg, [xg, bg], gy.to(DEV, dtype), create_graph=True)
    close(dxg, dxr, **tol)
    close(dbg, dbr, rtol=tol["rtol"], atol_rel=max(tol["atol_rel"], 1e-4))
    if dtype == torch.float32:
        # second order: d/d(gy) of sum(dx * v) -- the R1 path
        v = torch.randn(shape, generator=g)
        # the oracle's bias_act is piecewise linear: grad-grad flows through gy only
        gy_r = gy.clone().requires_grad_()
        dx2 = torch.where(yr.detach() > 0, gy_r, gy_r * 0.2) * O.SQRT2
        (ggr,) = torch.autograd.grad((dx2 * v).sum(), gy_r)
        gy_g = gy.to(DEV).requires_grad_()
        (dxg2,) = torch.autograd.grad(yg, xg, gy_g, create_graph=True)
        (ggg,) = torch.autograd.grad((dxg2 * v.to(DEV)).sum(), gy_g)
        close(ggg, ggr, rtol=1e-5, atol_rel=1e-6)


def test_fused_bias_act_native_contract(ops):
    from dusty_gan_v2_b200.gans.models.ops.fused_act.fused_act import fused
    x = torch.randn(2, 3, 4, device=DEV)
    empty = x.new_empty(0)
    y = fused.fused_bias_act(x, empty, empty, 3, 0, 0.2, 1.0)
    close(y, torch.nn.functional.leaky_relu(x, 0.2), 1e-6, 1e-7)
    y2 = fused.fused_bias_act(x, empty, y, 3, 1, 0.2, 2.0)
    close(y2, torch.where(y > 0, x, x * 0.2) * 2.0, 1e-6, 1e-7)
    assert float(fused.fused_bias_act(x, empty, empty, 3, 2, 0.2, 1.0).abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        fused.fused_bias_act(x.cpu(), empty.cpu(), empty.cpu(), 3, 0, 0.2, 1.0)
    with pytest.raises(RuntimeError):
        fused.fused_bias_act(x.transpose(0, 1), empty, empty, 3, 0, 0.2, 1.0)
    with pytest.raises(RuntimeError):
        ops.fused_leaky_relu(torch.randn(2, 3), torch.zeros(3))          # CPU: no fallback


# ----------------------------------------------------------------------------- a5 upfirdn2d
@pytest.mark.parametrize("name", ["ada_upx", "ada_upy", "ada_dnx", "ada_dny", "gen2d", "ident"])
def test_upfirdn2d_golden(g_ops, name):
    from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d, upfirdn2d_op
    cfg = g_ops[f"ufd_{name}_cfg"].tolist()
    x, k = dev(g_ops[f"ufd_{name}_x"]), dev(g_ops[f"ufd_{name}_k"])
    y = upfirdn2d(x, k, up=tuple(cfg[0:2]), down=tuple(cfg[2:4]), pad=tuple(cfg[4:8]))
    assert tuple(y.shape) == g_ops[f"ufd_{name}_y"].shape
    close(y, g_ops[f"ufd_{name}_y"], rtol=1e-4, atol_rel=1e-6)
    n, c, h, w = x.shape
    y2 = upfirdn2d_op.upfirdn2d(x.reshape(-1, h, w, 1), k, *cfg)
    close(y2.reshape(y.shape), g_ops[f"ufd_{name}_y"], rtol=1e-4, atol_rel=1e-6)


@pytest.mark.parametrize("cfg", [((2, 1), (1, 1), (6, 5, 0, 0), (1, 12)), ((1, 2), (1, 1), (0, 0, 6, 5), (12, 1)),
                                 ((1, 1), (2, 1), (-1, -1, 0, 0), (1, 12)), ((1, 1), (1, 2), (0, 0, -1, -1), (12, 1)),
                                 ((2, 2), (1, 1), (2, 1, 2, 1), (4, 4)), ((1, 1), (2, 2), (1, 1, 1, 1), (4, 4)),
                                 ((3, 2), (2, 3), (1, 4, -2, 3), (5, 3))])
def test_upfirdn2d_grad_and_gradgrad_vs_oracle(cfg):
    from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d
    up, down, pad, kshape = cfg
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 26, 28, generator=g)
    k = torch.randn(*kshape, generator=g)
    xr = x.clone().requires_grad_()
    yr = O.upfirdn2d(xr, k, up=up, down=down, pad=pad)
    xg = x.to(DEV).requires_grad_()
    yg = upfirdn2d(xg, k.to(DEV), up=up, down=down, pad=pad)
    close(yg, yr, rtol=1e-4, atol_rel=1e-6)
    gy = torch.randn(yr.shape, generator=g)
    (gxr,) = torch.autograd.grad(yr, xr, gy)
    gy_g = gy.to(DEV).requires_grad_()
    (gxg,) = torch.autograd.grad(yg, xg, gy_g, create_graph=True)
    close(gxg, gxr, rtol=1e-4, atol_rel=1e-6)
    # the op is linear: grad-grad w.r.t. gy of <gx, v> is the forward applied to v
    v = torch.randn(x.shape, generator=g)
    (gg,) = torch.autograd.grad((gxg * v.to(DEV)).sum(), gy_g)
    close(gg, O.upfirdn2d(v, k, up=up, down=down, pad=pad), rtol=1e-4, atol_rel=1e-6)


# ----------------------------------------------------------------------------- a4 Resample family
def test_resample_family_golden(ops, g_ops):
    x = dev(g_ops["rs_x"])
    close(ops.Resample(up=2).to(DEV)(x), g_ops["rs_up2"], 1e-5, 1e-6)
    close(ops.Resample(down=2).to(DEV)(x), g_ops["rs_down2"], 1e-5, 1e-6)
    close(ops.Resample().to(DEV)(x), g_ops["rs_blur4"], 1e-5, 1e-6)
    close(ops.Resample(window=[1, 2, 1], direction="h").to(DEV)(x), g_ops["rs_blur3_h"], 1e-5, 1e-6)
    close(ops.Resample(window=[1, 2, 1], direction="w").to(DEV)(x), g_ops["rs_blur3_w"], 1e-5, 1e-6)
    close(ops.Resample(up=2, ring=False).to(DEV)(x), g_ops["rs_up2_noring"], 1e-5, 1e-6)
    close(ops.BlurVH().to(DEV)(x), g_ops["rs_blurvh"], 1e-5, 1e-6)
    assert np.array_equal(ops.Pad(1, ring=True)(x).cpu().numpy(), g_ops["rs_pad1"])
    assert np.array_equal(ops.Pad(1, ring=True, mode="reflect")(x).cpu().numpy(), g_ops["rs_pad1_reflect"])
    close(ops.filter2d(x, dev(g_ops["rs_filter2d_k"])), g_ops["rs_filter2d"], 1e-5, 1e-6)
    for nm, mod in (("up2", ops.Resample(up=2)), ("down2", ops.Resample(down=2)), ("blur4", ops.Resample())):
        xg = x.clone().requires_grad_()
        (gx,) = torch.autograd.grad(mod.to(DEV)(xg), xg, dev(g_ops[f"rs_{nm}_gy"]))
        close(gx, g_ops[f"rs_{nm}_gx"], 1e-5, 1e-6)


@pytest.mark.parametrize("kw", [dict(up=2), dict(down=2), dict(), dict(window=[1, 2, 1], direction="h"),
                                dict(window=[1, 2, 1], direction="w"), dict(up=2, ring=False)])
@pytest.mark.parametrize("shape", [(2, 3, 4, 32), (1, 2, 64, 512), (2, 1, 6, 10)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_resample_vs_oracle(ops, kw, shape, dtype):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    xr = x.clone().requires_grad_()
    yr = O.resample(xr, **{k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items()})
    xg = x.to(DEV, dtype).requires_grad_()
    yg = ops.Resample(**kw).to(DEV)(xg)
    tol = dict(rtol=1e-3, atol_rel=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol_rel=1e-2)
    close(yg, yr, **tol)
    gy = torch.randn(yr.shape, generator=g)
    (gxr,) = torch.autograd.grad(yr, xr, gy)
    gy_g = gy.to(DEV, dtype).requires_grad_()
    (gxg,) = torch.autograd.grad(yg, xg, gy_g, create_graph=True)
    close(gxg, gxr, **tol)
    if dtype == torch.float32:
        v = torch.randn(shape, generator=g)
        (gg,) = torch.autograd.grad((gxg * v.to(DEV)).sum(), gy_g)
        close(gg, O.resample(v, **{k: (tuple(v2) if isinstance(v2, list) else v2) for k, v2 in kw.items()}),
              rtol=1e-4, atol_rel=1e-6)


@pytest.mark.parametrize("pad,ring,mode", [(1, True, "replicate"), (2, True, "reflect"), ((3, 0, 2, 1), False, "replicate"),
                                           ((5, 7, 0, 0), True, "reflect")])
def test_pad_fwd_bwd_exact(ops, pad, ring, mode):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 6, 9, generator=g)
    xr = x.clone().requires_grad_()
    yr = O.pad2d(xr, pad, ring=ring, mode=mode)
    xg = x.to(DEV).requires_grad_()
    yg = ops.Pad(pad, ring=ring, mode=mode)(xg)
    assert torch.equal(yg.cpu(), yr.detach())
    gy = torch.randn(yr.shape, generator=g)
    (gxr,) = torch.autograd.grad(yr, xr, gy)
    (gxg,) = torch.autograd.grad(yg, xg, gy.to(DEV))
    close(gxg, gxr, rtol=1e-5, atol_rel=1e-6)


# ----------------------------------------------------------------------------- a2 Fourier features
def test_fourier_golden_and_angle_pyramid(DF, g_ops):
    out = DF.fourier_features(dev(g_ops["ff_angle"]), dev(g_ops["ff_freqs"]), dev(g_ops["ff_phase"]))
    close(out, g_ops["ff_out"], rtol=0, atol_rel=5e-5)
    ob = DF.fourier_features(dev(g_ops["ff_angle"]), dev(g_ops["ff_freqs"]), dev(g_ops["ff_phase"]),
                             torch.bfloat16)
    close(ob, g_ops["ff_out"], rtol=0, atol_rel=8e-3)
    # angle pyramid step, incl. azimuth beyond +-pi (training-time shift)
    g = torch.Generator().manual_seed(7)
    el = torch.empty(3, 1, 16, 64).uniform_(-0.41, 0.05, generator=g)
    az = (torch.linspace(3.1, -3.1, 64)[None, None, None].expand(3, 1, 16, 64)
          + torch.tensor([0.0, 2.5, 6.0]).view(3, 1, 1, 1))
    ang = torch.cat([el, az], 1).contiguous()
    ref = O.downsample_angle(ang)
    got = DF.angle_down2(ang.to(DEV)).cpu()
    d = torch.remainder(got - ref + np.pi, 2 * np.pi) - np.pi        # compare modulo 2pi
    assert float(d.abs().max()) < 2e-6


@pytest.mark.parametrize("B,F,H,W", [(1, 256, 64, 512), (3, 256, 4, 32), (2, 5, 3, 7)])
def test_fourier_vs_oracle_full_size(DF, B, F, H, W):
    g = torch.Generator().manual_seed(8)
    ang = torch.stack([torch.empty(B, H, W).uniform_(-0.41, 0.05, generator=g),
                       torch.empty(B, H, W).uniform_(-3.14, 9.42, generator=g)], 1)
    freqs = torch.stack([torch.empty(F).uniform_(-256, 256, generator=g),
                         torch.randint(-7, 8, (F,), generator=g).float() * 16], 1)
    phase = torch.rand(F, generator=g) * 2 * np.pi
    ref = O.fourier_feature(ang, freqs, phase)
    got = DF.fourier_features(ang.to(DEV), freqs.to(DEV), phase.to(DEV))
    # identical fp32 argument arithmetic; device sincosf is within 2 ulp
    assert float((got.cpu() - ref).abs().max()) < 1e-6


# ----------------------------------------------------------------------------- a1 modulated conv
@pytest.mark.parametrize("tag,demod", [("dm", True), ("hd", False)])
def test_modconv_golden_module(ops, g_ops, tag, demod):
    sd = {k[len(f"mc_{tag}_sd_"):]: T(v) for k, v in g_ops.items() if k.startswith(f"mc_{tag}_sd_")}
    m = ops.ModConv2d(in_ch=12, out_ch=8 if demod else 1, mod_ch=16, ksize=1, stride=1, padding=0,
                      demod=demod, bias=not demod, ema=True)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    x = dev(g_ops[f"mc_{tag}_x"]).requires_grad_()
    st = dev(g_ops[f"mc_{tag}_style"]).requires_grad_()
    y = m(x, st)
    close(y, g_ops[f"mc_{tag}_y_eval"], rtol=1e-3, atol_rel=1e-5)
    params = [x, st, m.weight, m.mod.module.weight, m.mod.module.bias]
    grads = torch.autograd.grad(y, params, dev(g_ops[f"mc_{tag}_gy"]))
    for nm, gr in zip(("gx", "gstyle", "gw", "gmodw", "gmodb"), grads):
        close(gr, g_ops[f"mc_{tag}_{nm}"], rtol=1e-3, atol_rel=1e-4)
    m.train()
    y2 = m(x, st)
    close(y2, g_ops[f"mc_{tag}_y_train"], rtol=1e-3, atol_rel=1e-5)
    close(m.ema_var, g_ops[f"mc_{tag}_ema_after"], rtol=1e-5, atol_rel=0)


@pytest.mark.parametrize("B,Oc,C1,C2,B2,HW,act", [
    (2, 32, 64, 512, 2, (64, 512), 3),      # top-level conv1 shape (per-sample Fourier block)
    (3, 32, 64, 512, 1, (8, 64), 3),        # batch-shared Fourier block
    (2, 512, 0, 512, 1, (4, 32), 3),        # level 0: Fourier block only
    (2, 64, 64, 0, 1, (32, 256), 3),        # conv2: features only
    (2, 2, 32, 0, 1, (64, 512), 1),         # heads (O = 2)
    (1, 70, 33, 19, 1, (5, 12), 1),         # ragged sizes
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_modconv_bmm_vs_oracle(DF, B, Oc, C1, C2, B2, HW, act, dtype):
    g = torch.Generator().manual_seed(9)
    H, W = HW
    K = C1 + C2
    wb = torch.randn(B, Oc, K, generator=g) / np.sqrt(K)
    x1 = torch.randn(B, C1, H, W, generator=g) if C1 else None
    x2 = torch.randn(B2, C2, H, W, generator=g) if C2 else None
    bias = torch.randn(Oc, generator=g)
    if dtype == torch.bfloat16:
        wb = wb.bfloat16().float()
        x1 = None if x1 is None else x1.bfloat16().float()
        x2 = None if x2 is None else x2.bfloat16().float()
    wr = wb.clone().requires_grad_()
    x1r = None if x1 is None else x1.clone().requires_grad_()
    parts = ([x1r] if C1 else []) + ([x2.expand(B, -1, -1, -1)] if C2 else [])
    xin = torch.cat(parts, 1).reshape(B, K, H * W)
    yr = torch.bmm(wr, xin).reshape(B, Oc, H, W) + bias.view(1, -1, 1, 1)
    if act == 3:
        yr = O_lrelu(yr)
    wg = wb.to(DEV, dtype).requires_grad_()
    x1g = None if x1 is None else x1.to(DEV, dtype).requires_grad_()
    x2g = None if x2 is None else x2.to(DEV, dtype)
    bg = bias.to(DEV).requires_grad_()
    yg = DF.modconv_bmm(wg, x1g, x2g, bg, act, 0.2, O.SQRT2 if act == 3 else 1.0)
    tol = dict(rtol=1e-3, atol_rel=2e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol_rel=1e-2)
    close(yg, yr, **tol)
    gy = torch.randn(B, Oc, H, W, generator=g)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    # backward reference: the lrelu gate is discontinuous, so gate on the sign the device
    # actually produced (a |y| ~ 1e-7 element may round to either side) and check the
    # linear parts (dX = W^T g, dW = g X^T, db = sum g) against fp32 CPU matmuls
    if act == 3:
        gate = torch.where(yg.detach().float().cpu() > 0, 1.0, 0.2) * O.SQRT2
        gpre = gy * gate
    else:
        gpre = gy
    gp = gpre.reshape(B, Oc, H * W)
    dw_ref = torch.bmm(gp, xin.detach().transpose(1, 2))
    dx_ref = torch.bmm(wb.transpose(1, 2), gp)[:, :C1].reshape(B, C1, H, W) if C1 else None
    wanted_g = [wg, bg] + ([x1g] if C1 else [])
    gg = torch.autograd.grad(yg, wanted_g, gy.to(DEV, dtype))
    close(gg[0], dw_ref, **tol)
    if C1:
        close(gg[2], dx_ref, **tol)
    close(gg[1], gpre.sum((0, 2, 3)), rtol=tol["rtol"], atol_rel=max(tol["atol_rel"], 1e-4))


def O_lrelu(y):
    return torch.where(y > 0, y, y * 0.2) * O.SQRT2


# ----------------------------------------------------------------------------- a10 raydrop
def test_gumbel_raydrop_golden(DF, g_ops):
    logit = dev(g_ops["gs_logit"]).requires_grad_()
    img = dev(g_ops["gs_img"]).requires_grad_()
    u = dev(g_ops["gs_u"])
    mask, out = DF.gumbel_raydrop(logit, img, u, -1.0, 1.0)
    ref_mask = g_ops["gs_mask"]
    # device transcendentals differ in ulps: only near-ties (|logit + noise| ~ 0) may flip
    lo = np.log(g_ops["gs_u"]) - np.log1p(-g_ops["gs_u"]) + g_ops["gs_logit"]
    sure = np.abs(lo) > 1e-4
    assert np.array_equal(mask.detach().cpu().numpy()[sure], ref_mask[sure])
    assert np.array_equal(out.detach().cpu().numpy()[sure], g_ops["gs_image"][sure])
    _, _, count = DF.raydrop_count(logit.detach(), img.detach(), u, -1.0, 1.0)
    assert abs(int(count.item()) - int(g_ops["gs_count"])) <= int((~sure).sum())
    gl, gi = torch.autograd.grad(out, [logit, img], dev(g_ops["gs_gout"]))
    close(gl, g_ops["gs_glogit"], rtol=1e-3, atol_rel=1e-5)
    close(gi, g_ops["gs_gimg"], rtol=1e-6, atol_rel=0)


def test_gumbel_raydrop_full_size_count(DF):
    g = torch.Generator().manual_seed(11)
    logit = torch.randn(8, 1, 64, 512, generator=g) * 2
    img = torch.tanh(torch.randn(8, 1, 64, 512, generator=g))
    u = torch.rand(8, 1, 64, 512, generator=g)
    ref = O.raydrop(img, logit, u)
    mask, out, count = DF.raydrop_count(logit.to(DEV), img.to(DEV), u.to(DEV), -1.0, 1.0)
    lo = u.clamp(1e-7, 1 - 1e-7).logit() + logit
    sure = lo.abs() > 1e-4
    assert torch.equal(mask.cpu()[sure], ref["raydrop_mask"][sure])
    assert int(count.item()) == int(mask.sum().item())                 # kernel's own count
    assert abs(int(count.item()) - int(ref["raydrop_mask"].sum().item())) <= int((~sure).sum())
    assert torch.equal(out.cpu()[sure], ref["image"][sure])


# ----------------------------------------------------------------------------- a14 projection
def test_coord_bridge_bit_exact(g_coords):
    from dusty_gan_v2_b200.gans.coords import CoordBridge
    cb = CoordBridge(64, 512, 1.45, 80.0, "data/coords/kitti_raw.npy")
    assert np.array_equal(cb.angle.numpy(), g_coords["angle"])          # angle grid: bit exact
    cb = cb.to(DEV)
    xin = dev(g_coords["xin"])
    pm = cb.convert(xin, "inv_depth_norm", "point_map")
    assert int(cb.last_valid_count.item()) == int(g_coords["valid_count"])    # mask count
    ps = cb.convert(xin, "inv_depth_norm", "point_set")
    assert int(cb.last_valid_count.item()) == int(g_coords["valid_count"])
    # pixel indexing: point_set[b, h*W + w, :] == point_map[b, :, h, w]
    assert torch.equal(ps, pm.flatten(2).permute(0, 2, 1))
    # values: trig table comes from the host, products are plain fp32 -> bit exact
    assert np.array_equal(pm.cpu().numpy()[:, :, ::4, ::8], g_coords["point_map"])
    assert np.array_equal(ps.cpu().numpy()[:, :2048], g_coords["point_set_head"])
    dn = cb.convert(xin, "inv_depth_norm", "depth_norm")
    close(dn[:, :, ::4, ::8], g_coords["depth_norm"], rtol=1e-6, atol_rel=0)
    with pytest.raises(RuntimeError):
        cb.convert(xin.cpu(), "inv_depth_norm", "point_map")


# ----------------------------------------------------------------------------- a12 minibatch stddev
@pytest.mark.parametrize("shape", [(8, 6, 4, 4), (2, 6, 4, 4), (64, 512, 4, 32), (4, 3, 2, 2)])
def test_minibatch_stddev(ops, g_ops, shape):
    g = torch.Generator().manual_seed(12)
    x = torch.randn(shape, generator=g)
    xr = x.clone().requires_grad_()
    yr = O.minibatch_stddev(xr)
    xg = x.to(DEV).requires_grad_()
    yg = ops.MinibatchStdDev(4, 1)(xg)
    close(yg, yr, rtol=1e-4, atol_rel=1e-6)
    gy = torch.randn(yr.shape, generator=g)
    (gxr,) = torch.autograd.grad(yr, xr, gy, create_graph=True)
    gy_g = gy.to(DEV)
    (gx1,) = torch.autograd.grad(yg, xg, gy_g, retain_graph=True)           # fused kernel
    close(gx1, gxr, rtol=1e-3, atol_rel=1e-5)
    (gx2,) = torch.autograd.grad(yg, xg, gy_g, create_graph=True)           # differentiable path
    close(gx2, gxr, rtol=1e-3, atol_rel=1e-5)
    v = torch.randn(shape, generator=g)
    (hr,) = torch.autograd.grad((gxr * v).sum(), xr)
    (hg,) = torch.autograd.grad((gx2 * v.to(DEV)).sum(), xg)
    np.testing.assert_allclose(hg.cpu().numpy(), hr.numpy(), rtol=2e-3,
                               atol=max(1e-4 * float(hr.abs().max()), 1e-5))
    if shape == (8, 6, 4, 4):
        close(ops.MinibatchStdDev(4, 1)(dev(g_ops["mb_x"])), g_ops["mb_y"], 1e-4, 1e-6)
        close(ops.MinibatchStdDev(4, 1)(dev(g_ops["mb2_x"])), g_ops["mb2_y"], 1e-4, 1e-6)


@pytest.mark.parametrize("shape,cl", [((8, 6, 4, 4), False), ((8, 16, 4, 32), True), ((2, 8, 4, 4), True)])
def test_minibatch_std_stat_only(DF, shape, cl):
    """Statistic-only variant (D's NHWC epilogue): value of the appended channel and its
    gradient, first order (kernel) and differentiable (R1) paths, vs the oracle."""
    g = torch.Generator().manual_seed(13)
    x = torch.randn(shape, generator=g)
    xr = x.clone().requires_grad_()
    stat_r = O.minibatch_stddev(xr)[:, -1, 0, 0]                 # [B]
    xg = x.to(DEV)
    if cl:
        xg = xg.contiguous(memory_format=torch.channels_last)
    xg.requires_grad_()
    stat_g = DF.minibatch_std_stat(xg, 4)
    close(stat_g, stat_r, rtol=1e-4, atol_rel=1e-6)
    gs = torch.randn(shape[0], generator=g)
    (gxr,) = torch.autograd.grad(stat_r, xr, gs, create_graph=True)
    (gx1,) = torch.autograd.grad(stat_g, xg, gs.to(DEV), retain_graph=True)
    close(gx1, gxr, rtol=1e-3, atol_rel=1e-5)
    (gx2,) = torch.autograd.grad(stat_g, xg, gs.to(DEV), create_graph=True)
    close(gx2, gxr, rtol=1e-3, atol_rel=1e-5)
    v = torch.randn(shape, generator=g)
    (hr,) = torch.autograd.grad((gxr * v).sum(), xr)
    (hg,) = torch.autograd.grad((gx2 * v.to(DEV)).sum(), xg)
    np.testing.assert_allclose(hg.cpu().numpy(), hr.numpy(), rtol=2e-3,
                               atol=max(1e-4 * float(hr.abs().max()), 1e-5))


# ----------------------------------------------------------------------------- a13 / a7
def test_sumsq_rows_and_r1_grad(DF):
    g = torch.Generator().manual_seed(13)
    x = torch.randn(6, 1, 64, 512, generator=g)
    xg = x.to(DEV).requires_grad_()
    s = DF.sumsq_rows(xg)
    close(s, x.double().pow(2).sum((1, 2, 3)).float(), rtol=1e-5, atol_rel=0)
    (gx,) = torch.autograd.grad(s.mean(), xg)
    close(gx, 2 * x / 6, rtol=1e-6, atol_rel=0)
    close(DF.sumsq_total(x.to(DEV).bfloat16()), x.bfloat16().double().pow(2).sum().float(), 1e-4, 0)


def test_circular_unshift_vs_oracle(DF):
    g = torch.Generator().manual_seed(14)
    v = torch.randn(4, 2, 8, 64, generator=g)
    s01 = torch.tensor([0.0, 0.25, 0.7303, 0.999])
    vr = v.clone().requires_grad_()
    ref = O.circular_unshift(vr, s01 * 2 * np.pi) * 0.25
    vg = v.to(DEV).requires_grad_()
    got = DF.circular_unshift(vg, s01.to(DEV), 0.25)
    # the reference derives the sub-pixel offset through normalised fp32 grid coordinates
    # (error ~ W * 2^-23 per unit slope): compare at that resolution
    close(got, ref, rtol=1e-3, atol_rel=2e-4)
    gy = torch.randn(ref.shape, generator=g)
    (gr,) = torch.autograd.grad(ref, vr, gy)
    (gg,) = torch.autograd.grad(got, vg, gy.to(DEV))
    close(gg, gr, rtol=1e-3, atol_rel=2e-4)


# ----------------------------------------------------------------------------- a1 tcgen05 path
@pytest.mark.parametrize("B,Oc,C1,C2,B2,HW", [
    (2, 32, 64, 512, 2, (64, 512)),      # level 4 conv1, per-sample Fourier block
    (3, 32, 64, 512, 1, (8, 64)),        # batch-shared Fourier block
    (2, 512, 0, 512, 1, (4, 32)),        # level 0 (Fourier only), two N tiles of 256
    (2, 256, 512, 512, 2, (8, 64)),      # level 1 conv1, K = 1024
    (2, 128, 256, 512, 1, (16, 128)),    # level 2 conv1
    (2, 64, 64, 0, 1, (32, 256)),        # level 3 conv2
    (2, 32, 32, 0, 1, (64, 512)),        # level 4 conv2: K = 32 < one stage (TMA zero fill)
    (1, 48, 64, 0, 1, (2, 64)),          # O not a multiple of the N tile
    (8, 32, 64, 512, 1, (16, 128)),      # batch-fused tiles: 8 samples side by side (BN = 256)
    (16, 32, 64, 512, 1, (8, 64)),       # ... two column tiles
    (6, 32, 64, 512, 1, (8, 64)),        # ... group of 6 (BN = 192)
    (4, 64, 128, 512, 1, (16, 128)),     # level 3 conv1, 4 samples per tile
    (4, 128, 256, 512, 1, (8, 64)),      # level 2 conv1, 2 samples per tile
])
def test_modconv_tcgen05_vs_fp32_reference(DF, B, Oc, C1, C2, B2, HW):
    import dusty_gan_v2_b200 as pkg
    g = torch.Generator().manual_seed(21)
    H, W = HW
    K = C1 + C2
    bf = torch.bfloat16
    wb = (torch.randn(B, Oc, K, generator=g) / np.sqrt(K)).to(bf)
    x1 = torch.randn(B, C1, H, W, generator=g).to(bf) if C1 else None
    x2 = torch.randn(B2, C2, H, W, generator=g).to(bf) if C2 else None
    bias = torch.randn(Oc, generator=g)
    parts = ([x1.float()] if C1 else []) + ([x2.float().expand(B, -1, -1, -1)] if C2 else [])
    xin = torch.cat(parts, 1).reshape(B, K, H * W)
    ref = O_lrelu(torch.bmm(wb.float(), xin).reshape(B, Oc, H, W) + bias.view(1, -1, 1, 1))
    gy = torch.randn(B, Oc, H, W, generator=g).to(bf)
    res = {}
    try:
        for impl in (2, 1):
            pkg.set_modconv_impl(impl)
            wg = wb.to(DEV).requires_grad_()
            x1g = None if x1 is None else x1.to(DEV).requires_grad_()
            out = DF.modconv_bmm(wg, x1g, None if x2 is None else x2.to(DEV), bias.to(DEV), 3, 0.2,
                                 O.SQRT2)
            grads = torch.autograd.grad(out, [wg] + ([x1g] if C1 else []), gy.to(DEV))
            res[impl] = (out.detach(), grads)
    finally:
        pkg.set_modconv_impl(0)
    torch.cuda.synchronize()
    got, simt = res[2][0], res[1][0]
    # bf16 output rounding only: both kernels accumulate in fp32
    close(simt, ref, rtol=2e-2, atol_rel=4e-3)
    close(got, ref, rtol=2e-2, atol_rel=4e-3)
    assert float((got.float() - simt.float()).abs().max()) <= 2e-2 * float(ref.abs().max())
    # backward: fp32 reference gated on the tensor-core forward's own sign pattern
    gate = torch.where(got.float().cpu() > 0, 1.0, 0.2) * O.SQRT2
    gp = (gy.float() * gate).to(bf).float().reshape(B, Oc, H * W)        # g_pre is stored in bf16
    dw_ref = torch.bmm(gp, xin.transpose(1, 2))
    close(res[2][1][0], dw_ref, rtol=2e-2, atol_rel=1e-2)
    if C1:
        dx_ref = torch.bmm(wb.float().transpose(1, 2), gp)[:, :C1].reshape(B, C1, H, W)
        close(res[2][1][1], dx_ref, rtol=2e-2, atol_rel=1e-2)


@pytest.mark.parametrize("B,Oc,C1,C2,B2,HW", [
    (2, 32, 64, 512, 2, (16, 128)),      # level 4 conv1 family, per-sample Fourier block
    (3, 32, 64, 512, 1, (8, 64)),        # batch-shared Fourier block
    (2, 512, 0, 512, 1, (4, 32)),        # level 0 (Fourier only)
    (2, 128, 256, 512, 1, (16, 128)),    # level 2 conv1
    (2, 64, 64, 0, 1, (32, 256)),        # conv2 (no Fourier block)
    (2, 32, 32, 0, 1, (16, 128)),        # 96-channel split operand: zero-filled K tail
    (1, 48, 64, 0, 1, (2, 64)),          # O not a multiple of the N tile
])
def test_modconv_fp32_on_tensor_cores_split_bf16(DF, B, Oc, C1, C2, B2, HW):
    """g1: the fp32 mode's modulated contraction (forward, dX, dW) on tcgen05 through split-bf16
    operands and an fp32 epilogue, against an fp64 CPU bmm.  rtol 1e-4 (fp32 mode bound: 1e-3)."""
    g = torch.Generator().manual_seed(22)
    H, W = HW
    K = C1 + C2
    wb = torch.randn(B, Oc, K, generator=g) / np.sqrt(K)
    x1 = torch.randn(B, C1, H, W, generator=g) if C1 else None
    x2 = torch.randn(B2, C2, H, W, generator=g) if C2 else None
    bias = torch.randn(Oc, generator=g)
    parts = ([x1.double()] if C1 else []) + ([x2.double().expand(B, -1, -1, -1)] if C2 else [])
    xin = torch.cat(parts, 1).reshape(B, K, H * W)
    pre = torch.bmm(wb.double(), xin).reshape(B, Oc, H, W) + bias.double().view(1, -1, 1, 1)
    ref = torch.where(pre > 0, pre, 0.2 * pre) * O.SQRT2
    gy = torch.randn(B, Oc, H, W, generator=g)
    names = []
    orig = DF.K.call
    # impl argument: (..., impl, ema_var, ema_rows, sumsq, stream) for fwd, (..., impl, ema_var, ema_rows, stream)
    # for dx, (..., impl, dw_ld, stream) for dw
    impl_of = lambda name, a: {"dusty_modconv_bwd_dw": -3, "dusty_modconv_bwd_dx": -4, "dusty_modconv_fwd": -5}.get(name)  # noqa: E731
    impl_of = (lambda f: (lambda name, a: a[f(name, a)] if f(name, a) is not None else None))(impl_of)
    DF.K.call = lambda name, *a: (names.append((name, impl_of(name, a))), orig(name, *a))[1]
    try:
        wg = wb.to(DEV).requires_grad_()
        x1g = None if x1 is None else x1.to(DEV).requires_grad_()
        out = DF.modconv_bmm(wg, x1g, None if x2 is None else x2.to(DEV), bias.to(DEV), 3, 0.2, O.SQRT2)
        grads = torch.autograd.grad(out, [wg] + ([x1g] if C1 else []), gy.to(DEV))
    finally:
        DF.K.call = orig
    assert out.dtype == torch.float32
    assert ("dusty_modconv_fwd", 4) in names and ("dusty_modconv_bwd_dw", 2) in names, names
    assert not C1 or ("dusty_modconv_bwd_dx", 4) in names, names
    close(out, ref.float(), rtol=1e-4, atol_rel=2e-5)
    # gate from the device forward's own sign pattern (a pre-activation within 1e-5 of zero may
    # land on either side of it)
    gate = torch.where(out.detach().double().cpu() > 0, 1.0, 0.2) * O.SQRT2
    gp = (gy.double() * gate).reshape(B, Oc, H * W)
    # dW contracts over 3 * H * W pixels in ONE TMEM accumulator: the tensor core's fp32
    # accumulation truncates, a bias of ~2^-24 per UMMA step that grows with the chain length
    # (measured 1.7e-4 of the rms at 24576 terms, 6e-4 at 98304) -- inside the fp32 bound
    close(grads[0], torch.bmm(gp, xin.transpose(1, 2)).float(), rtol=1e-3, atol_rel=2e-4)
    if C1:
        dx_ref = torch.bmm(wb.double().transpose(1, 2), gp)[:, :C1].reshape(B, C1, H, W)
        close(grads[1], dx_ref.float(), rtol=1e-4, atol_rel=2e-5)


@pytest.mark.parametrize("demod", [True, False])
@pytest.mark.parametrize("B,Oc,I", [(3, 8, 12), (64, 32, 576), (5, 256, 1024), (4, 1, 32)])
def test_modprep_fused_vs_composite(ops, demod, B, Oc, I):
    """Fused weight-prep kernels (value + analytic gradient) against the same algebra in
    autograd-traced tensor ops."""
    torch.manual_seed(31)
    m = ops.ModConv2d(in_ch=I, out_ch=Oc, mod_ch=16, ksize=1, stride=1, padding=0, demod=demod,
                      bias=False, ema=True).to(DEV)
    m.mod.module.bias.data.normal_(0, 0.3)
    m.ema_var.fill_(0.8)
    style = torch.randn(B, 16, device=DEV, requires_grad=True)
    params = [style, m.weight, m.mod.module.weight, m.mod.module.bias]
    wb = m.effective_weights(style)
    ref = m._effective_weights_composite(m.mod(style.float()))
    close(wb, ref, rtol=1e-4, atol_rel=1e-6)
    gw = torch.randn_like(ref)
    g_fused = torch.autograd.grad(wb, params, gw)
    g_ref = torch.autograd.grad(ref, params, gw)
    for a, b_ in zip(g_fused, g_ref):
        close(a, b_, rtol=2e-3, atol_rel=2e-4)
    wb16 = m.effective_weights(style, torch.bfloat16)
    close(wb16, ref, rtol=1e-2, atol_rel=4e-3)


# ----------------------------------------------------------------------------- NHWC variants
@pytest.mark.parametrize("dtype,C", [(torch.float32, 4), (torch.float32, 32), (torch.bfloat16, 32),
                                     (torch.bfloat16, 256)])
def test_channels_last_ops_match_nchw(DF, ops, dtype, C):
    """bias_act / pad / blur on NHWC-stored tensors give the NCHW results (values, first and
    second order), and keep the NHWC layout."""
    g = torch.Generator().manual_seed(41)
    x = torch.randn(3, C, 8, 16, generator=g).to(DEV, dtype)
    b = torch.randn(C, generator=g).to(DEV)
    CL = torch.channels_last
    tol = dict(rtol=1e-5, atol_rel=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol_rel=1e-2)

    def run(fn, xin):
        xin = xin.clone().requires_grad_()
        y = fn(xin)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(DEV, dtype)
        gy = gy.contiguous(memory_format=CL) if DF._is_cl(xin) else gy
        gy = gy.requires_grad_()
        (gx,) = torch.autograd.grad(y, xin, gy, create_graph=True)
        v = torch.randn(xin.shape, generator=torch.Generator().manual_seed(6)).to(DEV, dtype)
        (gg,) = torch.autograd.grad((gx.float() * v.float()).sum(), gy)
        return y, gx, gg

    blur, pad = ops.Resample().to(DEV), ops.Pad(1, ring=True)
    for name, fn in (("bias_act", lambda t: DF.bias_act(t, b)), ("pad", pad), ("blur", blur)):
        ref = run(fn, x)
        got = run(fn, x.contiguous(memory_format=CL))
        assert DF._is_cl(got[0]), name
        for a, r in zip(got, ref):
            if name == "pad":
                assert torch.equal(a.contiguous(), r.contiguous()), name
            else:
                close(a, r, **tol)
    # bias gradient through the NHWC reduction
    bb = b.clone().requires_grad_()
    xc = x.contiguous(memory_format=CL)
    y = DF.bias_act(xc, bb)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(7)).to(DEV, dtype)
    (db,) = torch.autograd.grad(y, bb, gy.contiguous(memory_format=CL))
    bb2 = b.clone().requires_grad_()
    (db2,) = torch.autograd.grad(DF.bias_act(x, bb2), bb2, gy)
    close(db, db2, rtol=1e-3, atol_rel=1e-3)


@pytest.mark.parametrize("dtype,C", [(torch.float32, 8), (torch.bfloat16, 32)])
def test_fused_blur_pad_nhwc(DF, ops, dtype, C):
    """blur + ring pad as one NHWC kernel == Pad(1, ring)(Resample()(x)); value, adjoint and
    second order (the op is linear)."""
    g = torch.Generator().manual_seed(43)
    CL = torch.channels_last
    x = torch.randn(2, C, 8, 16, generator=g).to(DEV, dtype)
    blur, pad = ops.Resample().to(DEV), ops.Pad(1, ring=True)
    taps = tuple(blur.kernel.tolist())
    tol = dict(rtol=1e-5, atol_rel=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol_rel=1e-2)
    xr = x.clone().requires_grad_()
    ref = pad(blur(xr))
    xf = x.contiguous(memory_format=CL).requires_grad_()
    got = DF.blur_pad_cl(xf, taps)
    assert got.shape == ref.shape and DF._is_cl(got)
    close(got, ref, **tol)
    gy = torch.randn(ref.shape, generator=g).to(DEV, dtype)
    (gr,) = torch.autograd.grad(ref, xr, gy)
    gyf = gy.contiguous(memory_format=CL).requires_grad_()
    (gf,) = torch.autograd.grad(got, xf, gyf, create_graph=True)
    close(gf, gr, **tol)
    v = torch.randn(x.shape, generator=g).to(DEV, dtype)
    (gg,) = torch.autograd.grad((gf.float() * v.float()).sum(), gyf)
    close(gg, pad(blur(v)), **tol)


@pytest.mark.parametrize("dtype,C,H,W", [(torch.float32, 8, 8, 16), (torch.bfloat16, 32, 6, 20),
                                         (torch.bfloat16, 64, 2, 4), (torch.float32, 4, 64, 512)])
def test_blur_down2_nhwc(DF, ops, dtype, C, H, W):
    """blur evaluated at even positions only == Resample()(x)[:, :, ::2, ::2] (what the 1x1
    stride-2 skip convolution reads); value, adjoint and second order."""
    g = torch.Generator().manual_seed(44)
    CL = torch.channels_last
    x = torch.randn(2, C, H, W, generator=g).to(DEV, dtype)
    blur = ops.Resample().to(DEV)
    taps = tuple(blur.kernel.tolist())
    tol = dict(rtol=1e-5, atol_rel=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol_rel=1e-2)
    xr = x.clone().requires_grad_()
    ref = blur(xr)[:, :, ::2, ::2]
    xf = x.contiguous(memory_format=CL).requires_grad_()
    assert DF.blur_down2_cl_supported(xf)
    got = DF.blur_down2_cl(xf, taps)
    assert got.shape == ref.shape and DF._is_cl(got)
    close(got, ref, **tol)
    gy = torch.randn(ref.shape, generator=g).to(DEV, dtype)
    (gr,) = torch.autograd.grad(ref, xr, gy)
    gyf = gy.contiguous(memory_format=CL).requires_grad_()
    (gf,) = torch.autograd.grad(got, xf, gyf, create_graph=True)
    close(gf, gr, **tol)
    v = torch.randn(x.shape, generator=g).to(DEV, dtype)
    (gg,) = torch.autograd.grad((gf.float() * v.float()).sum(), gyf)
    close(gg, blur(v)[:, :, ::2, ::2], **tol)


@pytest.mark.parametrize("dtype,shape", [(torch.float32, (2, 3, 4, 8)), (torch.bfloat16, (2, 5, 6, 16)),
                                         (torch.float32, (3, 16, 32, 256)),
                                         (torch.bfloat16, (4, 64, 32, 256))])
def test_up2_with_sumsq(DF, ops, dtype, shape):
    """2x upsampling that also emits sum(y^2) (ModConv2d's EMA statistic, style.py:99-102): same y
    as the plain kernel bit for bit, statistic against the oracle, same adjoint."""
    g = torch.Generator().manual_seed(46)
    x = torch.randn(shape, generator=g)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    up = ops.Resample(up=2).to(DEV)
    taps = tuple(up.kernel.tolist())
    ref = O.resample(x, up=2)
    xd = x.to(DEV, dtype).requires_grad_()
    y, ss = DF.up2_with_sumsq(xd, taps)
    assert torch.equal(y, DF.resample4(xd.detach(), taps, 2))
    assert not ss.requires_grad and tuple(ss.shape) == (1,) and ss.dtype == torch.float32
    close(ss, ref.double().square().sum().float().reshape(1), rtol=1e-4 if dtype == torch.float32 else 2e-3,
          atol_rel=0)
    y2, ss2 = up.forward_with_sumsq(xd.detach())
    assert torch.equal(y2, y.detach())
    close(ss2, ss, rtol=1e-5, atol_rel=0)          # float atomics: summation order varies
    gy = torch.randn(ref.shape, generator=g).to(DEV, dtype)
    (gx,) = torch.autograd.grad(y, xd, gy)
    xr = xd.detach().clone().requires_grad_()
    (gr,) = torch.autograd.grad(DF.resample4(xr, taps, 2), xr, gy)
    assert torch.equal(gx, gr)


@pytest.mark.parametrize("dtype,C,H,W", [(torch.float32, 8, 8, 16), (torch.bfloat16, 32, 6, 20),
                                         (torch.bfloat16, 64, 2, 4), (torch.float32, 4, 64, 512),
                                         (torch.bfloat16, 128, 16, 128)])
def test_residual_fork_nhwc(DF, ops, dtype, C, H, W):
    """ResidualBlock input fork (dusty_v2.py:387-396): (Pad(1, ring)(x), blur(x)[::2, ::2]) from one
    autograd node whose backward folds both gradients in one kernel; against the oracle's
    pad2d / resample, first order (one and both consumers) and the R1-style second order."""
    g = torch.Generator().manual_seed(45)
    CL = torch.channels_last
    x = torch.randn(2, C, H, W, generator=g)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    tol = dict(rtol=1e-5, atol_rel=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol_rel=1e-2)
    xr = x.clone().requires_grad_()
    ref_p = O.pad2d(xr, 1, ring=True, mode="replicate")
    ref_d = O.resample(xr)[:, :, ::2, ::2]
    taps = tuple(ops.Resample().kernel.tolist())
    xf = x.to(DEV, dtype).contiguous(memory_format=CL).requires_grad_()
    assert DF.residual_fork_supported(xf)
    xp, xd = DF.residual_fork(xf, taps)
    assert DF._is_cl(xp) and DF._is_cl(xd)
    close(xp, ref_p, rtol=0, atol_rel=0)            # padding copies values: bit-exact
    close(xd, ref_d, **tol)
    gp = torch.randn(ref_p.shape, generator=g)
    gd = torch.randn(ref_d.shape, generator=g)
    if dtype == torch.bfloat16:
        gp, gd = gp.bfloat16().float(), gd.bfloat16().float()
    (gr,) = torch.autograd.grad([ref_p, ref_d], xr, [gp, gd], retain_graph=True)
    gpf = gp.to(DEV, dtype).contiguous(memory_format=CL).requires_grad_()
    gdf = gd.to(DEV, dtype).contiguous(memory_format=CL).requires_grad_()
    (gf,) = torch.autograd.grad([xp, xd], xf, [gpf, gdf], retain_graph=True)     # fused kernel
    assert DF._is_cl(gf)
    close(gf, gr, **tol)
    # one consumer only
    (g1,) = torch.autograd.grad(xp, xf, gpf, retain_graph=True)
    close(g1, torch.autograd.grad(ref_p, xr, gp, retain_graph=True)[0], **tol)
    (g2,) = torch.autograd.grad(xd, xf, gdf, retain_graph=True)
    close(g2, torch.autograd.grad(ref_d, xr, gd, retain_graph=True)[0], **tol)
    # second order: d/d(gp, gd) of <backward(gp, gd), v> = (pad(v), blur_down2(v))
    (gc,) = torch.autograd.grad([xp, xd], xf, [gpf, gdf], create_graph=True)
    close(gc, gr, **tol)
    v = torch.randn(x.shape, generator=g)
    if dtype == torch.bfloat16:
        v = v.bfloat16().float()
    ggp, ggd = torch.autograd.grad((gc.float() * v.to(DEV)).sum(), [gpf, gdf])
    close(ggp, O.pad2d(v, 1, ring=True, mode="replicate"), **tol)
    close(ggd, O.resample(v)[:, :, ::2, ::2], **tol)


# ----------------------------------------------------------------------------- a11 dense convs
@pytest.mark.parametrize("B,C,Oc,H,W,k,stride", [
    (2, 32, 32, 18, 66, 3, 1),       # RB0 conv1 shape family (C = O = 32, K_g = 96 -> zero-filled chunk)
    (2, 32, 64, 19, 67, 3, 2),       # strided 3x3, odd padded size
    (1, 64, 64, 34, 258, 3, 1),      # several 128-pixel patches per row
    (2, 64, 128, 18, 130, 3, 2),
    (2, 128, 128, 18, 34, 3, 1),     # patch = 4 rows x 32
    (1, 256, 512, 10, 66, 3, 2),     # two N tiles of 256
    (2, 256, 256, 6, 34, 3, 1),
    (2, 32, 64, 16, 64, 1, 2),       # skip branch: 1x1 stride 2
    (2, 48, 40, 9, 21, 3, 1),        # channel counts that are not powers of two
    (1, 16, 24, 7, 9, 2, 1),
    (2, 40, 64, 12, 70, 3, 1),       # halo kernel on the dgrad side only (contracts 64 channels)
    (2, 32, 32, 10, 20, 3, 1),       # narrow image: 32-pixel patch pitch
    (2, 32, 64, 8, 32, 1, 1),        # skip branch after the decimating blur: 1x1, unit stride
    (3, 64, 128, 11, 131, 3, 1),     # odd sizes, several tiles with ragged edges
])
@pytest.mark.parametrize("halo", [True, False])
def test_conv2d_tcgen05_fprop_dgrad_wgrad(DF, B, C, Oc, H, W, k, stride, halo):
    """Own implicit-GEMM convolution vs the oracle's F.conv2d (fp32, CPU) on bf16-rounded
    operands; tolerance = bf16 output rounding (rtol 2e-2)."""
    g = torch.Generator().manual_seed(33)
    bf = torch.bfloat16
    x = torch.randn(B, C, H, W, generator=g).to(bf)
    w = (torch.randn(Oc, C, k, k, generator=g) / np.sqrt(C * k * k)).to(bf)
    xr = x.float().requires_grad_()
    wr = w.float().requires_grad_()
    ref = torch.nn.functional.conv2d(xr, wr, None, stride)
    gy = torch.randn(ref.shape, generator=g).to(bf)
    gx_ref, gw_ref = torch.autograd.grad(ref, [xr, wr], gy.float())

    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    wd = w.to(DEV)
    s2 = (stride, stride)
    if halo and not (stride == 1 and (DF.conv_halo_ok(wd, "fprop") or DF.conv_halo_ok(wd, "dgrad"))):
        pytest.skip("shape does not take the halo-resident kernel")
    DF.set_conv_halo(halo)
    try:
        assert DF.conv_tc_supported(xd, wd, s2)
        _conv_checks(DF, xd, wd, w, s2, ref, gy, gx_ref, gw_ref, g, Oc, H, W)
    finally:
        DF.set_conv_halo(True)


def _conv_checks(DF, xd, wd, w, s2, ref, gy, gx_ref, gw_ref, g, Oc, H, W):
    y = DF.conv2d_fprop_tc(xd, wd, s2)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    close(y, ref, rtol=2e-2, atol_rel=4e-3)
    gx = DF.conv2d_dgrad_tc(gy.to(DEV), wd, s2, (H, W))
    close(gx, gx_ref, rtol=2e-2, atol_rel=4e-3)
    gw = DF.conv2d_wgrad_tc(gy.to(DEV), xd, s2, w.shape, torch.float32)
    assert gw.shape == w.shape
    close(gw, gw_ref, rtol=2e-2, atol_rel=4e-3)
    # fused bias + leaky-ReLU epilogue
    bias = torch.randn(Oc, generator=g)
    yb = DF.conv2d_fprop_tc(xd, wd, s2, bias.to(DEV), 3, 0.2, O.SQRT2)
    close(yb, O_lrelu(ref.detach() + bias.view(1, -1, 1, 1)), rtol=2e-2, atol_rel=4e-3)


@pytest.mark.parametrize("B,C,Oc,H,W,k,stride,layout", [
    (2, 32, 32, 18, 66, 3, 1, "nhwc"),      # RB0 conv1 family
    (2, 32, 64, 19, 67, 3, 2, "nhwc"),      # strided, odd padded size, four parity classes
    (2, 64, 128, 18, 130, 3, 2, "nchw"),    # NCHW source through the strided split
    (1, 256, 512, 10, 66, 3, 2, "nhwc"),    # 768-channel split operand, two N tiles
    (2, 128, 128, 6, 34, 3, 1, "nhwc"),
    (2, 32, 64, 16, 64, 1, 2, "nhwc"),      # 1x1 stride 2 (positions no tap reaches)
    (2, 48, 40, 9, 21, 3, 1, "nhwc"),       # channel counts that are not powers of two
])
def test_conv2d_fp32_on_tensor_cores_split_bf16(DF, B, C, Oc, H, W, k, stride, layout):
    """g1: fp32 convolutions as [hi|hi|lo] x [hi|lo|hi] bf16 contractions on the tcgen05 kernels
    with an fp32 epilogue, against fp64 CPU autograd of the same map.  Held to rtol 1e-4 (the
    fp32 mode's bound is 1e-3): the dropped terms are ~2^-16 per product."""
    from dusty_gan_v2_b200.gans.models.ops.common import conv2d_valid
    g = torch.Generator().manual_seed(36)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Oc, C, k, k, generator=g) / np.sqrt(C * k * k)
    xr, wr = x.double().requires_grad_(), w.double().requires_grad_()
    yr = torch.nn.functional.conv2d(xr, wr, None, stride)
    gy = torch.randn(yr.shape, generator=g)
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy.double(), create_graph=True)
    ggwr, = torch.autograd.grad(gxr.pow(2).sum(), [wr])
    fmt = torch.channels_last if layout == "nhwc" else torch.contiguous_format
    xg = x.to(DEV).contiguous(memory_format=fmt).requires_grad_()
    wg = w.to(DEV).requires_grad_()
    s2 = (stride, stride)
    assert DF.fp32_on_tensor_cores() and DF.conv_x3_supported(xg, wg, s2)
    # the three primitives
    y0 = DF.conv2d_fprop_x3(xg.detach(), wg.detach(), s2)
    assert y0.dtype == torch.float32 and y0.shape == yr.shape
    close(y0, yr.float(), rtol=1e-4, atol_rel=2e-5)
    close(DF.conv2d_dgrad_x3(gy.to(DEV), wg.detach(), s2, (H, W)), gxr.float(), rtol=1e-4, atol_rel=2e-5)
    close(DF.conv2d_wgrad_x3(gy.to(DEV), xg.detach(), s2, w.shape), gwr.float(), rtol=1e-4, atol_rel=2e-5)
    # first and second order through the module-level op (what Conv2d / R1 run in fp32 mode)
    names = []
    orig = DF.K.call
    DF.K.call = lambda name, *a: (names.append(name), orig(name, *a))[1]
    try:
        y = conv2d_valid(xg, wg, s2)
        gx, gw = torch.autograd.grad(y, [xg, wg], gy.to(DEV), create_graph=True)
        ggw, = torch.autograd.grad(gx.pow(2).sum(), [wg])
    finally:
        DF.K.call = orig
    assert "dusty_conv2d_simt" not in names and "dusty_split_bf16x3" in names, names
    for a, r in ((y, yr), (gx, gxr), (gw, gwr), (ggw, ggwr)):
        assert a.dtype == torch.float32
        close(a, r.float(), rtol=1e-4, atol_rel=2e-5)


def test_conv2d_valid_autograd_routes_through_tcgen05(DF, ops):
    """conv2d_valid: first and second order through the own tcgen05 kernels against fp32 CPU
    autograd of the same bilinear map (bf16-rounded operands)."""
    from dusty_gan_v2_b200.gans.models.ops.common import conv2d_valid
    g = torch.Generator().manual_seed(34)
    bf = torch.bfloat16
    x = torch.randn(2, 32, 12, 36, generator=g).to(bf)
    w = (torch.randn(64, 32, 3, 3, generator=g) / 17.0).to(bf)
    gy = torch.randn(2, 64, 5, 17, generator=g).to(bf)
    xr, wr = x.float().requires_grad_(), w.float().requires_grad_()
    yr = torch.nn.functional.conv2d(xr, wr, None, 2)
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy.float(), create_graph=True)
    ggwr, = torch.autograd.grad(gxr.pow(2).sum(), [wr])
    xg = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_()
    wg = w.to(DEV).requires_grad_()
    n0 = DF.K.launch_count()
    y = conv2d_valid(xg, wg, (2, 2))
    gx, gw = torch.autograd.grad(y, [xg, wg], gy.to(DEV), create_graph=True)
    ggw, = torch.autograd.grad(gx.float().pow(2).sum(), [wg])
    assert DF.K.launch_count() - n0 >= 6
    for a, b in ((y, yr), (gx, gxr), (gw, gwr), (ggw, ggwr)):
        close(a, b, rtol=3e-2, atol_rel=1e-2)


@pytest.mark.parametrize("B,C,Oc,H,W,k,stride,pad", [
    (2, 1, 8, 9, 20, 3, 1, 0),        # single-channel first layers
    (2, 2, 4, 16, 64, 1, 1, 0),       # the discriminator stem's 1x1 convolution
    (2, 9, 8, 6, 18, 3, 1, 1),        # odd channel count, zero padding (the 513-channel epilogue conv, small)
    (3, 4, 8, 18, 34, 3, 2, 0),       # stride 2
    (2, 6, 10, 12, 20, 4, 2, 1),      # 4x4 stride-2 padding-1 (vanilla discriminator)
])
@pytest.mark.parametrize("dtype,layout", [(torch.float32, "nchw"), (torch.float32, "nhwc"), (torch.bfloat16, "nhwc")])
def test_conv2d_simt_family_vs_cpu_autograd(DF, B, C, Oc, H, W, k, stride, pad, dtype, layout):
    """The CUDA-core convolution family (fp32 parity mode, shapes outside the tcgen05 domain):
    forward, data gradient, filter gradient and the R1-style second order through conv2d()
    against CPU autograd; exact fp32 FMAs -> rtol 1e-4 in fp32."""
    from dusty_gan_v2_b200.gans.models.ops.common import conv2d
    g = torch.Generator().manual_seed(35)
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    w = (torch.randn(Oc, C, k, k, generator=g) / np.sqrt(C * k * k)).to(dtype)
    b = torch.randn(Oc, generator=g)
    xr, wr = x.float().requires_grad_(), w.float().requires_grad_()
    yr = torch.nn.functional.conv2d(xr, wr, b, stride, pad)
    gy = torch.randn(yr.shape, generator=g).to(dtype)
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy.float(), create_graph=True)
    ggwr, = torch.autograd.grad(gxr.pow(2).sum(), [wr])
    fmt = torch.channels_last if layout == "nhwc" else torch.contiguous_format
    xg = x.to(DEV).contiguous(memory_format=fmt).requires_grad_()
    wg = w.to(DEV).requires_grad_()
    assert not DF.conv_tc_supported(xg, wg, (stride, stride)) or pad
    n0 = DF.K.launch_count()
    y = conv2d(xg, wg, b.to(DEV), (stride, stride), (pad, pad))
    gx, gw = torch.autograd.grad(y, [xg, wg], gy.to(DEV), create_graph=True)
    ggw, = torch.autograd.grad(gx.float().pow(2).sum(), [wg])
    assert DF.K.launch_count() - n0 >= 5
    tol = (1e-4, 1e-5) if dtype == torch.float32 else (2e-2, 1e-2)
    for a, r in ((y, yr), (gx, gxr), (gw, gwr), (ggw, ggwr)):
        assert a.dtype == dtype
        close(a, r, rtol=tol[0], atol_rel=tol[1])


@pytest.mark.parametrize("C,Oc", [(2, 32), (1, 16), (4, 64)])
def test_pointwise_fan_out_conv_writes_nhwc_and_takes_nhwc_gradients(DF, C, Oc):
    """The stem's 1x1 convolution as a single op (functional._Stem's composite under
    create_graph, i.e. the R1 double backward): a few-channel NCHW bf16 image goes in, the
    feature tensor comes out NHWC, an NHWC gradient is consumed in place (no layout copy) and
    the second-order terms come back NHWC too; values against CPU autograd."""
    from dusty_gan_v2_b200.gans.models.ops.common import conv2d_valid
    prev = "bf16" if DF.act_dtype() == torch.bfloat16 else "fp32"
    DF.set_precision("bf16")                     # the rule belongs to the bf16 NHWC feature stack
    g = torch.Generator().manual_seed(77)
    B, H, W = 3, 10, 24
    x = torch.randn(B, C, H, W, generator=g).bfloat16()
    w = (torch.randn(Oc, C, 1, 1, generator=g) / np.sqrt(C)).bfloat16()
    xr, wr = x.float().requires_grad_(), w.float().requires_grad_()
    yr = torch.nn.functional.conv2d(xr, wr)
    gy = torch.randn(yr.shape, generator=g).bfloat16()
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy.float(), create_graph=True)
    ggyr_src = gxr.pow(2).sum()
    ggwr, = torch.autograd.grad(ggyr_src, [wr], retain_graph=True)
    xg, wg = x.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    gyg = gy.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_()
    try:
        y = conv2d_valid(xg, wg, (1, 1))
        assert DF._is_cl(y), "fan-out convolution must produce NHWC"
        gx, gw = torch.autograd.grad(y, [xg, wg], gyg, create_graph=True)
        ggw, ggy = torch.autograd.grad(gx.float().pow(2).sum(), [wg, gyg])
        assert DF._is_cl(ggy), "the gradient w.r.t. the incoming gradient must stay NHWC"
    finally:
        DF.set_precision(prev)
    gyr = gy.float().requires_grad_()
    gxr2, = torch.autograd.grad(yr, [xr], gyr, create_graph=True)
    ggyr, = torch.autograd.grad(gxr2.pow(2).sum(), [gyr])
    for a, r in ((y, yr), (gx, gxr), (gw, gwr), (ggw, ggwr), (ggy, ggyr)):
        close(a, r, rtol=2e-2, atol_rel=1e-2)


@pytest.mark.parametrize("dtype,C,Oc", [(torch.float32, 6, 5), (torch.bfloat16, 32, 16), (torch.bfloat16, 12, 6)])
def test_conv_transpose2d_vs_cpu_autograd(DF, dtype, C, Oc):
    """ConvTranspose2d(4x4, stride 2, padding 1) of the vanilla / dusty_v1 generators
    (reference vanilla.py:18-27): the data-gradient kernels (tcgen05 for bf16 channel counts
    that qualify, CUDA cores otherwise) against CPU autograd, first order."""
    from dusty_gan_v2_b200.gans.models.ops.common import conv_transpose2d
    g = torch.Generator().manual_seed(36)
    x = torch.randn(2, C, 5, 9, generator=g).to(dtype)
    w = (torch.randn(C, Oc, 4, 4, generator=g) / np.sqrt(C * 4)).to(dtype)
    b = torch.randn(Oc, generator=g)
    xr, wr = x.float().requires_grad_(), w.float().requires_grad_()
    yr = torch.nn.functional.conv_transpose2d(xr, wr, b, 2, 1)
    gy = torch.randn(yr.shape, generator=g).to(dtype)
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy.float())
    xg, wg = x.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    y = conv_transpose2d(xg, wg, b.to(DEV), (2, 2), (1, 1))
    assert tuple(y.shape) == tuple(yr.shape)
    gx, gw = torch.autograd.grad(y, [xg, wg], gy.to(DEV))
    tol = (1e-4, 1e-5) if dtype == torch.float32 else (2e-2, 1e-2)
    for a, r in ((y, yr), (gx, gxr), (gw, gwr)):
        close(a, r, rtol=tol[0], atol_rel=tol[1])


@pytest.mark.parametrize("M,N,Kd", [(64, 512, 65536), (128, 512, 4096), (100, 72, 200), (64, 1, 512), (3, 512, 1)])
def test_linear_nt_forward_dgrad_wgrad_vs_cpu(DF, M, N, Kd):
    """The epilogue linears on own GEMMs (linear_tc.cu): y = x @ W.T, the data gradient
    dy @ W (W read as an MN-major operand) and the weight gradient dy.T @ x (both operands
    MN-major) against fp64 CPU products; TF32 operand rounding (rtol 2e-3) for the forward GEMM
    on the fp32 master weight, bf16 operands (2e-2) for the gradient GEMMs, exact fp32 on the
    CUDA-core path; then first and second order through autograd."""
    g = torch.Generator().manual_seed(50)
    x = torch.randn(M, Kd, generator=g)
    w = torch.randn(N, Kd, generator=g) / np.sqrt(Kd)
    dy = torch.randn(M, N, generator=g)
    xd, wd, dyd = x.to(DEV), w.to(DEV), dy.to(DEV)
    tol = dict(rtol=2e-3, atol_rel=2e-3)                   # TF32 operands (forward)
    tolg = dict(rtol=2e-2, atol_rel=1e-2)                 # bf16 operands (gradient GEMMs)
    if min(M, N) < 16:
        tolg = tol
    close(DF.matmul_nt(xd, wd, 0.5), 0.5 * (x.double() @ w.double().t()), **tol)
    close(DF.matmul_nt(dyd, wd.t()), dy.double() @ w.double(), **tolg)                 # dgrad: B MN-major
    close(DF.matmul_nt(dyd.t(), xd.t()), dy.double().t() @ x.double(), **tolg)         # wgrad: both MN-major
    xg, wg = xd.clone().requires_grad_(), wd.clone().requires_grad_()
    y = DF.linear_nt(xg, wg)
    gx, gw = torch.autograd.grad(y, [xg, wg], dyd, create_graph=True)
    close(gx, dy.double() @ w.double(), **tolg)
    close(gw, dy.double().t() @ x.double(), **tolg)
    (ggw,) = torch.autograd.grad(gx.square().sum(), [wg])           # R1-style second order
    xr, wr = x.double().requires_grad_(), w.double().requires_grad_()
    (gxr,) = torch.autograd.grad(xr @ wr.t(), [xr], dy.double(), create_graph=True)
    (ggwr,) = torch.autograd.grad(gxr.square().sum(), [wr])
    close(ggw, ggwr, rtol=3e-2, atol_rel=2e-2)


# ----------------------------------------------------------------------------- a11 fused tails
@pytest.mark.parametrize("C,HW", [(64, (8, 16)), (256, (4, 8)), (32, (16, 64))])
def test_residual_tail_vs_single_ops(DF, C, HW):
    """(lrelu(x + b) * sqrt2 + skip) / sqrt2 as one NHWC kernel against bias_act, add, mul:
    values, first-order gradients, and the create_graph re-expression."""
    g = torch.Generator().manual_seed(41)
    CL = torch.channels_last
    x = torch.randn(3, C, *HW, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=CL)
    sk = torch.randn(3, C, *HW, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=CL)
    b = (0.3 * torch.randn(C, generator=g)).to(DEV)
    gy = torch.randn(3, C, *HW, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=CL)
    c = 2 ** -0.5
    assert DF.residual_tail_supported(x, sk)
    res = []
    for fused in (True, False):
        xs, ss, bs = x.clone().requires_grad_(), sk.clone().requires_grad_(), b.clone().requires_grad_()
        y = DF.residual_tail(xs, bs, ss) if fused else (DF.bias_act(xs, bs) + ss) * c
        g1 = torch.autograd.grad(y, [xs, bs, ss], gy, create_graph=True)
        # second order: d/d(gy-path) through the gate is zero a.e.; check the linear map instead
        (g2,) = torch.autograd.grad(g1[0], [xs], torch.ones_like(g1[0]), allow_unused=True)
        res.append((y.detach().float(), [t.detach().float() for t in g1]))
        assert g2 is None or torch.isfinite(g2).all()
    # fp32 reference with bf16 operands
    pre = x.float() + b.to(torch.bfloat16).float().view(1, -1, 1, 1)
    ref = (torch.where(pre > 0, pre, 0.2 * pre) * 2 ** 0.5 + sk.float()) * c
    close(res[0][0], ref, rtol=1e-2, atol_rel=4e-3)
    close(res[0][0], res[1][0], rtol=2e-2, atol_rel=1e-2)
    gate = torch.where(pre > 0, 1.0, 0.2) * 2 ** 0.5
    gs = gy.float() * c
    close(res[0][1][2], gs, rtol=1e-2, atol_rel=4e-3)                        # d/dskip
    close(res[0][1][0], gs * gate, rtol=1e-2, atol_rel=4e-3)                 # d/dx
    close(res[0][1][1], (gs.bfloat16().float() * gate).sum((0, 2, 3)), rtol=2e-2, atol_rel=1e-2)   # d/dbias


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 16, 64), (3, 64, 10, 36), (2, 32, 64, 512)])
def test_conv_bias_act_blur_pad_fused_backward(ops, DF, B, C, H, W):
    """conv1 + bias / leaky ReLU epilogue + blur + ring pad as ONE autograd node whose backward runs
    the blur adjoint, the ring fold, the activation gate and the bias-gradient reduction in one
    kernel (dusty_blur4_cl_adj_act), against the same chain run stage by stage on the device (same
    bf16 roundings up to the intermediate the fused kernel never writes) and, at 2e-2, against fp32
    CPU autograd through the oracle's restatement; second order goes through the composite."""
    g = torch.Generator().manual_seed(41)
    bf = torch.bfloat16
    taps = (0.125, 0.375, 0.375, 0.125)
    x = torch.randn(B, C, H + 2, W + 2, generator=g).to(bf)
    w = (torch.randn(C, C, 3, 3, generator=g) / np.sqrt(9 * C)).to(bf)
    bias = torch.randn(C, generator=g) * 0.2
    gy = torch.randn(B, C, H + 2, W + 2, generator=g).to(bf)
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    res = {}
    for fused in (True, False):
        xg, wg, bg = xd.clone().requires_grad_(), w.to(DEV).requires_grad_(), bias.to(DEV).requires_grad_()
        names = []
        orig = DF.K.call
        DF.K.call = lambda name, *a: (names.append(name), orig(name, *a))[1]
        try:
            if fused:
                out = ops.conv_bias_act(xg, wg, bg, (1, 1), 0.2, O.SQRT2, None, blur_taps=taps)
            else:
                out = DF.blur_pad_cl(ops.conv_bias_act(xg, wg, bg, (1, 1), 0.2, O.SQRT2), taps)
            grads = torch.autograd.grad(out, [xg, wg, bg], gy.to(DEV))
        finally:
            DF.K.call = orig
        assert ("dusty_blur4_cl_adj_act" in names) == fused and (("dusty_bias_act_bwd_cl" in names) != fused), names
        res[fused] = (out.detach(), grads)
    assert torch.equal(res[True][0], res[False][0])
    for a, b_ in zip(res[True][1], res[False][1]):
        # the unfused chain rounds the blur adjoint's output to bf16 before the gate
        close(a, b_, rtol=2e-2, atol_rel=6e-3)
    # fp32 reference of the whole chain
    xr, wr, br = x.float().requires_grad_(), w.float().requires_grad_(), bias.clone().requires_grad_()
    pre = torch.nn.functional.conv2d(xr, wr) + br.view(1, -1, 1, 1)
    yact = torch.where(pre > 0, pre, 0.2 * pre) * O.SQRT2
    ref = O.pad2d(O.resample(yact), 1, ring=True, mode="replicate")
    close(res[True][0], ref, rtol=2e-2, atol_rel=6e-3)
    # gradients in the linear region the device took (gate of ITS activation output)
    y_dev = ops.conv_bias_act(xd, w.to(DEV), bias.to(DEV), (1, 1), 0.2, O.SQRT2).float().cpu()
    gate = torch.where(y_dev > 0, 1.0, 0.2) * O.SQRT2
    gblur, = torch.autograd.grad(O.pad2d(O.resample(yact), 1, ring=True, mode="replicate"), yact, gy.float())
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(pre, [xr, wr, br], gblur * gate)
    for a, r in zip(res[True][1], (gx_ref, gw_ref, gb_ref)):
        close(a, r, rtol=3e-2, atol_rel=1e-2)


def test_conv_bias_act_epilogue_vs_single_ops(ops, DF):
    """3x3 unit-stride convolution of a thin NHWC layer with bias + leaky ReLU in the tcgen05
    kernel's epilogue against conv2d_valid followed by bias_act."""
    g = torch.Generator().manual_seed(43)
    CL = torch.channels_last
    x = torch.randn(2, 32, 18, 66, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=CL)
    w = (torch.randn(32, 32, 3, 3, generator=g) / 17).to(DEV, torch.bfloat16).contiguous(memory_format=CL)
    b = (0.2 * torch.randn(32, generator=g)).to(DEV)
    gy = torch.randn(2, 32, 16, 64, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=CL)
    if not ops.conv_bias_act_supported(x, w, (1, 1)):
        pytest.skip("halo-resident convolution not selected for this shape")

    def rel_l2(a, b_):
        return float((a.float() - b_.float()).norm() / b_.float().norm())

    res = []
    for fused in (True, False):
        xs, ws, bs = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
        y = ops.conv_bias_act(xs, ws, bs, (1, 1)) if fused else DF.bias_act(ops.conv2d_valid(xs, ws, (1, 1)), bs)
        res.append((y.detach(), torch.autograd.grad(y, [xs, ws, bs], gy)))
    ref = torch.nn.functional.conv2d(x.float(), w.float()) + b.to(torch.bfloat16).float().view(1, -1, 1, 1)
    ref = torch.where(ref > 0, ref, 0.2 * ref) * 2 ** 0.5
    close(res[0][0], ref, rtol=2e-2, atol_rel=5e-3)
    for a, b_ in zip(res[0][1], res[1][1]):        # two gate patterns (bf16 rounding of ~0 values)
        assert rel_l2(a, b_) < 5e-2


@pytest.mark.parametrize("shape", [(64, 32, 3, 3), (128, 64, 1, 1), (24, 40, 3, 3)])
def test_equal_lr_weight_prep_kernel(DF, shape):
    """scale + cast + OIHW -> OHWI in one kernel against the three tensor ops, with its adjoint
    (first order) and the adjoint's adjoint (second order: the map is linear)."""
    g = torch.Generator().manual_seed(47)
    w = torch.randn(*shape, generator=g).to(DEV).requires_grad_()
    s = 0.37
    CL = torch.channels_last
    out = DF.prep_conv_weight(w, s, torch.bfloat16)
    ref = (w * s).to(torch.bfloat16).contiguous(memory_format=CL)
    assert out.is_contiguous(memory_format=CL) and out.dtype == torch.bfloat16
    assert torch.equal(out.detach().float().cpu(), ref.detach().float().cpu())
    out2, tco = DF.prep_conv_weight(w, s, torch.bfloat16, with_tco=True)
    assert torch.equal(out2.detach().float().cpu(), ref.detach().float().cpu()) and not tco.requires_grad
    O_, C_, R_, S_ = shape
    assert torch.equal(tco.float().cpu(), ref.detach().permute(2, 3, 1, 0).reshape(R_ * S_, C_, O_).float().cpu())
    for fmt in (CL, torch.contiguous_format):
        gy = torch.randn(*shape, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=fmt)
        (gw,) = torch.autograd.grad(out, w, gy, retain_graph=True)
        close(gw, gy.float() * s, rtol=1e-6, atol_rel=1e-7)
    # second order: d/dg of <prep_adj(g), v> = prep(v)
    gyr = torch.randn(*shape, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=CL).requires_grad_()
    (gw,) = torch.autograd.grad(out, w, gyr, create_graph=True)
    v = torch.randn(*shape, generator=g).to(DEV)
    (gg,) = torch.autograd.grad(gw, gyr, v)
    close(gg, (v * s).to(torch.bfloat16), rtol=1e-2, atol_rel=1e-3)


@pytest.mark.parametrize("tag,unfold", [("unfold", True), ("elev", False)])
def test_kitti_scan_to_image_exact(g_kitti, tag, unfold):
    """f4: scan -> [6, 64, 512] range image on the device (depth-ordered scatter as an atomic
    min, NEAREST column selection, mask product) against the reference's golden output and the
    oracle: integer-exact pixel indexing, bit-equal values."""
    from dusty_gan_v2_b200.gans.datasets.kitti import KITTIRaw, scan_to_image
    pts = g_kitti["points"]
    img = scan_to_image(pts, (64, 512), 1.45, 80.0, unfold)
    assert img.shape == (6, 64, 512) and img.is_cuda
    assert np.array_equal(img.cpu().numpy(), g_kitti[f"{tag}_image"])
    full = scan_to_image(pts, (64, 2048), 1.45, 80.0, unfold)
    assert np.array_equal(full.cpu().numpy(), O.scan_to_image(pts, 64, 2048, 2048, 1.45, 80.0, unfold))
    import os, tempfile
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "0000000000.bin")
        pts.tofile(f)
        ds = KITTIRaw(shape=(64, 512), min_depth=1.45, max_depth=80.0, scan_unfolding=unfold, files=[f])
        item = ds[0]
    assert len(ds) == 1 and set(item) == {"xyz", "reflectance", "depth", "mask"}
    assert np.array_equal(item["depth"].cpu().numpy(), g_kitti[f"{tag}_image"][4:5])
    assert int(item["mask"].sum()) == int(g_kitti[f"{tag}_image"][5].sum())          # valid-point count
    with pytest.raises(RuntimeError):
        scan_to_image(pts, (64, 512), device="cpu")



def test_modconv_epilogue_sumsq_matches_separate_pass(DF):
    """sum(y^2) accumulated by the tcgen05 contraction's epilogue (the next ModConv2d's EMA
    statistic) against the stand-alone reduction over the stored tensor, per-sample tiles and
    batch-fused tiles."""
    g = torch.Generator().manual_seed(23)
    bf = torch.bfloat16
    for (B, Oc, C1, C2, B2, HW) in [(3, 64, 64, 0, 1, (16, 128)), (8, 32, 64, 512, 1, (16, 128)),
                                     (2, 48, 64, 0, 1, (2, 64))]:
        K_ = C1 + C2
        wb = (torch.randn(B, Oc, K_, generator=g) / np.sqrt(K_)).to(bf).to(DEV)
        x1 = torch.randn(B, C1, *HW, generator=g).to(bf).to(DEV)
        x2 = torch.randn(B2, C2, *HW, generator=g).to(bf).to(DEV) if C2 else None
        bias = torch.randn(Oc, generator=g).to(DEV)
        y = DF.modconv_bmm(wb, x1, x2, bias, 3, 0.2, O.SQRT2, want_sumsq=True)
        assert hasattr(y, "_dusty_sumsq")
        ref = float(y.float().pow(2).sum())
        assert abs(float(y._dusty_sumsq) - ref) <= 2e-4 * ref, (float(y._dusty_sumsq), ref)
        y2 = DF.modconv_bmm(wb, x1, x2, bias, 3, 0.2, O.SQRT2)
        assert torch.equal(y, y2) and not hasattr(y2, "_dusty_sumsq")
