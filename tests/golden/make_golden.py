"""Generate golden vectors by running the UNMODIFIED reference (CPU path) in the
build container.  Usage (only where /root/reference exists):

    python tests/golden/make_golden.py

Writes small .npz fixtures next to this file.  The fixtures are committed; the
GPU box never needs /root/reference.  Seeds are fixed, so re-running reproduces
the same files (up to BLAS summation order).
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

ref_import.install()

from gans.augment import adaptive_augment as ref_ada  # noqa: E402
from gans.coords import CoordBridge  # noqa: E402
from gans.models import ops as rops  # noqa: E402
from gans.models.builder import build_discriminator, build_generator  # noqa: E402
from gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d_native  # noqa: E402

torch.set_num_threads(4)


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrs)} arrays")


def seed(s):
    torch.manual_seed(s)
    np.random.seed(s)


# ---------------------------------------------------------------------------- ops
def golden_ops():
    out = {}
    seed(10)
    # --- fused bias act (reference CPU branch + autograd) ---
    x = torch.randn(2, 3, 4, 5, requires_grad=True)
    b = torch.randn(3, requires_grad=True)
    y = rops.fused_leaky_relu(x, b)
    dy = torch.randn_like(y)
    dx, db = torch.autograd.grad(y, [x, b], dy)
    out.update(ba_x=npy(x), ba_b=npy(b), ba_y=npy(y), ba_dy=npy(dy), ba_dx=npy(dx), ba_db=npy(db))
    x2 = torch.randn(4, 6)
    b2 = torch.randn(6)
    out.update(ba2_x=npy(x2), ba2_b=npy(b2), ba2_y=npy(rops.fused_leaky_relu(x2, b2)))

    # --- upfirdn2d_native ---
    sym6 = torch.tensor(ref_ada.SYM6)
    cases = [
        ("ada_upx", sym6[None], (2, 1), (1, 1), (6, 5, 0, 0), (2, 1, 9, 14)),
        ("ada_upy", sym6[:, None], (1, 2), (1, 1), (0, 0, 6, 5), (2, 1, 9, 28)),
        ("ada_dnx", sym6.flip(0)[None], (1, 1), (2, 1), (-1, -1, 0, 0), (2, 1, 18, 36)),
        ("ada_dny", sym6.flip(0)[:, None], (1, 1), (1, 2), (0, 0, -1, -1), (2, 1, 18, 16)),
        ("gen2d", torch.randn(3, 4), (2, 3), (3, 2), (2, 1, -1, 3), (2, 2, 7, 9)),
        ("ident", torch.ones(1, 1), (1, 1), (1, 1), (0, 0, 0, 0), (1, 2, 3, 4)),
    ]
    for name, k, up, down, pad, shape in cases:
        xin = torch.randn(*shape)
        yo = upfirdn2d_native(xin, k, *up, *down, *pad)
        out.update({f"ufd_{name}_x": npy(xin), f"ufd_{name}_k": npy(k),
                    f"ufd_{name}_cfg": np.array([*up, *down, *pad]), f"ufd_{name}_y": npy(yo)})

    # --- Resample family ---
    xr = torch.randn(2, 3, 8, 16)
    out["rs_x"] = npy(xr)
    out["rs_up2"] = npy(rops.Resample(up=2)(xr))
    out["rs_down2"] = npy(rops.Resample(down=2)(xr))
    out["rs_blur4"] = npy(rops.Resample()(xr))
    out["rs_blur3_h"] = npy(rops.Resample(window=[1, 2, 1], direction="h")(xr))
    out["rs_blur3_w"] = npy(rops.Resample(window=[1, 2, 1], direction="w")(xr))
    out["rs_up2_noring"] = npy(rops.Resample(up=2, ring=False)(xr))
    out["rs_blurvh"] = npy(rops.BlurVH()(xr))
    out["rs_pad1"] = npy(rops.Pad(1, ring=True)(xr))
    out["rs_pad1_reflect"] = npy(rops.Pad(1, ring=True, mode="reflect")(xr))
    kk = torch.tensor([0.25, 0.5, 1.0, 0.5, 0.25])
    out["rs_filter2d_k"] = npy(kk)
    out["rs_filter2d"] = npy(rops.filter2d(xr, kk))
    # gradient of up2 / down2 / blur (adjoint with boundary folding)
    for nm, mod in (("up2", rops.Resample(up=2)), ("down2", rops.Resample(down=2)),
                    ("blur4", rops.Resample())):
        xg = xr.clone().requires_grad_()
        yg = mod(xg)
        g = torch.randn_like(yg)
        (gx,) = torch.autograd.grad(yg, xg, g)
        out[f"rs_{nm}_gy"] = npy(g)
        out[f"rs_{nm}_gx"] = npy(gx)

    # --- FourierFeature ---
    seed(11)
    ff = rops.FourierFeature(resolution=(8, 16), num_freqs=32)
    ang = torch.stack([torch.empty(2, 8, 16).uniform_(-0.4, 0.05),
                       torch.empty(2, 8, 16).uniform_(-3.1, 9.4)], dim=1)
    out.update(ff_freqs=npy(ff.freqs), ff_phase=npy(ff.phase), ff_angle=npy(ang),
               ff_out=npy(ff(ang)), ff_L=np.array([ff.L_h, ff.L_w]))

    # --- ModConv2d 1x1 ---
    seed(12)
    for tag, demod, bias in (("dm", True, False), ("hd", False, True)):
        m = rops.ModConv2d(in_ch=12, out_ch=8 if demod else 1, mod_ch=16, ksize=1, stride=1,
                           padding=0, demod=demod, bias=bias, ema=True)
        if bias:
            m.bias.data.normal_()
        m.mod.module.bias.data.normal_(0, 0.3)
        m.ema_var.fill_(0.7)
        xm = torch.randn(3, 12, 4, 6, requires_grad=True)
        st = torch.randn(3, 16, requires_grad=True)
        sd = {k: npy(v) for k, v in m.state_dict().items()}
        m.eval()
        y_eval = m(xm, st)
        gy = torch.randn_like(y_eval)
        params = [xm, st, m.weight, m.mod.module.weight, m.mod.module.bias]
        grads = torch.autograd.grad(y_eval, params, gy)
        m.train()
        y_train = m(xm, st)
        out.update({f"mc_{tag}_sd_{k}": v for k, v in sd.items()})
        out.update({f"mc_{tag}_x": npy(xm), f"mc_{tag}_style": npy(st), f"mc_{tag}_y_eval": npy(y_eval),
                    f"mc_{tag}_gy": npy(gy), f"mc_{tag}_y_train": npy(y_train),
                    f"mc_{tag}_ema_after": npy(m.ema_var)})
        for nm, g in zip(("gx", "gstyle", "gw", "gmodw", "gmodb"), grads):
            out[f"mc_{tag}_{nm}"] = npy(g)

    # --- GumbelSigmoid + RayDropModel ---
    from gans.models.dusty_v1 import RayDropModel
    rd = RayDropModel(raydrop_const=-1, gumbel_temperature=1)
    logit = (torch.randn(2, 1, 8, 16) * 3).requires_grad_()
    img = torch.tanh(torch.randn(2, 1, 8, 16)).requires_grad_()
    torch.manual_seed(77)
    o = rd({"image": img, "raydrop_logit": logit})
    torch.manual_seed(77)
    u = torch.rand(logit.shape)
    gi = torch.randn_like(img)
    g_logit, g_img = torch.autograd.grad(o["image"], [logit, img], gi)
    out.update(gs_logit=npy(logit), gs_img=npy(img), gs_u=npy(u), gs_mask=npy(o["raydrop_mask"]),
               gs_image=npy(o["image"]), gs_gout=npy(gi), gs_glogit=npy(g_logit), gs_gimg=npy(g_img),
               gs_count=np.array(int(o["raydrop_mask"].sum().item())))

    # --- MinibatchStdDev ---
    xs = torch.randn(8, 6, 4, 4)
    out.update(mb_x=npy(xs), mb_y=npy(rops.MinibatchStdDev(4, 1)(xs)))
    xs2 = torch.randn(2, 6, 4, 4)
    out.update(mb2_x=npy(xs2), mb2_y=npy(rops.MinibatchStdDev(4, 1)(xs2)))

    # --- PixelNorm / EqualLR linear ---
    zz = torch.randn(3, 16)
    out.update(pn_x=npy(zz), pn_y=npy(rops.PixelNorm()(zz)))
    save("ops.npz", **out)


# ------------------------------------------------------------------------- coords
def golden_coords():
    angle_file = os.path.join(ref_import.REFERENCE_ROOT, "data/coords/kitti_raw.npy")
    raw = np.load(angle_file)
    cb = CoordBridge(64, 512, 1.45, 80.0, angle_file)
    seed(20)
    depth = 1.45 + (80 - 1.45) * torch.rand(2, 1, 64, 512)
    depth[0, 0, :4] = 0.3          # below min depth -> invalid
    depth[1, 0, 5:7] = 120.0       # above max depth -> invalid
    mask = (torch.rand(2, 1, 64, 512) < 0.85).float()
    x = cb.convert(depth, "depth", "inv_depth_norm")
    reals = mask * (x * 2 - 1) + (1 - mask) * -1.0
    xin = (reals + 1) / 2
    pm = cb.convert(xin.clone(), "inv_depth_norm", "point_map")
    ps = cb.convert(xin.clone(), "inv_depth_norm", "point_set")
    inv = xin / 1.45
    valid = (xin > 1e-11).float() * cb.get_mask(inv, "inv_depth").float()
    dn = cb.convert(xin.clone(), "inv_depth_norm", "depth_norm")
    save("coords.npz",
         raw_sha256=np.array(hashlib.sha256(raw.tobytes()).hexdigest()),
         angle=npy(cb.angle), depth=npy(depth)[:, :, ::4, ::8], mask=npy(mask)[:, :, ::4, ::8],
         reals=npy(reals)[:, :, ::4, ::8], xin=npy(xin), point_map=npy(pm)[:, :, ::4, ::8],
         point_set_head=npy(ps)[:, :2048], depth_norm=npy(dn)[:, :, ::4, ::8],
         valid_count=np.array(int(valid.sum().item())))


# ------------------------------------------------------------------------- models
G_SMALL = dict(
    arch="dusty_v2",
    mapping_kwargs=dict(in_ch=16, out_ch=16, depth=2),
    synthesis_kwargs=dict(
        in_ch=16,
        out_ch=[dict(name="image", ch=1, act="nn.Tanh"), dict(name="raydrop_logit", ch=1, act=None)],
        ch_base=4, ch_max=16, resolution=[16, 64], layers=[2, 2, 2, 2], ring=True,
        num_fp16_layers=-1, use_noise=False, pe_type="random", pe_scale_offset=[3, -1],
        aug_coords=True, aug_coords_blitting=False),
    measurement_kwargs=dict(raydrop_const=-1, gumbel_temperature=1),
)
D_SMALL = dict(arch="dusty_v2", layer_kwargs=dict(in_ch=1, ring=True, ch_base=4, ch_max=8,
                                                  resolution=[16, 64], mbdis_group=4, mbdis_feat=1,
                                                  num_fp16_layers=-1, pre_blur=True))


def small_angle(b, h=16, w=64):
    el = torch.linspace(0.05, -0.41, h)[:, None].expand(h, w)
    az = (torch.arange(w) + 0.5) / w * 2 * np.pi - np.pi
    az = -az[None].expand(h, w)
    return torch.stack([el, az], 0)[None].repeat(b, 1, 1, 1).contiguous()


def golden_generator():
    seed(30)
    G = build_generator(ref_import.to_attr(G_SMALL))
    # de-trivialise zero-initialised params/buffers so parity is sensitive to them
    with torch.no_grad():
        for n, p in G.named_parameters():
            if "bias" in n:
                p.normal_(0, 0.2)
        for n, bf in G.named_buffers():
            if n.endswith("ema_var"):
                bf.fill_(float(torch.empty(1).uniform_(0.5, 1.5)))
        G.w_avg.normal_(0, 0.1)
    sd0 = {k: npy(v).copy() for k, v in G.state_dict().items()}
    B = 4
    z = torch.randn(B, 16)
    angle = small_angle(B)
    out = {f"sd_{k}": v for k, v in sd0.items()}
    out.update(z=npy(z), angle=npy(angle))

    # eval forward, psi = 1 and 0.7
    G.eval()
    for psi, tag in ((1.0, "eval"), (0.7, "psi")):
        torch.manual_seed(31)
        with torch.no_grad():
            o = G(z, angle=angle, truncation_psi=psi)
        torch.manual_seed(31)
        u = torch.rand(B, 1, 16, 64)
        out[f"{tag}_u"] = npy(u)
        for k in ("image", "image_orig", "raydrop_logit", "raydrop_mask"):
            out[f"{tag}_{k}"] = npy(o[k])
        out[f"{tag}_w0"] = npy(o["w"][:, 0])

    # training forward + backward (aug-coords shift, EMA side effects)
    G.train()
    for p in G.parameters():
        p.requires_grad_(True)
    torch.manual_seed(32)
    o = G(z, angle=angle)
    torch.manual_seed(32)
    shifts01 = torch.rand(B)            # what shifts[:,1].uniform_(0,1) consumed
    u = torch.rand(B, 1, 16, 64)
    gi = torch.randn(B, 1, 16, 64)
    gl = torch.randn(B, 1, 16, 64)
    loss = (o["image"] * gi).sum() + (o["raydrop_logit"] * gl).sum()
    names = [n for n, _ in G.named_parameters()]
    grads = torch.autograd.grad(loss, list(G.parameters()), allow_unused=True)
    out.update(train_shift01=npy(shifts01), train_u=npy(u), train_gi=npy(gi), train_gl=npy(gl))
    for k in ("image", "image_orig", "raydrop_logit", "raydrop_mask"):
        out[f"train_{k}"] = npy(o[k])
    for n, g in zip(names, grads):
        if g is not None:
            out[f"grad_{n}"] = npy(g)
    for n, bf in G.named_buffers():
        if n.endswith("ema_var") or n == "w_avg":
            out[f"after_{n}"] = npy(bf)
    save("g_small.npz", **out)


def golden_discriminator():
    seed(40)
    D = build_discriminator(ref_import.to_attr(D_SMALL))
    with torch.no_grad():
        for n, p in D.named_parameters():
            if "bias" in n:
                p.normal_(0, 0.2)
    out = {f"sd_{k}": npy(v).copy() for k, v in D.state_dict().items()}
    B = 8
    x = torch.tanh(torch.randn(B, 1, 16, 64)).requires_grad_()
    for p in D.parameters():
        p.requires_grad_(True)
    y = D(x)
    names = [n for n, _ in D.named_parameters()]
    # first order (non-saturating D loss on "real")
    loss = torch.nn.functional.softplus(-y).mean()
    g1 = torch.autograd.grad(loss, [x] + list(D.parameters()), retain_graph=True)
    # R1: double backward
    (gx,) = torch.autograd.grad(y.sum(), x, create_graph=True)
    r1 = gx.pow(2).sum(dim=[1, 2, 3]).mean()
    g2 = torch.autograd.grad(r1, list(D.parameters()), allow_unused=True)
    out.update(x=npy(x), y=npy(y), gx_loss=npy(g1[0]), r1=npy(r1), r1_gx=npy(gx))
    for n, g in zip(names, g1[1:]):
        out[f"g1_{n}"] = npy(g)
    for n, g in zip(names, g2):
        if g is not None:
            out[f"g2_{n}"] = npy(g)
    save("d_small.npz", **out)


def golden_ada():
    """Reference AdaptiveAugment.forward with its samplers pinned to recorded matrices."""
    seed(50)
    ada = ref_ada.AdaptiveAugment(p_init=0.9, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1,
                                  frac_trans=1, brightness=1, contrast=1, luma_flip=1, hue=1,
                                  saturation=1)
    B, H, W = 4, 16, 64
    G = ada.sample_affine(B, H, W)
    C = ada.sample_color(B)
    G[0] = torch.eye(3)                      # keep one identity sample
    ada.sample_affine = lambda *a, **k: G.clone()
    ada.sample_color = lambda *a, **k: C.clone()
    x = torch.tanh(torch.randn(B, 1, H, W)).requires_grad_()
    y = ada(x)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    save("ada.npz", G=npy(G), G_inv=npy(torch.inverse(G)), C=npy(C), x=npy(x), y=npy(y), gy=npy(gy),
         gx=npy(gx))


V_SYN = dict(in_ch=16, ch_base=4, ch_max=16, resolution=[32, 64], ring=True)
V1_SMALL = dict(arch="dusty_v1", synthesis_kwargs=dict(V_SYN, out_ch=[
    dict(name="image", ch=1, act=None), dict(name="raydrop_logit", ch=1, act=None)]),
    measurement_kwargs=dict(raydrop_const=-1, gumbel_temperature=1))
VD_SMALL = dict(arch="vanilla", layer_kwargs=dict(in_ch=1, ring=True, ch_base=4, ch_max=16,
                                                  resolution=[32, 64]))


def golden_vanilla():
    """dusty_v1 generator (vanilla synthesis + raydrop) and vanilla discriminator."""
    seed(60)
    G = build_generator(ref_import.to_attr(V1_SMALL)).eval()
    D = build_discriminator(ref_import.to_attr(VD_SMALL))
    with torch.no_grad():
        for n, p_ in list(G.named_parameters()) + list(D.named_parameters()):
            if "bias" in n:
                p_.normal_(0, 0.2)
    out = {f"sdG_{k}": npy(v).copy() for k, v in G.state_dict().items()}
    out.update({f"sdD_{k}": npy(v).copy() for k, v in D.state_dict().items()})
    B = 4
    z = torch.randn(B, 16, requires_grad=True)
    for p_ in list(G.parameters()) + list(D.parameters()):
        p_.requires_grad_(True)
    torch.manual_seed(61)
    o = G(z)
    torch.manual_seed(61)
    u = torch.rand(B, 1, 32, 64)
    y = D(o["image"])
    loss = torch.nn.functional.softplus(-y).mean()
    namesG = [n for n, _ in G.named_parameters()]
    namesD = [n for n, _ in D.named_parameters()]
    grads = torch.autograd.grad(loss, [z] + list(G.parameters()) + list(D.parameters()),
                                allow_unused=True)
    out.update(z=npy(z), u=npy(u), image=npy(o["image"]), image_orig=npy(o["image_orig"]),
               raydrop_logit=npy(o["raydrop_logit"]), raydrop_mask=npy(o["raydrop_mask"]),
               y=npy(y), gz=npy(grads[0]))
    for n, g in zip(namesG, grads[1:1 + len(namesG)]):
        if g is not None:
            out[f"gG_{n}"] = npy(g)
    for n, g in zip(namesD, grads[1 + len(namesG):]):
        if g is not None:
            out[f"gD_{n}"] = npy(g)
    save("vanilla_small.npz", **out)


if __name__ == "__main__":
    if "--vanilla-only" in sys.argv:
        golden_vanilla()
        sys.exit(0)
    golden_ada()
    if "--ada-only" in sys.argv:
        sys.exit(0)
    golden_ops()
    golden_coords()
    golden_generator()
    golden_discriminator()
    golden_vanilla()
