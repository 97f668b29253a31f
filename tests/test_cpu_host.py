"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares,
ops refuse CPU tensors (no fallback), host-side logic (configs, state_dict layout, ADA
sampling, padding geometry) behaves, and the oracle never leaks into the product."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dusty_gan_v2_b200 import _cabi
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "dusty_b200.h")).read()
    declared = set(re.findall(r"\b(dusty_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    assert lib.dusty_abi_version() == _cabi.ABI_VERSION


def test_no_cpu_fallback():
    import dusty_gan_v2_b200.functional as DF
    from dusty_gan_v2_b200.gans.models import ops
    from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d
    x = torch.randn(2, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        DF.bias_act(x, torch.zeros(3))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.Resample(up=2)(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        upfirdn2d(x, torch.ones(2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.MinibatchStdDev()(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.FourierFeature((8, 8), num_freqs=8)(torch.zeros(1, 2, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dusty_gan_v2_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), os.path.join(d, f)
    code = ("import sys; sys.path.insert(0, %r); import dusty_gan_v2_b200, "
            "dusty_gan_v2_b200.gans.trainer; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_missing_library_fails_loudly(monkeypatch):
    from dusty_gan_v2_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libdusty_b200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        _cabi.load()


def test_state_dict_layout_matches_reference(g_gen, g_disc):
    from dusty_gan_v2_b200.gans.models.builder import build_discriminator, build_generator
    from small_cfgs import D_SMALL, G_SMALL
    G, D = build_generator(G_SMALL), build_discriminator(D_SMALL)
    ref_g = {k[3:]: v.shape for k, v in g_gen.items() if k.startswith("sd_")}
    ref_d = {k[3:]: v.shape for k, v in g_disc.items() if k.startswith("sd_")}
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == ref_g
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == ref_d


def test_full_size_parameter_counts_and_presets():
    from dusty_gan_v2_b200.gans.models.builder import build_discriminator, build_generator
    from dusty_gan_v2_b200.presets import preset
    expect = {"dusty_v2": (4367594, 38445569), "dusty_v1": (36309954, 2821057),
              "vanilla": (36308929, 2821057)}
    for arch, (ng, nd) in expect.items():
        cfg = preset(arch)
        G, D = build_generator(cfg.model.generator), build_discriminator(cfg.model.discriminator)
        assert sum(p.numel() for p in G.parameters()) == ng
        assert sum(p.numel() for p in D.parameters()) == nd


def test_fourier_init_consumes_rng_like_reference(g_ops):
    from dusty_gan_v2_b200.gans.models.ops import FourierFeature
    torch.manual_seed(11)
    np.random.seed(11)
    ff = FourierFeature(resolution=(8, 16), num_freqs=32)
    assert np.array_equal(ff.freqs.numpy(), g_ops["ff_freqs"])
    assert np.array_equal(ff.phase.numpy(), g_ops["ff_phase"])
    assert [ff.L_h, ff.L_w] == g_ops["ff_L"].tolist()


def test_coord_bridge_angle_bit_exact_on_host(g_coords):
    from dusty_gan_v2_b200.gans.coords import CoordBridge
    cb = CoordBridge(64, 512, 1.45, 80.0, os.path.join(ROOT, "data/coords/kitti_raw.npy"))
    assert np.array_equal(cb.angle.numpy(), g_coords["angle"])
    depth = torch.from_numpy(g_coords["depth"])
    mask = torch.from_numpy(g_coords["mask"])
    x = cb.convert(depth, "depth", "inv_depth_norm") * 2 - 1
    assert np.array_equal((mask * x + (1 - mask) * -1.0).numpy(), g_coords["reals"])
    ps = cb.convert(torch.arange(24.).reshape(1, 3, 2, 4), "point_map", "point_set")
    assert ps.shape == (1, 8, 3) and float(ps[0, 5, 1]) == 8 + 5     # index h*W + w


def test_ada_host_sampling_and_padding():
    from dusty_gan_v2_b200.gans.augment import adaptive_augment as A
    ada = A.AdaptiveAugment(p_init=0.0, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1, frac_trans=1,
                            brightness=1, contrast=1, luma_flip=1, hue=1, saturation=1)
    ada.generator = torch.Generator().manual_seed(0)
    G = ada.sample_affine(5, 64, 512)
    C = ada.sample_color(5)
    assert torch.equal(G, torch.eye(3).repeat(5, 1, 1)) and torch.equal(C, torch.eye(4).repeat(5, 1, 1))
    # identity transform needs only the filter margin (76 x 524 at p = 0)
    px1, px2, py1, py2 = A.padding_for(torch.inverse(G), 64, 512, 12)
    assert (64 + py1 + py2, 512 + px1 + px2) == (76, 524)
    ada._p_host = 0.9
    G = ada.sample_affine(64, 64, 512)
    assert G.shape == (64, 3, 3) and float((G - torch.eye(3)).abs().sum()) > 0
    assert torch.allclose(G[:, 2], torch.tensor([0.0, 0.0, 1.0]).expand(64, 3))
    pads = A.padding_for(torch.inverse(G), 64, 512, 12)
    assert all(0 <= p for p in pads) and pads[0] <= 511 and pads[2] <= 63
    assert tuple(ada.Hz_fbank.shape) == (4, 43)


def test_ada_fused_parameter_rows_and_policy_vector():
    """Host side of the fused ADA op (adaptive_augment.py:271-291,386-545 as two launches): the
    [B, 8] parameter rows carry the axis-aligned inverse transform and the one-channel colour
    gain / offset of the sampled matrices; anything with shear / rotation is refused (None ->
    the single-op path); the policy vector the device sampler reads lists the eleven multipliers
    in the kernel's order."""
    from dusty_gan_v2_b200.gans.augment import adaptive_augment as A
    ada = A.AdaptiveAugment(p_init=0.0, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1, frac_trans=1,
                            brightness=1, contrast=1, luma_flip=1, hue=1, saturation=1)
    ada.generator = torch.Generator().manual_seed(3)
    ada._p_host = 0.8
    G_inv = torch.inverse(ada.sample_affine(16, 64, 512))
    C = ada.sample_color(16)
    rows = A.AdaptiveAugment.fused_params(G_inv, C)
    assert rows.shape == (16, 8)
    # x' = a x + tx, y' = d y + ty of the inverse map
    rebuilt = torch.eye(3).repeat(16, 1, 1)
    rebuilt[:, 0, 0], rebuilt[:, 0, 2], rebuilt[:, 1, 1], rebuilt[:, 1, 2] = rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3]
    assert torch.allclose(rebuilt, G_inv, atol=1e-6)
    # one-channel images: the colour matrix collapses to a gain and an offset (mean over the RGB rows)
    x = torch.rand(16, 1, 7)
    full = (C[:, :3, :3] @ x.expand(16, 3, 7) + C[:, :3, 3:]).mean(dim=1, keepdim=True)
    assert torch.allclose(x * rows[:, 4].view(-1, 1, 1) + rows[:, 5].view(-1, 1, 1), full, atol=1e-5)
    assert torch.equal(rows[:, 6:], torch.zeros(16, 2))
    sheared = G_inv.clone()
    sheared[3, 0, 1] = 0.1
    assert A.AdaptiveAugment.fused_params(sheared, C) is None
    pv = ada.policy_vector()
    assert len(pv) == 11 and pv[:10] == [1] * 10 and pv[10] == ada.h_trans_factor


def test_split_bf16x3_scheme_error_bound():
    """The arithmetic behind fp32 mode on the tensor cores (csrc/split3.cu), restated
    on the host: a = a_hi + a_lo in bf16, the K axis tripled as [a_hi|a_hi|a_lo] . [b_hi|b_lo|b_hi]
    with fp32 accumulation reproduces the fp32 dot product to ~2^-16 relative to sum |a||b| --
    two orders below the rtol 1e-3 the mode is held to -- while plain bf16 operands do not."""
    g = torch.Generator().manual_seed(11)
    a = torch.randn(64, 1536, generator=g)
    b = torch.randn(1536, 48, generator=g)

    def split(t):
        hi = t.bfloat16()
        lo = (t - hi.float()).bfloat16()
        return hi.float(), lo.float()

    ah, al = split(a)
    bh, bl = split(b)
    a3 = torch.cat([ah, ah, al], dim=1)                   # pattern 0 (the "a" side)
    b3 = torch.cat([bh, bl, bh], dim=0)                   # pattern 1 (the "b" side)
    ref = a.double() @ b.double()
    scale = (a.abs().double() @ b.abs().double())
    err3 = ((a3 @ b3).double() - ref).abs() / scale
    err1 = ((ah @ bh).double() - ref).abs() / scale
    assert float(err3.max()) < 4e-5, float(err3.max())    # dropped lo*lo + residuals: ~2^-16
    assert float(err1.max()) > 20 * float(err3.max())      # bf16 operands alone: ~2^-9


def test_fir_geometry_matches_oracle_sizes():
    from dusty_gan_v2_b200.functional import FirCfg
    from oracle import dusty_oracle as O
    for up, down, pad, k in [((2, 1), (1, 1), (0, 0, 6, 5), (1, 12)), ((1, 1), (2, 3), (1, 2, -1, 3), (4, 5))]:
        cfg = FirCfg(k[0], k[1], 1, up=up, down=down, pad=pad)
        oh, ow = cfg.out_hw(17, 23)
        assert oh == O.upfirdn2d_out_size(17, up[0], down[0], pad[0], pad[1], k[0])
        assert ow == O.upfirdn2d_out_size(23, up[1], down[1], pad[2], pad[3], k[1])


def test_c_abi_argument_errors_are_status_codes_not_crashes():
    """Error behaviour of the C ABI (no GPU needed: arguments are validated before any CUDA
    call): bad arguments return a negative status and set dusty_last_error()."""
    from dusty_gan_v2_b200 import _cabi as K
    lib = K.load()
    one = torch.zeros(64)
    p = one.data_ptr()
    rc = lib.dusty_stem_fwd(p, p, None, p, 1, 8, 8, 12, 0.25, 0.5, 0.25, 0.2, 1.41, K.F32, None)   # O = 12
    assert rc != 0 and "O must be" in K.last_error()
    rc = lib.dusty_weight_prep(None, p, None, 4, 4, 9, 1.0, K.BF16, None)
    assert rc != 0 and "null" in K.last_error()
    rc = lib.dusty_bias_act_add_cl(p, None, p, p, 64, 12, 0.2, 1.41, 0.7, K.BF16, None)           # C % 8 != 0
    assert rc != 0 and "channel" in K.last_error()
    rc = lib.dusty_modconv_fwd(p, p, p, None, p, 2, 32, 64, 0, 1, 128, 3, 0.2, 1.0, K.BF16, K.BF16, 7, None, None, None, None)
    assert rc != 0 and "impl" in K.last_error()
    rc = lib.dusty_filter_rsco_to_ohwi(p, p, 0, 4, 9, K.BF16, None)
    assert rc != 0 and "shape" in K.last_error()
    with pytest.raises(RuntimeError, match="status"):
        K.call("dusty_pad2d_cl", p, p, 1, 8, 8, 8, 9, 0, 0, 0, K.PAD_REPLICATE, K.PAD_CIRCULAR, 0, K.BF16, None)
    # entries added late in round 1
    rc = lib.dusty_residual_fork_bwd_cl(p, None, p, 0.125, 0.375, 0.375, 0.125, 1, 8, 8, 8, K.BF16, None)
    assert rc != 0 and "null" in K.last_error()
    rc = lib.dusty_residual_fork_bwd_cl(p, p, p, 0.125, 0.375, 0.375, 0.125, 1, 7, 8, 8, K.BF16, None)     # odd H
    assert rc != 0 and "shape" in K.last_error()
    rc = lib.dusty_residual_fork_bwd_cl(p, p, p, 0.125, 0.375, 0.375, 0.125, 1, 8, 8, 12, K.BF16, None)    # C % 8
    assert rc != 0 and "vector" in K.last_error()
    rc = lib.dusty_up2_sumsq(p, p, None, 0.25, 0.75, 0.75, 0.25, 1, 8, 8, K.BF16, None)
    assert rc != 0 and "null" in K.last_error()
    rc = lib.dusty_up2_sumsq(p, p, p, 0.25, 0.75, 0.75, 0.25, 1, 8, 12, K.BF16, None)                      # W % 8
    assert rc != 0 and "multiple" in K.last_error()
    rc = lib.dusty_fir1d(p, p, p, 65, 1, 1, 8, 8, 1, 2, 1, 6, 5, None)                                     # > 64 taps
    assert rc != 0 and "tap count" in K.last_error()


def test_inversion_host_logic_and_no_cpu_fallback():
    """gans/inversion.py mirror: the learning-rate schedule is the oracle's (demo_inversion.py:140-146),
    the losses and the loop refuse CPU tensors, SphericalOptimizer / geocross are plain torch."""
    from dusty_gan_v2_b200.gans import inversion as inv
    from oracle import dusty_oracle as O
    for n in (3, 13, 500):
        for i in range(0, n, max(1, n // 7)):
            assert inv.lr_schedule(i, n) == pytest.approx(O.inversion_lr_schedule(i, n), abs=1e-15)
    assert inv.lr_schedule(0, 500) == 0.0 and inv.lr_schedule(25, 500) == pytest.approx(1.0)
    x = torch.rand(2, 1, 16, 64) + 0.1
    with pytest.raises(RuntimeError, match="CUDA"):
        inv.MultiScaleMaskedLoss(torch.nn.functional.l1_loss, level=2)(x, x, torch.ones_like(x))
    with pytest.raises(RuntimeError, match="CUDA"):
        inv.LatentInversion(None, None, x, torch.ones_like(x))
    with pytest.raises(ValueError):
        inv.LatentInversion(None, None, x, torch.ones_like(x), latent_type="q")
    lat = torch.randn(2, 10, 16)
    assert torch.allclose(inv.geocross_loss(lat), O.geocross_loss(lat))
    assert torch.equal(inv.tanh_to_sigmoid(torch.tensor([-1.0, 0.0, 1.0])), torch.tensor([0.0, 0.5, 1.0]))


def test_reference_extension_recipe():
    """oracle/build_ref.py: compiles the reference's sources where they lie (no copy in the repo),
    outputs only under oracle/_ref (git-ignored, not gpurun-ignored), loader returns None when a
    module is not there."""
    from oracle import build_ref
    assert build_ref.load_built("no_such_module") is None
    for srcs in build_ref.MODULES.values():
        for f in srcs:
            assert not os.path.abspath(f).startswith(ROOT + os.sep), f
    assert build_ref.OUT == os.path.join(ROOT, "oracle", "_ref")
    ignore = open(os.path.join(ROOT, ".gitignore")).read().split()
    assert "oracle/_ref/" in ignore
    gpurunignore = os.path.join(ROOT, ".gpurunignore")
    if os.path.exists(gpurunignore):
        assert "oracle/_ref" not in open(gpurunignore).read()
    # no reference source file was copied into the tree
    for d, _, files in os.walk(ROOT):
        if ".git" in d.split(os.sep) or os.sep + "_ref" in d:
            continue
        assert not {"fused_bias_act_kernel.cu", "upfirdn2d_kernel.cu"} & set(files), d


def test_ncu_window_summary_tool():
    csv_path = os.path.join(ROOT, "profiles", "r01_ncu_launches_window.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_window_summary.py"), csv_path],
                         check=True, capture_output=True, text=True).stdout
    head = out.splitlines()[0]
    assert "2500 launches" in head and "dusty:: kernels" in head


def test_bench_reference_legs_on_cpu(g_gen, g_invloop):
    """bench.py is the one measurement entry that executes oracle/: its kernel-level reference arm
    reports `unavailable` without a GPU (never raises into the bench line), the config-5 CPU leg
    runs the oracle's inversion step."""
    import bench
    res = bench.reference_kernel_baseline()
    assert "unavailable" in res
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--ref-kernels-only"],
                        capture_output=True, text=True, timeout=300, cwd=ROOT)
    import json
    assert "unavailable" in json.loads(cp.stdout.strip().splitlines()[-1])
    sd = {k[3:]: torch.from_numpy(v) for k, v in g_gen.items() if k.startswith("sd_")}
    T = torch.from_numpy
    out = bench.inversion_cpu_baseline(sd, T(g_invloop["w+_z0"]), T(g_invloop["angle"]), T(g_invloop["depth"]),
                                       T(g_invloop["mask"]), "w+")
    assert out["batch"] == 3 and out["target_iterations_per_s"] > 0 and out["kind"] == "port"


def test_utils_mirror_host_helpers():
    """gans/utils.py mirror: the sampler's stream is reproducible and rank-disjoint, the
    visualisation helpers refuse, and the module is registered by install_as_gans."""
    import dusty_gan_v2_b200 as pkg
    from dusty_gan_v2_b200.gans import utils as U
    assert "gans.utils" in pkg._MIRRORED and "gans.inversion" in pkg._MIRRORED
    data = list(range(10))
    streams = []
    for rank in range(2):
        it = iter(U.InfiniteSampler(data, rank=rank, num_replicas=2, seed=4))
        streams.append([int(next(it)) for _ in range(40)])
    it = iter(U.InfiniteSampler(data, rank=0, num_replicas=2, seed=4))
    assert streams[0] == [int(next(it)) for _ in range(40)]
    assert streams[0] != streams[1]
    it = iter(U.InfiniteSampler(data, shuffle=False))
    assert [int(next(it)) for _ in range(12)] == list(range(10)) + [0, 1]
    with pytest.raises(NotImplementedError):
        U.colorize(torch.zeros(1))
    lin = torch.nn.Linear(2, 2)
    U.set_requires_grad(lin, False)
    assert not any(p.requires_grad for p in lin.parameters())
    assert torch.equal(U.sigmoid_to_tanh(U.tanh_to_sigmoid(torch.tensor([-1.0, 0.25]))), torch.tensor([-1.0, 0.25]))


def test_install_as_gans_resolves_reference_import_lines():
    """after install_as_gans() the reference's own import lines
    (trainer.py:13-27, demo_inversion.py:11-28) resolve to the mirror."""
    code = "\n".join([
        "import sys; sys.path.insert(0, %r)" % ROOT,
        "import dusty_gan_v2_b200 as b200; b200.install_as_gans()",
        "import gans.models.ops as ops",
        "from gans.coords import CoordBridge",
        "from gans.augment.adaptive_augment import AdaptiveAugment",
        "from gans.inversion import MultiScaleMaskedLoss, SphericalOptimizer, geocross_loss, normalize_noise_",
        "from gans.models.builder import build_discriminator, build_generator",
        "from gans.models.loss import GANLoss",
        "from gans.models.ops.common import filter2d",
        "from gans.utils import InfiniteSampler, set_requires_grad, sigmoid_to_tanh, tanh_to_sigmoid, "
        "init_random_seed, cycle",
        "from gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d",
        "from gans.models.ops.fused_act.fused_act import FusedLeakyReLU, fused_leaky_relu",
        "assert ops.ModConv2d.__module__.startswith('dusty_gan_v2_b200')",
        "assert CoordBridge.__module__.startswith('dusty_gan_v2_b200')",
    ])
    subprocess.run([sys.executable, "-c", code], check=True, timeout=300)


def test_ada_controller_matches_reference_trainer_step(g_step):
    """ADA's probability controller (adaptive_augment.py:368-384, host logic of the mirror) against
    what the reference's real Trainer.step did in the recorded iteration: rt = mean sign of D(real),
    p moves by sign(rt - p_target) * n_pred / (kimg * 1000)."""
    from dusty_gan_v2_b200.gans.augment.adaptive_augment import AdaptiveAugment
    ada = AdaptiveAugment(p_init=0.5, p_target=0.6, kimg=500, lr_flip=1, ud_flip=1, int_trans=1, iso_scale=1,
                          frac_trans=1, brightness=1, contrast=1, luma_flip=1, hue=1, saturation=1)
    rt_ref = float(g_step["ada_rt"])
    n_pos = int(round((rt_ref + 1) / 2 * 4))
    y_real = torch.tensor([1.0] * n_pos + [-1.0] * (4 - n_pos)).reshape(4, 1)
    ada.cumulate(y_real)
    rt = ada.update_p()
    assert float(rt) == pytest.approx(rt_ref)
    assert float(ada.p) == pytest.approx(float(g_step["ada_p_after"].reshape(-1)[0]), abs=1e-7)
    assert float(ada.sign_cum) == 0.0 and float(ada.n_pred_cum) == 0.0


def test_trainer_checkpoint_round_trip_on_host():
    """`Trainer.state_dict(step)` (the reference's checkpoint keys, trainer.py:551-567) ->
    `Trainer.load_state_dict` (resume, trainer.py:184-196): weights, EMA copy, ADA state, both Adam
    states and the start iteration survive; host-only (construction needs no GPU)."""
    from small_cfgs import D_SMALL, G_SMALL
    from dusty_gan_v2_b200.config import to_attr
    from dusty_gan_v2_b200.gans.trainer import Trainer
    from dusty_gan_v2_b200.presets import preset

    def make(seed):
        torch.manual_seed(seed)
        np.random.seed(seed)
        cfg = preset("dusty_v2", batch_size=4)
        cfg.model.generator, cfg.model.discriminator = to_attr(G_SMALL), to_attr(D_SMALL)
        return Trainer(cfg, iter([]), device="cpu", precision="fp32",
                       angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))

    a, b = make(1), make(2)
    for opt, net in ((a.optim_G, a.G_module), (a.optim_D, a.D_module)):      # give Adam some state
        for p in net.parameters():              # (the update itself is a CUDA kernel: GPU tests)
            st = opt._state_of(p)
            st["step"] += 3
            st["exp_avg"].normal_()
            st["exp_avg_sq"].uniform_()
        with pytest.raises(RuntimeError):       # no CPU fallback
            next(iter(net.parameters())).grad = torch.zeros_like(next(iter(net.parameters())))
            opt.step()
        for p in net.parameters():
            p.grad = None
    a.A.p.fill_(0.25)
    payload = a.state_dict(step=7 * 4)
    assert set(payload) == {"cfg", "step", "angle", "G", "D", "G_ema", "A", "optim_G", "optim_D"}
    assert b.load_state_dict(payload) == 7
    for x, y in ((a.G_module, b.G_module), (a.D_module, b.D_module), (a.G_ema, b.G_ema), (a.A, b.A)):
        for (k, u), v in zip(x.state_dict().items(), y.state_dict().values()):
            assert torch.equal(u, v), k
    for oa, na, ob, nb in ((a.optim_G, a.G_module, b.optim_G, b.G_module), (a.optim_D, a.D_module, b.optim_D, b.D_module)):
        for pa, pb in zip(na.parameters(), nb.parameters()):
            assert torch.equal(oa.state[pa]["exp_avg"], ob.state[pb]["exp_avg"])
            assert torch.equal(oa.state[pa]["exp_avg_sq"], ob.state[pb]["exp_avg_sq"])
