"""Shared helpers of the step-level parity tests: the fixtures written by
tests/golden/make_golden_trainer_step.py hold ONE FULL ITERATION of the reference's real
`Trainer.step` (recorded random draws, losses, the gradients each optimiser consumed, the
updated weights).  Here: comparison against a (possibly compact) fixture, the CPU oracle advanced
through the same three phases, and the replay of the recorded draws through the mirror
`Trainer.step` on the GPU (eager or CUDA-graphed)."""
import numpy as np
import torch

from small_cfgs import sample_flat

T = torch.from_numpy
NOGRAD_G = ("ema_var", "w_avg", "kernel", "pe.", "raydrop_const")


def check(got, g, key, rtol, atol_rel, what=""):
    """`got` against fixture entry `key`; compact fixtures store a fixed-stride sample of the
    tensor plus its L2 norm (key 'n' + key)."""
    ref = np.asarray(g[key])
    got = got.detach().float().cpu()
    compact = ("n" + key) in g
    a = sample_flat(got).numpy() if compact else got.numpy().reshape(ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-12)
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=atol_rel * scale, err_msg=f"{what}{key}")
    if compact:
        n_ref = float(g["n" + key])
        assert abs(float(got.double().norm()) - n_ref) <= rtol * n_ref + 1e-12, (key, float(got.norm()), n_ref)


def check_grads(named_grads, g, prefix, rtol, atol_rel, min_n):
    n = 0
    for k, gr in named_grads.items():
        if gr is None or prefix + k not in g:
            continue
        check(gr, g, prefix + k, rtol, atol_rel)
        n += 1
    assert n >= min_n, (prefix, n)
    return n


def check_updated_weights(named, g, prefix, lr, min_total=1000, frac=0.98):
    """Weights after the first Adam step: the update is lr * g / (|g| + eps), so an entry whose
    gradient is ~0 may move the other way -- hence a fraction, with 10 % of the step as the bar."""
    near = total = 0
    for k, v in named.items():
        key = prefix + k
        if key not in g:
            continue
        ref = T(np.asarray(g[key]))
        got = v.detach().float().cpu()
        got = sample_flat(got) if ("n" + key) in g else got.reshape(ref.shape)
        d = (got - ref).abs()
        near += int((d < 0.1 * lr).sum())
        total += d.numel()
    assert total >= min_total and near / total > frac, (prefix, near, total)


def fixture_draws(g):
    """The recorded random draws of a fixture as tensors (affine transforms as sampled)."""
    d = {k: T(np.asarray(g[k])) for k in ("z_g", "z_d", "shift_g", "shift_d", "u_g", "u_d") if k in g}
    for tag in ("g_fake", "d_real", "d_fake", "r1"):
        for k in ("keep", "G", "C"):
            d[f"{k}_{tag}"] = T(np.asarray(g[f"{k}_{tag}"]))
    return d


def oracle_rnd(draws):
    """Draws in the form `O.train_iteration` takes (inverse affine transforms)."""
    rnd = {k: v for k, v in draws.items() if not k.startswith("G_")}
    for tag in ("g_fake", "d_real", "d_fake", "r1"):
        rnd[f"Ginv_{tag}"] = torch.inverse(draws[f"G_{tag}"])
    return rnd


def oracle_iteration(O, sdG, sdD, x_real, angle, rnd, lr=0.002, with_r1=True):
    """The reference's iteration restated on the oracle: G step, D step on the updated
    generator, lazy R1 step on the updated discriminator (Adam settings of trainer.py:128-171).
    `sdG` / `sdD` are updated in place; returns losses and the three gradient sets."""
    sdG = {k: v.requires_grad_(v.dtype.is_floating_point and not any(t in k for t in NOGRAD_G))
           for k, v in sdG.items()}
    sdD = {k: v.requires_grad_("kernel" not in k) for k, v in sdD.items()}
    lazy = 16 / 17.0
    optG = torch.optim.Adam([v for v in sdG.values() if v.requires_grad], lr=lr, betas=(0.0, 0.99))
    optD = torch.optim.Adam([v for v in sdD.values() if v.requires_grad], lr=lr * lazy,
                            betas=(0.0, 0.99 ** lazy))

    def apply(opt, sd, grads):
        for k, gr in grads.items():
            sd[k].grad = gr
        opt.step()
        opt.zero_grad(set_to_none=True)

    out = {}
    r = O.train_iteration(sdG, sdD, x_real, angle, rnd, with_r1=False)
    out["loss_G"], out["grads_G"] = r["loss_G"], r["grads_G"]
    nb = {}
    with torch.no_grad():
        O.generator(sdG, rnd["z_g"], angle, rnd["u_g"], training=True,
                    shifts_rad=rnd["shift_g"] * (2 * np.pi), new_buffers=nb)
    apply(optG, sdG, r["grads_G"])
    for k, v in nb.items():
        sdG[k] = v.detach().clone()
    r = O.train_iteration(sdG, sdD, x_real, angle, rnd, with_r1=False)
    out["loss_D"], out["grads_D"] = r["loss_D"], r["grads_D"]
    apply(optD, sdD, r["grads_D"])
    if with_r1:
        r = O.train_iteration(sdG, sdD, x_real, angle, rnd, with_r1=True)
        out["r1"], out["grads_R1"] = r["r1"], r["grads_R1"]
        apply(optD, sdD, r["grads_R1"])
    out["sdG"], out["sdD"] = sdG, sdD
    return out


class MirrorReplay:
    """Feeds recorded random draws to the mirror `Trainer.step`.  The mirror stacks real + fake
    in the D step, so its single dropout / ADA draw of 2B samples is the concatenation of the
    reference's two.  Draws that the CUDA-graphed trainer consumes INSIDE a captured segment
    (latents, aug-coords shifts, Gumbel uniforms) are served as clones of device-resident
    tensors selected by the phase of the step (graph-safe: a device-to-device copy node), so
    warm-up passes and the capture may ask for them any number of times."""

    def __init__(self, tr, draws, monkeypatch, dev):
        rnd = draws
        self.tr, self.dev = tr, dev
        self.phase = "g"
        dv = lambda k: rnd[k].to(dev) if k in rnd else None      # noqa: E731
        self.z = {"g": dv("z_g"), "d": dv("z_d")}
        self.u = {"g": dv("u_g"), "d": dv("u_d")}
        self.shift = {"g": dv("shift_g"), "d": dv("shift_d")}
        me = self

        tr.sample_z = lambda n: me.z[me.phase].clone()
        monkeypatch.setattr(torch, "rand", lambda *a, **k: me.u[me.phase].clone())
        monkeypatch.setattr(torch.Tensor, "uniform_",
                            lambda self_, a=0, b=1, **k: self_.copy_(me.shift[me.phase]), raising=True)
        fake_nograd = tr._fake_images_nograd

        def d_phase(B):
            me.phase = "d"
            return fake_nograd(B)
        tr._fake_images_nograd = d_phase

        def seq(*tensors):
            it = iter(tensors)
            return lambda *a, **k: next(it)

        keeps = seq(rnd["keep_g_fake"], torch.cat([rnd["keep_d_real"], rnd["keep_d_fake"]]), rnd["keep_r1"])
        monkeypatch.setattr(torch, "bernoulli", lambda p, **k: keeps().to(p.device))
        affines = seq(rnd["G_g_fake"], torch.cat([rnd["G_d_real"], rnd["G_d_fake"]]), rnd["G_r1"])
        colors = seq(rnd["C_g_fake"], torch.cat([rnd["C_d_real"], rnd["C_d_fake"]]), rnd["C_r1"])
        tr.A.sample_affine = lambda *a, **k: affines()
        tr.A.sample_color = lambda *a, **k: colors()

        # gradients each optimiser step consumed / weights right after it
        self.d_grads, self.g_grads = [], {}
        d_step, g_step = tr.optim_D.step, tr.optim_G.step

        def rec_d(*a, **k):
            me.d_grads.append({n: p.grad.detach().clone() for n, p in tr.D_module.named_parameters()
                               if p.grad is not None})
            return d_step(*a, **k)

        def rec_g(*a, **k):
            me.g_grads = {n: p.grad.detach().clone() for n, p in tr.G_module.named_parameters()
                          if p.grad is not None}
            return g_step(*a, **k)
        tr.optim_D.step, tr.optim_G.step = rec_d, rec_g

    def run(self, iteration=0):
        self.phase = "g"
        return self.tr.scalars_to_host(self.tr.step(iteration))
