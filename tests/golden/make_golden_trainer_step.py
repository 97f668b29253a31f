"""Golden vectors of ONE FULL TRAINING ITERATION produced by the reference's real
`gans.trainer.Trainer.step` (gans/trainer.py:247-482) on CPU: the trainer object is assembled by
tests/ref_trainer_harness.py (no `__init__`: that needs a CUDA rank and the KITTI files; one-rank
gloo DDP), small G / D, iteration 0 (G step, D step, lazy R1 step, ADA p = 0.5, warm-up dropout
0.5).  Every random draw of the step is recorded and stored, together with the losses and the
gradients each optimiser step consumed.

    python tests/golden/make_golden_trainer_step.py      ->  tests/golden/trainer_step.npz
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ref_trainer_harness import build_reference_trainer, record_step  # noqa: E402
from small_cfgs import D_MID, D_SMALL, G_MID, G_SMALL, V1_SMALL, V_SMALL, VD_SMALL, sample_flat  # noqa: E402

npy = lambda t: t.detach().cpu().numpy().copy()  # noqa: E731


def main():
    dist.init_process_group("gloo", init_method=f"file://{tempfile.mkdtemp()}/pg", rank=0, world_size=1)
    one(G_SMALL, D_SMALL, (16, 64), "trainer_step.npz", 1100)
    one(V1_SMALL, VD_SMALL, (32, 64), "trainer_step_dusty_v1.npz", 1200)       # BASELINE config 3
    one(V_SMALL, VD_SMALL, (32, 64), "trainer_step_vanilla.npz", 1300)
    # channel counts that qualify for the tcgen05 kernels (bf16 / CUDA-graph twin of the replay
    # test); gradients and updated weights stored as fixed-stride samples + norms to stay small
    one(G_MID, D_MID, (32, 128), "trainer_step_mid.npz", 1400, compact=True, logit_gain=True)
    dist.destroy_process_group()


def one(g_cfg, d_cfg, res, fname, seed, compact=False, logit_gain=False):
    B, (H, W) = 4, res
    v2 = g_cfg["arch"] == "dusty_v2"
    torch.manual_seed(seed)
    np.random.seed(seed)
    g = torch.Generator().manual_seed(seed + 1)
    batch = {"depth": 1.45 + 78.55 * torch.rand(B, 1, H, W, generator=g),
             "mask": (torch.rand(B, 1, H, W, generator=g) < 0.85).float()}
    T, G, D = build_reference_trainer(g_cfg, d_cfg, B, (H, W), [batch], p_init=0.5)
    with torch.no_grad():                         # de-trivialise the zero-initialised biases
        for net in (G, D):
            for n, p in net.named_parameters():
                if "bias" in n:
                    p.normal_(0, 0.2)
    if logit_gain:
        # a state with O(1) logits (random-init logits are ~0.05: a bf16 comparison of them would
        # measure rounding noise): widen D's last linear until std(D(x)) ~ 1 on the real batch
        with torch.no_grad():
            x = T.fetch_reals(batch)["image"]
            last = D.epilogue[-1].module
            s0 = float(D(x).std())
            last.weight.mul_(1.0 / max(s0, 1e-6))
            print(f"{fname}: logit std {s0:.4f} -> {float(D(x).std()):.4f}")
    out = {f"sdG_{k}": npy(v) for k, v in G.state_dict().items()}
    out.update({f"sdD_{k}": npy(v) for k, v in D.state_dict().items()})
    scalars, log, g_grads, d_grads = record_step(T, G, D, 0)
    gumbel = g_cfg["arch"] != "vanilla"
    assert [len(log[k]) for k in ("randn", "uniform_", "rand", "bernoulli", "affine", "color")] == \
        [2, 2 if v2 else 0, 2 if gumbel else 0, 4, 4, 4]
    out.update(depth=npy(batch["depth"]), mask=npy(batch["mask"]), z_g=npy(log["randn"][0]),
               z_d=npy(log["randn"][1]))
    if gumbel:
        out.update(u_g=npy(log["rand"][0]), u_d=npy(log["rand"][1]))
    if v2:
        out.update(angle=npy(T.auxin["angle"]), shift_g=npy(log["uniform_"][0]), shift_d=npy(log["uniform_"][1]))
    for i, tag in enumerate(("g_fake", "d_real", "d_fake", "r1")):
        out[f"keep_{tag}"] = npy(log["bernoulli"][i])
        out[f"G_{tag}"] = npy(log["affine"][i])
        out[f"C_{tag}"] = npy(log["color"][i])
    out.update(loss_G=np.array(scalars["loss/G/adversarial"]), loss_D=np.array(scalars["loss/D/adversarial"]),
               r1=np.array(scalars["loss/D/gradient_penalty"]), ada_rt=np.array(scalars["stats/ada_rt"]),
               ada_p_after=npy(T.A.p), ema_decay=np.array(scalars["stats/ema_decay"]))
    def put(prefix, items):
        for k, v in items:
            if compact:
                out[f"{prefix}_{k}"] = npy(sample_flat(v))
                out[f"n{prefix}_{k}"] = np.array(float(v.detach().double().norm()))
            else:
                out[f"{prefix}_{k}"] = npy(v)

    put("gG", g_grads.items())
    put("gD", d_grads[0].items())
    put("gR1", d_grads[1].items())
    put("afterG", [(k, v) for k, v in G.state_dict().items() if "kernel" not in k and "pe." not in k])
    put("afterD", [(k, v) for k, v in D.state_dict().items() if "kernel" not in k])
    out.update({f"afterGema_{k}": npy(v) for k, v in T.G_ema.state_dict().items() if k.endswith("ema_var") or k == "w_avg"})
    path = os.path.join(HERE, fname)
    np.savez_compressed(path, **out)
    print(f"{fname}: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


if __name__ == "__main__":
    main()
