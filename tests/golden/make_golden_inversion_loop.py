"""Golden vectors for the latent-optimisation loop of BASELINE config 5 (demo_inversion.py
89-217), produced with the UNMODIFIED reference's own parts on CPU: its dusty_v2 Generator (the
small configuration and weights of g_small.npz), CoordBridge, MultiScaleMaskedLoss,
geocross_loss, tanh_to_sigmoid, torch.optim.Adam + LambdaLR.  demo_inversion.py keeps its loop
inside `if __name__ == "__main__":` (needs a checkpoint and KITTI files), so the loop body is
re-assembled here statement by statement from those lines; everything it calls is the reference.

    python tests/golden/make_golden_inversion_loop.py     ->  tests/golden/inversion_loop.npz
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_import  # noqa: E402

ref_import.install()

from gans.coords import CoordBridge  # noqa: E402
from gans.inversion import MultiScaleMaskedLoss, geocross_loss  # noqa: E402
from gans.models.builder import build_generator  # noqa: E402
from gans.utils import set_requires_grad, tanh_to_sigmoid  # noqa: E402
from small_cfgs import G_SMALL  # noqa: E402

torch.set_num_threads(4)
npy = lambda t: t.detach().cpu().numpy().copy()  # noqa: E731

H, W, MIN_D, MAX_D = 16, 64, 1.45, 80.0
STEPS, LR, RAMPUP, RAMPDOWN, NUM_Z = 3, 5e-2, 0.05, 0.25, 256


def main():
    gold = np.load(os.path.join(HERE, "g_small.npz"))
    torch.manual_seed(50)
    np.random.seed(50)
    G = build_generator(ref_import.to_attr(G_SMALL))
    G.load_state_dict({k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("sd_")})
    G.eval()
    angle_file = os.path.join(ROOT, "data/coords/kitti_raw.npy")
    coord = CoordBridge(num_ring=H, num_points=W, min_depth=MIN_D, max_depth=MAX_D,
                        angle_file=angle_file)
    g = torch.Generator().manual_seed(51)
    B = 3
    depth = MIN_D + (MAX_D - MIN_D) * torch.rand(B, 1, H, W, generator=g) ** 2
    mask = (torch.rand(B, 1, H, W, generator=g) < 0.8).float()
    depth = depth * mask
    out = dict(depth=npy(depth), mask=npy(mask), angle=npy(coord.angle))

    # demo_inversion.py:89-94
    t_depth = coord.convert(depth.clone(), "depth", "depth_norm")
    t_inv_depth = coord.convert(t_depth, "depth_norm", "inv_depth_norm")
    t_inv_depth = t_inv_depth * mask
    out.update(t_depth=npy(t_depth), t_inv_depth=npy(t_inv_depth))

    # demo_inversion.py:99-107
    with torch.no_grad():
        torch.manual_seed(0)
        z_samples = G.mapping_network(torch.randn(NUM_Z, 16))
        z_avg = z_samples.mean(dim=0, keepdim=True)
        z_std = (((z_samples - z_avg) ** 2).sum() / NUM_Z).sqrt()
    out.update(z_avg=npy(z_avg), z_std=npy(z_std))
    criterion = MultiScaleMaskedLoss(loss_fn=F.l1_loss, level=2)
    set_requires_grad(G, False)

    def lr_schedule(iteration):                                   # demo_inversion.py:140-146
        t = iteration / STEPS
        gamma = min(1.0, (1.0 - t) / RAMPDOWN)
        gamma = 0.5 - 0.5 * np.cos(gamma * np.pi)
        return gamma * min(1.0, t / RAMPUP)

    for latent_type in ("w", "w+"):
        n_styles = G.synthesis_network.num_styles
        z = z_avg.repeat_interleave(B, dim=0)
        if latent_type == "w+":
            z = torch.stack([z] * n_styles, dim=1)
            # spread the styles so the geodesic cross term has a non-degenerate gradient
            z = z + 0.05 * torch.randn(z.shape, generator=g)
        z = torch.nn.Parameter(z.clone()).requires_grad_()
        out[f"{latent_type}_z0"] = npy(z)
        phase = torch.zeros((B, 2, 1, 1))
        optim = torch.optim.Adam(params=[z], lr=LR)
        sched = torch.optim.lr_scheduler.LambdaLR(optim, lr_lambda=lr_schedule)
        for step in range(STEPS):
            # demo_inversion.py:149-191 (perturb_z off)
            w = torch.stack([z] * n_styles, dim=1) if latent_type == "w" else z
            imgs = G(w, angle=coord.angle + phase, input_w=True)
            g_inv_depth_orig = tanh_to_sigmoid(imgs["image_orig"])
            g_depth = coord.convert(g_inv_depth_orig, "inv_depth_norm", "depth_norm")
            loss = 0
            if latent_type == "w+":
                loss += 5e-3 * geocross_loss(w)
            loss += criterion(g_depth, t_depth, mask)
            loss += criterion(g_inv_depth_orig, t_inv_depth, mask)
            optim.zero_grad(set_to_none=True)
            loss.backward(gradient=torch.ones_like(loss))
            out[f"{latent_type}_loss{step}"] = npy(loss)
            out[f"{latent_type}_grad{step}"] = npy(z.grad)
            if step == 0:
                out[f"{latent_type}_inv_depth_orig0"] = npy(g_inv_depth_orig)
                out[f"{latent_type}_g_depth0"] = npy(g_depth)
                out[f"{latent_type}_raydrop_logit0"] = npy(imgs["raydrop_logit"])
            out[f"{latent_type}_lr{step}"] = np.array(optim.param_groups[0]["lr"])
            optim.step()
            sched.step()
            out[f"{latent_type}_z{step + 1}"] = npy(z)
    path = os.path.join(HERE, "inversion_loop.npz")
    np.savez_compressed(path, **out)
    print(f"inversion_loop.npz: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


if __name__ == "__main__":
    main()
