"""Golden vectors for BASELINE config 5 (GAN inversion losses) from the UNMODIFIED reference
`gans/inversion.py` on CPU.  Usage (only where /root/reference exists):

    python tests/golden/make_golden_inversion.py      ->  tests/golden/inversion.npz
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

ref_import.install()

from gans import inversion as ref_inv  # noqa: E402

torch.set_num_threads(4)
npy = lambda t: t.detach().cpu().numpy().copy()  # noqa: E731  (copy: p is updated in place later)


def main():
    g = torch.Generator().manual_seed(123)
    B, H, W = 3, 16, 64
    gen = (torch.rand(B, 1, H, W, generator=g) * 0.8 + 0.1).requires_grad_()
    ref = torch.rand(B, 1, H, W, generator=g) * 0.8 + 0.1
    mask = (torch.rand(B, 1, H, W, generator=g) < 0.7).float()
    mask[0, :, 4:12, 8:40] = 0.0                      # a hole that survives two pooling levels
    out = {"gen": npy(gen), "ref": npy(ref), "mask": npy(mask)}
    for tag, kw in (("l1_rel_full", dict(loss_fn=F.l1_loss, level=None, relative=True)),
                    ("l1_rel_l2", dict(loss_fn=F.l1_loss, level=2, relative=True)),
                    ("l2_abs_l3", dict(loss_fn=F.mse_loss, level=3, relative=False))):
        crit = ref_inv.MultiScaleMaskedLoss(**kw)
        loss = crit(gen, ref, mask)
        (gg,) = torch.autograd.grad(loss.sum(), gen)
        out[f"{tag}_loss"] = npy(loss)
        out[f"{tag}_grad"] = npy(gg)
    lat = torch.randn(2, 10, 16, generator=g).requires_grad_()
    gl = ref_inv.geocross_loss(lat)
    (glg,) = torch.autograd.grad(gl.sum(), lat)
    out.update(lat=npy(lat), geocross=npy(gl), geocross_grad=npy(glg))
    # SphericalOptimizer: two steps with fixed gradients
    p = torch.nn.Parameter(torch.randn(2, 10, 16, generator=g))
    out["sph_p0"] = npy(p)
    opt = ref_inv.SphericalOptimizer([p], lr=0.1, betas=(0.9, 0.999))
    for i in range(2):
        p.grad = torch.randn(2, 10, 16, generator=g)
        out[f"sph_g{i}"] = npy(p.grad)
        opt.step()
        out[f"sph_p{i + 1}"] = npy(p)
    path = os.path.join(HERE, "inversion.npz")
    np.savez_compressed(path, **out)
    print(f"inversion.npz: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


if __name__ == "__main__":
    main()
