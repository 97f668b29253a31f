"""Where does the bf16 discriminator lose precision?  Full-size D, same weights: per-layer
relative L2 error of the bf16 GPU path against the fp32 GPU path (forward activations), and of
logits / input gradient against the fp32 CPU oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dusty_gan_v2_b200 as pkg  # noqa: E402
from dusty_gan_v2_b200.gans.models.builder import build_discriminator  # noqa: E402
from dusty_gan_v2_b200.presets import preset  # noqa: E402


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def run(D, x, precision):
    pkg.set_precision(precision)
    acts = {}
    hooks = []

    def fwd_hook(name):
        def fn(mod, inp, out):
            if isinstance(out, torch.Tensor):
                acts[name] = out.detach().float().clone()
                if out.requires_grad:
                    out.register_hook(lambda g: acts.__setitem__("grad@" + name, g.detach().float().clone()))
        return fn
    for name, m in D.named_modules():
        if name.count(".") <= 1 and name:
            hooks.append(m.register_forward_hook(fwd_hook(name)))
    xg = x.clone().requires_grad_()
    y = D(xg)
    for p in D.parameters():
        p.requires_grad_(True)
        p.grad = None
    torch.nn.functional.softplus(-y).mean().backward()
    gx = xg.grad
    for n, p in D.named_parameters():
        acts["wgrad@" + n] = p.grad.detach().float().clone()
    for h in hooks:
        h.remove()
    return y.detach().float(), gx.detach().float(), acts


def main():
    torch.manual_seed(0)
    D = build_discriminator(preset("dusty_v2").model.discriminator)
    with torch.no_grad():
        for n, p in D.named_parameters():
            if "bias" in n:
                p.normal_(0, 0.2)
    D = D.cuda()
    x = torch.tanh(torch.randn(8, 1, 64, 512, device="cuda"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y32, g32, a32 = run(D, x, "fp32")
    y16, g16, a16 = run(D, x, "bf16")
    print("logits fp32", y32.flatten().tolist())
    print("logits bf16", y16.flatten().tolist())
    print("logit rel", rel(y16, y32), " grad_x rel", rel(g16, g32))
    for k in a32:
        if k in a16 and a16[k].shape == a32[k].shape:
            print(f"{k:28s} rel_l2 {rel(a16[k], a32[k]):.5f}   |ref| rms {float(a32[k].pow(2).mean().sqrt()):.4f}")
        else:
            print(f"{k:28s} (not comparable: {tuple(a32[k].shape)} vs {tuple(a16.get(k, torch.zeros(0)).shape)})")


if __name__ == "__main__":
    main()
