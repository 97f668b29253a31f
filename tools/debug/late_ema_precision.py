"""ModConv2d forward / gradients in bf16 with the EMA normaliser folded into the weights vs applied
in the contraction epilogue, both against an fp64 evaluation of the reference algebra (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import dusty_gan_v2_b200 as pkg
import dusty_gan_v2_b200.functional as DF
from dusty_gan_v2_b200.gans.models import ops

dev = torch.device("cuda", 0)
pkg.set_precision("bf16")
torch.manual_seed(0)
B, C1, C2, O, H, W = 8, 64, 512, 64, 16, 128
for ema_val in (1.0, 0.37, 7.3):
    m = ops.ModConv2d(in_ch=C1 + C2, out_ch=O, mod_ch=32, ksize=1, stride=1, padding=0, bias=False, ema=True).to(dev)
    m.eval()
    m.ema_var.fill_(ema_val)
    act = ops.FusedLeakyReLU(O).to(dev)
    act.bias.data.normal_(0, 0.3)
    x = torch.randn(B, C1, H, W, device=dev)
    pe = torch.randn(1, C2, H, W, device=dev)
    style = torch.randn(B, 32, device=dev)
    gy = torch.randn(B, O, H, W, device=dev)
    # fp64 reference
    xd = x.bfloat16().double().requires_grad_()
    ped = pe.bfloat16().double()
    sd = style.double().requires_grad_()
    wd = m.weight.detach().double().requires_grad_()
    lw, lb = m.mod.module.weight.detach().double(), m.mod.module.bias.detach().double()
    s = torch.nn.functional.linear(sd, lw * m.mod.scale if hasattr(m.mod, "scale") else lw, lb)
    w = wd.reshape(O, C1 + C2) * m.scale
    w = w / w.abs().amax(); s2 = s / s.abs().amax(dim=1, keepdim=True)
    wb = w.unsqueeze(0) * (s2 + 1).unsqueeze(1)
    wb = wb * torch.rsqrt(wb.square().sum(dim=2, keepdim=True) + 1e-8) / (ema_val ** 0.5 + 1e-8)
    xin = torch.cat([xd, ped.expand(B, -1, -1, -1)], 1).reshape(B, C1 + C2, H * W)
    pre = torch.bmm(wb, xin).reshape(B, O, H, W) + act.bias.detach().double().view(1, -1, 1, 1)
    ref = torch.where(pre > 0, pre, 0.2 * pre) * 2 ** 0.5
    for late in (False, True):
        DF.set_late_ema(late)
        xg = x.bfloat16().requires_grad_()
        sg = style.clone().requires_grad_()
        m.weight.grad = None
        y = m(xg, sg, pe=pe.bfloat16(), fused_act=act)
        # gradients through the device's own gates
        gate = torch.where(y.detach().double() > 0, 1.0, 0.2) * 2 ** 0.5
        gx_ref, gs_ref, gw_ref = torch.autograd.grad(pre, [xd, sd, wd], gy.double() * gate, retain_graph=True)
        gx, gs, gw = torch.autograd.grad(y, [xg, sg, m.weight], gy.bfloat16())
        rel = lambda a, b: float((a.double() - b).norm() / b.norm())
        print(f"ema_var {ema_val:5.2f} late={late!s:5}  y {rel(y, ref):.5f}  dx {rel(gx, gx_ref):.5f}  "
              f"dstyle {rel(gs, gs_ref):.5f}  dweight {rel(gw, gw_ref.reshape(gw.shape)):.5f}")
