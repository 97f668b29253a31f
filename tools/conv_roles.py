"""Where do the tcgen05 convolution kernels wait?  Runs each own kernel of the chosen layers
once on the instrumented library (make -C dusty_gan_v2_b200/csrc prof) and prints, per launch,
the share of its lifetime each warp role spent waiting on the others.
    DUSTY_LIB=build/libdusty_b200_prof.so python tools/conv_roles.py [--layers 0,1,6]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("DUSTY_LIB", os.path.join(ROOT, "build", "libdusty_b200_prof.so"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import dusty_gan_v2_b200.functional as DF  # noqa: E402
from dusty_gan_v2_b200 import _cabi as K  # noqa: E402
from conv_once import layers  # noqa: E402


def read(reset=True):
    buf = (ctypes.c_double * 12)()
    K.call("dusty_conv_role_prof", ctypes.addressof(buf), 1 if reset else 0)
    return list(buf)


def report(tag, fn):
    torch.cuda.synchronize()
    read()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    p = read()
    ctas = max(p[7], 1)
    life = [max(p[4], 1), max(p[5], 1), max(p[6], 1)]
    print(f"{tag:34s} {s.elapsed_time(e) * 1e3:7.1f} us  ctas {int(ctas):4d}  cyc/cta {life[1] / ctas:9.0f} | "
          f"producer waits slot {100 * p[0] / life[0]:5.1f}% | mma waits operands {100 * p[1] / life[1]:5.1f}% "
          f"acc {100 * p[2] / life[1]:5.1f}% | epilogue waits acc {100 * p[3] / life[2]:5.1f}%"
          + (f" | issue loop {p[8] / p[9]:6.0f} cyc/block x {int(p[9] / ctas)} blocks, fence+elect {p[10] / p[9]:5.0f}, "
             f"commit {p[11] / p[9]:5.0f}" if p[9] else ""), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--layers", default=None)
    args = ap.parse_args()
    dev, cl = "cuda", torch.channels_last
    sel = None if args.layers is None else {int(v) for v in args.layers.split(",")}
    for idx, (name, C, O, H, W, k, s) in enumerate(layers(args.batch)):
        if sel is not None and idx not in sel:
            continue
        x = torch.randn(args.batch, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=cl)
        w = (torch.randn(O, C, k, k, device=dev) / (C * k * k) ** 0.5).to(torch.bfloat16)
        st = (s, s)
        y = DF.conv2d_fprop_tc(x, w, st)
        gy = torch.randn_like(y).contiguous(memory_format=cl)
        w_tco = DF.filter_tco(w)
        for _ in range(2):
            DF.conv2d_dgrad_tc(gy, w, st, (H, W), w_tco)
            DF.conv2d_wgrad_tc(gy, x, st, w.shape, torch.float32)
        report(f"{name} {C}->{O} fprop", lambda: DF.conv2d_fprop_tc(x, w, st))
        report(f"{name} {C}->{O} dgrad", lambda: DF.conv2d_dgrad_tc(gy, w, st, (H, W), w_tco))
        report(f"{name} {C}->{O} wgrad", lambda: DF.conv2d_wgrad_tc(gy, x, st, w.shape, torch.float32))


if __name__ == "__main__":
    main()
