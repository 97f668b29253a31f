#!/usr/bin/env python
"""Times the modulated-contraction kernels at the generator's conv1 shapes (B=64, bf16,
batch-shared Fourier block): per-sample tiles (impl 3) against the batch-fused tiles (impl 2),
forward and weight gradient, each checked against a bf16-operand / fp32-accumulate torch
reference on a subset of samples.

    python tools/modconv_bench.py [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import dusty_gan_v2_b200 as pkg  # noqa: E402
import dusty_gan_v2_b200.functional as DF  # noqa: E402
from dusty_gan_v2_b200 import _cabi as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pkg.set_precision("bf16")
    bf = torch.bfloat16
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, reps=5, inner=4):
        fn()
        torch.cuda.synchronize()
        if os.environ.get("DUSTY_KB_ONCE"):      # one launch per kernel: for `ncu --set full`
            return 0.0
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(inner):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / inner)
        return min(ts)

    rows = []
    for tag, O_, C1, C2, hh, ww in [("L4.conv1", 32, 64, 512, 64, 512), ("L3.conv1", 64, 128, 512, 32, 256),
                                    ("L2.conv1", 128, 256, 512, 16, 128), ("L1.conv1", 256, 512, 512, 8, 64)]:
        P = hh * ww
        Kt = C1 + C2
        g = torch.Generator(device=dev).manual_seed(3)
        wb = (torch.randn(B, O_, Kt, device=dev, generator=g) / Kt ** 0.5).to(bf)
        x1 = torch.randn(B, C1, hh, ww, device=dev, generator=g).to(bf)
        x2 = torch.randn(1, C2, hh, ww, device=dev, generator=g).to(bf)
        bias = torch.randn(O_, device=dev, generator=g)
        gy = torch.randn(B, O_, hh, ww, device=dev, generator=g).to(bf)
        gw = torch.empty(B, O_, Kt, device=dev)
        row = {"layer": tag, "B": B, "O": O_, "C1": C1, "C2": C2, "P": P,
               "gflop": 2.0 * B * O_ * Kt * P / 1e9,
               "hbm_mbytes_fwd": (x1.numel() + B * O_ * P + wb.numel() + x2.numel()) * 2 / 1e6}
        sub = [0, 1, B // 2, B - 1]
        xin = torch.cat([x1[sub].float(), x2.float().expand(len(sub), -1, -1, -1)], 1).reshape(len(sub), Kt, P)
        ref = torch.bmm(wb[sub].float(), xin).reshape(len(sub), O_, hh, ww) + bias.view(1, -1, 1, 1)
        ref = torch.where(ref > 0, ref, 0.2 * ref) * 1.41
        dw_ref = torch.bmm(gy[sub].float().reshape(len(sub), O_, P), xin.transpose(1, 2))
        for impl, name in [(3, "per_sample"), (2, "batch_fused")]:
            pkg.set_modconv_impl(impl)
            y = DF.modconv_bmm(wb, x1, x2, bias, 3, 0.2, 1.41)
            err = float((y[sub].float() - ref).abs().max() / ref.abs().max())
            row[f"fwd_{name}_us"] = timeit(lambda: DF.modconv_bmm(wb, x1, x2, bias, 3, 0.2, 1.41))
            row[f"fwd_{name}_relerr"] = err

            def dw():
                K.call("dusty_modconv_bwd_dw", K.ptr(gy), K.ptr(x1), K.ptr(x2), K.ptr(gw), B, O_, C1, C2, 1, P,
                       K.BF16, impl, 0, K.stream_of(gy))
            dw()
            errw = float((gw[sub] - dw_ref).abs().max() / dw_ref.abs().max())
            row[f"dw_{name}_us"] = timeit(dw)
            row[f"dw_{name}_relerr"] = errw
        pkg.set_modconv_impl(0)
        rows.append(row)
        print(json.dumps(row), flush=True)
    # conv2 (features only) forward and the input gradient of conv1 / conv2: pure streaming
    for tag, O_, C1, hh, ww in [("L4.conv2", 32, 32, 64, 512), ("L3.conv2", 64, 64, 32, 256),
                                ("L2.conv2", 128, 128, 16, 128)]:
        P = hh * ww
        g = torch.Generator(device=dev).manual_seed(5)
        wb = (torch.randn(B, O_, C1, device=dev, generator=g) / C1 ** 0.5).to(bf)
        x1 = torch.randn(B, C1, hh, ww, device=dev, generator=g).to(bf)
        bias = torch.randn(O_, device=dev, generator=g)
        gy = torch.randn(B, O_, hh, ww, device=dev, generator=g).to(bf)
        gx = torch.empty_like(x1)
        row = {"layer": tag, "B": B, "O": O_, "C1": C1, "P": P,
               "hbm_mbytes_fwd": (x1.numel() + B * O_ * P) * 2 / 1e6}
        row["fwd_us"] = timeit(lambda: DF.modconv_bmm(wb, x1, None, bias, 3, 0.2, 1.41))
        row["fwd_hbm_frac"] = row["hbm_mbytes_fwd"] / row["fwd_us"] / 6554.9 * 1e3
        row["dx_us"] = timeit(lambda: K.call("dusty_modconv_bwd_dx", K.ptr(wb), K.ptr(gy), K.ptr(gx), B, O_, C1, C1,
                                             P, K.BF16, K.BF16, 0, None, None, K.stream_of(gy)))
        sub = [0, B - 1]
        ref = torch.bmm(wb[sub].float().transpose(1, 2), gy[sub].float().reshape(2, O_, P)).reshape(2, C1, hh, ww)
        row["dx_relerr"] = float((gx[sub].float() - ref).abs().max() / ref.abs().max())
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.json:
        json.dump({"unit": "us", "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
