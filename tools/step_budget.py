#!/usr/bin/env python
"""Groups the kernel-time breakdown of tools/profile_step.py (JSON) into the cost centres of the
training step -- where the next millisecond is.  No GPU needed (reads the committed profile).

    python tools/step_budget.py gpurun_out/step_profile.json > profiles/rNN_step_budget.txt
"""
import json
import re
import sys
from collections import OrderedDict

GROUPS = OrderedDict([
    ("D convolutions (own tcgen05: fprop / dgrad / wgrad, halo)", r"conv_(fwd|halo|wgrad|pair)_tc_kernel|conv_simt"),
    ("D conv weight preparation / filter layout", r"weight_prep|filter_rsco|filter_tco"),
    ("NHWC stencils (blur, blur+pad, blur+decimate, pad, fork)", r"blur4_cl|blur4_down2|pad2d_cl|residual_fork"),
    ("bias_act / residual tail (NHWC + NCHW)", r"bias_act"),
    ("D stem + minibatch-stddev + R1 statistics", r"stem_|minibatch|mbstd|sumsq_rows"),
    ("modulated contractions (tcgen05) + heads", r"modconv_|small_o_kernel|heads_dw|gemm_n"),
    ("modprep (per-sample weights, styles)", r"modprep_"),
    ("G resampling / Fourier / raydrop / shift", r"up2_|blur4_fwd|blur4_adj|fourier|raydrop|circ|angle_down|ema_lerp|point_project"),
    ("ADA (device-side op: sample, row pass, column pass)", r"ada_rows|ada_cols|ada_sample|fir1d|fir2d|affine_warp|pad2d_(fwd|adj)_kernel"),
    ("linears (own GEMMs: epilogue; cuBLAS: mapping / styles)", r"cutlass|sgemm|cublas|gemv|gemm_tc_kernel|gemm_simt|nvjet"),
    ("optimizer (fused Adam, EMA foreach)", r"Adam|multi_tensor|FusedOptimizer|lerp|multi_adam|multi_copy"),
    ("ATen glue: gradient accumulation adds", r"CUDAFunctor_add"),
    ("ATen glue: copies / casts", r"copy_kernel|direct_copy|bfloat16_copy|CatArray"),
    ("memset", r"Memset"),
])


def main():
    d = json.load(open(sys.argv[1]))
    steps, total = d["steps"], d["kernel_ms"]
    acc = OrderedDict((k, [0.0, 0]) for k in GROUPS)
    other = [0.0, 0, []]
    for t in d["top"]:
        for name, pat in GROUPS.items():
            if re.search(pat, t["name"]):
                acc[name][0] += t["ms"]
                acc[name][1] += t["calls"]
                break
        else:
            other[0] += t["ms"]
            other[1] += t["calls"]
            other[2].append((t["ms"], t["name"][:70]))
    listed = sum(t["ms"] for t in d["top"])
    print(f"# {steps} iterations (one with R1): {total / steps:.2f} ms of kernel time per iteration, "
          f"{d['event_ms'] / steps:.2f} ms of device time per iteration")
    print(f"{'cost centre':66s} {'ms/iter':>8s} {'share':>6s} {'launches/iter':>14s}")
    for name, (ms, calls) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"{name:66s} {ms / steps:8.2f} {100 * ms / total:5.1f}% {calls / steps:14.1f}")
    print(f"{'other ATen / small kernels':66s} {(other[0] + total - listed) / steps:8.2f} "
          f"{100 * (other[0] + total - listed) / total:5.1f}% {other[1] / steps:14.1f}")
    for ms, name in sorted(other[2], reverse=True)[:6]:
        print(f"    {ms / steps:6.3f} ms/iter  {name}")


if __name__ == "__main__":
    main()
