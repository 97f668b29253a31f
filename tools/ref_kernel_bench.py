#!/usr/bin/env python
"""Kernel-level baseline: the REFERENCE'S OWN CUDA kernels (built from /root/reference into
oracle/_ref by oracle/build_ref.py, sm_100 recompiles of fused_bias_act_kernel.cu /
upfirdn2d_kernel.cu) timed beside ours on the same tensors, at the training step's shapes.
CUDA events, L2 flushed between timed groups.

    python tools/ref_kernel_bench.py --json gpurun_out/ref_kernels.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import build_ref  # noqa: E402  (measurement tool: the reference arm, not the product)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "ref_kernels.json"))
    args = ap.parse_args()
    import dusty_gan_v2_b200.functional as DF
    from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d
    fused = build_ref.load_built("dusty_ref_fused")
    ufd = build_ref.load_built("dusty_ref_upfirdn2d")
    if fused is None or ufd is None:
        print(json.dumps({"unavailable": "oracle/_ref not built"}))
        return
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, reps=6, inner=4):
        """Median device time per call in us.  The `inner` calls are captured into a CUDA graph
        and replayed, so that the host cost of either side's Python / pybind wrapper (20-40 us,
        more than some of these kernels take) stays out of the number."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    for _ in range(inner):
                        fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = gr
        except Exception as e:          # fall back to eager launches
            print(f"graph capture failed ({type(e).__name__}); eager timing", file=sys.stderr)
            torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if graph is not None:
                graph.replay()
            else:
                for _ in range(inner):
                    fn()
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) / inner)
        best.sort()
        return best[len(best) // 2] * 1e3          # us

    rows = []

    def row(name, ref_fn, our_fn, bytes_):
        r, o = timeit(ref_fn), timeit(our_fn)
        rows.append({"kernel": name, "reference_us": round(r, 1), "ours_us": round(o, 1),
                     "speedup": round(r / o, 2), "algorithmic_bytes": bytes_,
                     "reference_gbs": round(bytes_ / r / 1e3, 1), "ours_gbs": round(bytes_ / o / 1e3, 1)})

    B, H, W = 64, 64, 512
    for dt_ref, dt_our, tag in ((torch.float32, torch.float32, "f32"), (torch.float16, torch.bfloat16, "f16|bf16")):
        x = torch.randn(B, 32, H, W, device=dev)
        b = torch.randn(32, device=dev)
        xr, br, xo, bo = x.to(dt_ref), b.to(dt_ref), x.to(dt_our), b.to(dt_our)
        er, eo = xr.new_empty(0), xo.new_empty(0)
        es = x.element_size() if dt_ref == torch.float32 else 2
        row(f"bias_act_fwd[64,32,64,512]{tag}", lambda: fused.fused_bias_act(xr, br, er, 3, 0, 0.2, 1.41),
            lambda: DF.fused_bias_act(xo, bo, eo, 3, 0, 0.2, 1.41), 2 * x.numel() * es)
        yr = fused.fused_bias_act(xr, br, er, 3, 0, 0.2, 1.41)
        yo = DF.fused_bias_act(xo, bo, eo, 3, 0, 0.2, 1.41)
        # the reference's backward = the gradient kernel + a separate ATen reduction for db
        # (fused_act.py:28-40); ours produces dx and db in one pass
        row(f"bias_act_bwd+db[64,32,64,512]{tag}",
            lambda: fused.fused_bias_act(xr, er, yr, 3, 1, 0.2, 1.41).sum((0, 2, 3)),
            lambda: DF._BiasActBackward.apply(xo, yo, True, 0.2, 1.41), 3 * x.numel() * es)
        del x, xr, xo, yr, yo
    # AdaptiveAugment's four separable passes (adaptive_augment.py:497-545), 1-channel fp32
    k = torch.tensor([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057,
                      -0.048311742585633, 0.4910559419267466, 0.787641141030194, 0.3379294217276218,
                      -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
                      0.0017677118642428036, -0.007800708325034148], device=dev)
    for name, hw, ks, up, down, pad in (("ada_up_x", (76, 524), (1, 12), (2, 1), (1, 1), (6, 5, 0, 0)),
                                        ("ada_up_y", (76, 1048), (12, 1), (1, 2), (1, 1), (0, 0, 6, 5)),
                                        ("ada_down_x", (140, 1036), (1, 12), (1, 1), (2, 1), (-1, -1, 0, 0)),
                                        ("ada_down_y", (140, 512), (12, 1), (1, 1), (1, 2), (0, 0, -1, -1))):
        x = torch.randn(B, 1, hw[0], hw[1], device=dev)
        kk = k.reshape(ks).contiguous()
        x4 = x.reshape(B, hw[0], hw[1], 1)
        a = (up[0], up[1], down[0], down[1], pad[0], pad[1], pad[2], pad[3])
        out = ufd.upfirdn2d(x4, kk, *a)
        row(f"upfirdn2d_{name}[64,1,{hw[0]},{hw[1]}]f32", lambda: ufd.upfirdn2d(x4, kk, *a),
            lambda: upfirdn2d(x, kk, up=up, down=down, pad=pad), (x.numel() + out.numel()) * 4)
    res = {"device": torch.cuda.get_device_name(0), "note": "reference kernels = sm_100 builds of the "
           "reference's own .cu files (oracle/_ref); median of 6 graph replays of 4 launches (device time, no host "
           "wrapper cost), L2 flushed before each replay", "rows": rows}
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    json.dump(res, open(args.json, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
