#!/usr/bin/env python
"""Runs every dusty_b200 kernel family once or twice at the training step's real shapes
(B=64, 64x512, bf16).  Meant to be wrapped in `ncu --set full -k regex:dusty` so that one
capture holds the DRAM traffic / pipe utilisation of each kernel; also prints CUDA-event
timings (isolated, L2 flushed) as JSON when run bare.

    python tools/kernel_bench.py [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    dom, kernels, family = bench.kernel_rooflines(dev, peaks)
    out = {"dominant": dom, "kernels": kernels, "conv_family": family}
    print(json.dumps(out, indent=1))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
