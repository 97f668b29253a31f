#!/usr/bin/env python
"""Kernel-level baseline: the REFERENCE'S OWN CUDA kernels (oracle/_ref, built by
oracle/build_ref.py) timed beside ours -- a thin front end of `bench.reference_kernel_baseline`
(bench.py is the one measurement entry that executes anything under oracle/; the default
`python bench.py` run adds the same table to its JSON line as `ref_kernels`).

    python tools/ref_kernel_bench.py --json gpurun_out/ref_kernels.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "ref_kernels.json"))
    args = ap.parse_args()
    res = bench.reference_kernel_baseline()
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    json.dump(res, open(args.json, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
