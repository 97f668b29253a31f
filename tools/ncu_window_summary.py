#!/usr/bin/env python
"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py's timed
region by kernel name: total time, share, calls; and the share of `dusty::` kernels.

    python tools/ncu_window_summary.py gpurun_out/launches.csv > profiles/rNN_ncu_launches_window_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = defaultdict(lambda: [0.0, 0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*$", "", r["Kernel Name"])[:110]
        agg[name][0] += us
        agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    n = sum(v[1] for v in agg.values())
    own = sum(v[0] for k, v in agg.items() if "dusty" in k)
    print(f"# ncu launch list of the timed region: {n} launches, {total / 1e3:.2f} ms of kernel time "
          f"(cold-cache, serialised); dusty:: kernels {100 * own / total:.1f} % of it")
    print("#   us_total  share  calls  kernel")
    for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
        print(f"{us:10.1f} {100 * us / total:5.1f}% {c:6d}  {k}")


if __name__ == "__main__":
    main()
