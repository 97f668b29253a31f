#!/usr/bin/env python
"""Reads an `ncu --set full` report of tools/kernel_bench.py (DUSTY_KB_ONCE=1: one launch per
bench entry, in bench order) and writes profiles/ncu_traffic.json: for every entry of
bench.kernel_rooflines the DRAM bytes of its launch (dram__bytes_read.sum +
dram__bytes_write.sum), duration, DRAM %, tensor-pipe % -- bench.py copies `dram_bytes` into
the `traffic` field of the roofline block.

    ncu -i gpurun_out/kb.ncu-rep --page raw --csv > /tmp/kb.csv
    python tools/ncu_traffic.py /tmp/kb.csv gpurun_out/kernels.json profiles/ncu_traffic.json

Matching is by order: kernel_bench prints its entries in launch order and the report is
filtered (-k regex) to the kernels those entries launch; entries whose launch count differs
(e.g. an op that runs two kernels) are matched through the name hints below.
"""
import csv
import json
import sys

HINTS = [  # bench entry prefix -> substring of the kernel name that carries the traffic
    ("bias_act_fwd_nhwc", "bias_act_cl_kernel"), ("bias_act_bwd_nhwc", "bias_act_bwd_cl_kernel"),
    ("bias_act_fwd", "bias_act_vec_kernel"), ("bias_act_bwd", "bias_act_bwd_kernel"),
    ("resample_up2_adjoint", "up2_adj_kernel"), ("resample_up2", "up2_fwd_kernel"),
    ("resample_blur_adjoint", "blur4_adj_kernel"), ("resample_blur_nhwc", "blur4_cl_kernel"),
    ("resample_blur", "blur4_fwd_kernel"), ("blur+pad_fused_adjoint", "blur4_cl_kernel"),
    ("blur+pad_fused", "blur4_cl_kernel"), ("pad_ring1_adjoint_nhwc", "pad2d_cl_adj_kernel"),
    ("pad_ring1_nhwc", "pad2d_cl_fwd_kernel"), ("pad_ring1_adjoint", "pad2d_adj_kernel"),
    ("pad_ring1", "pad2d_fwd_kernel"), ("stem_fwd", "stem_fwd_kernel"), ("stem_bwd", "stem_bwd_kernel"),
    ("residual_fork_bwd", "residual_fork_bwd_cl_kernel"), ("residual_tail_fwd", "bias_act_add_cl_kernel"), ("heads_fwd", "small_o_kernel"),
    ("heads_dw", "heads_dw_bf16_kernel"), ("fourier", "fourier_kernel"),
    ("sumsq", "sumsq_rows_kernel"), ("gumbel_raydrop", "raydrop"), ("point_project", "point_project"),
    ("upfirdn2d_ada", "fir1d"), ("modconv_fwd", "modconv_fwd"), ("modconv_dw", "modconv_dw"),
    ("modconv_dx", "modconv_fwd_tc_kernel"),
    ("conv_fprop", ("conv_halo_tc_kernel", "conv_fwd_tc_kernel", "conv_pair_tc_kernel")),
    ("conv_dgrad", ("conv_halo_tc_kernel", "conv_fwd_tc_kernel", "conv_pair_tc_kernel")),
    ("conv_wgrad", "conv_wgrad_tc_kernel"),
]


def main():
    raw, kernels_json, out = sys.argv[1:4]
    rows = list(csv.reader(open(raw)))
    h = rows[0]
    col = {k: h.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                   "dram__bytes_write.sum",
                                   "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                   "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
    units = rows[1]

    def to_bytes(v, unit):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    launches = []
    for r in rows[2:]:
        launches.append({
            "name": r[col["Kernel Name"]],
            "us": float(r[col["gpu__time_duration.sum"]]) * (1e-3 if units[col["gpu__time_duration.sum"]] == "ns" else 1.0),
            "dram_bytes": to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) +
                          to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]),
            "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
            "tensor_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
        })
    entries = [k["kernel"] for k in json.load(open(kernels_json))["kernels"]]
    res, pos = {}, 0
    for e in entries:
        hint = next((sub for pre, sub in HINTS if e.startswith(pre)), None)
        if hint is None:
            continue
        hints = hint if isinstance(hint, tuple) else (hint,)
        j = next((i for i in range(pos, len(launches)) if any(h_ in launches[i]["name"] for h_ in hints)), None)
        if j is None:
            continue
        pos = j + 1
        l = launches[j]
        res[e] = {"dram_bytes": int(l["dram_bytes"]), "ncu_us": round(l["us"], 2), "dram_pct": round(l["dram_pct"], 1),
                  "tensor_pct": round(l["tensor_pct"], 1), "ncu_kernel": l["name"][:90]}
    json.dump({"source": raw, "note": "per launch, ncu --set full --clock-control none, DUSTY_KB_ONCE=1",
               "kernels": res}, open(out, "w"), indent=1)
    print(f"matched {len(res)} of {len(entries)} entries")


if __name__ == "__main__":
    main()
