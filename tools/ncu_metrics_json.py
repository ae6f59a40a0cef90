#!/usr/bin/env python
"""Reads an `ncu --set full` report of the SHIPPED kernels and merges, per C-ABI call, the measured DRAM bytes per launch
and the pipe / issue utilisations into profiles/r2_kernel_metrics.json -- the file bench.py takes `roofline.traffic` and
`roofline.issue` from (instead of constants in the source).
Usage: python tools/ncu_metrics_json.py gpurun_out/x.ncu-rep WORKLOAD profiles/x_summary.csv"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_kernel_metrics.json")
KERNEL_TO_CALL = {"blend_fwd_kernel": "dimo_raster_blend_fwd", "blend_bwd_kernel": "dimo_raster_blend_bwd",
                  "preprocess_fwd_kernel": "dimo_raster_preprocess", "preprocess_bwd_kernel": "dimo_raster_preprocess_bwd",
                  "preprocess_bwd_shared_kernel": "dimo_raster_preprocess_bwd",
                  "depth_sort_kernel": "depth_sort_kernel", "tile_scatter_kernel": "tile_scatter_kernel",
                  "ssim_fwd_kernel": "dimo_ssim_fwd", "ssim_bwd_kernel": "dimo_ssim_bwd", "lbs_bwd_kernel": "dimo_lbs_bwd",
                  "lbs_fwd_kernel": "dimo_lbs_fwd", "tn_gemm_kernel": "tn_gemm_kernel", "tn_wgrad_kernel": "tn_wgrad_kernel",
                  "tn_heads_bwd_kernel": "tn_heads_bwd_kernel", "tn_heads_fwd_kernel": "tn_heads_fwd_kernel",
                  "adam_kernel": "dimo_adam_step", "knn_kernel": "dimo_knn"}
COLS = {"dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "pipe_fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "pipe_alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "pipe_xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "issue_active": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "tensor_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "duration": "gpu__time_duration.sum"}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0}


def main(rep, workload, source):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in rows[2:]:
        kname = r[ix["Kernel Name"]].split("<")[0].split("(")[0].split("::")[-1].strip().split(" ")[-1]
        call = KERNEL_TO_CALL.get(kname)
        if call is None:
            continue
        if call == "dimo_ssim_bwd" and "<0>" in r[ix["Kernel Name"]].replace(" ", ""):
            continue                                  # the mask-MSE launch (no filter): keep the image launch only
        rec = acc.setdefault(call, {"n": 0})
        rec["n"] += 1
        for k, col in COLS.items():
            if col in ix and r[ix[col]] not in ("", "n/a"):
                v = float(r[ix[col]].replace(",", "")) * UNIT_SCALE.get(units[ix[col]], 1.0)
                rec[k] = rec.get(k, 0.0) + v
    data = {}
    if os.path.exists(OUT):
        data = json.load(open(OUT))
    wl = data.setdefault(workload, {})
    for call, rec in acc.items():
        n = rec.pop("n")
        m = {k: v / n for k, v in rec.items()}
        wl[call] = {"dram_bytes": m.get("dram_read", 0.0) + m.get("dram_write", 0.0), "dram_read": m.get("dram_read"),
                    "dram_write": m.get("dram_write"), "pipe_fma": m.get("pipe_fma"), "pipe_alu": m.get("pipe_alu"),
                    "pipe_xu": m.get("pipe_xu"), "issue_active": m.get("issue_active"),
                    "tensor_active": m.get("tensor_active"), "dram_pct": m.get("dram_pct"),
                    "duration_ms_under_ncu": m.get("duration"), "launches_averaged": n, "source": source}
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(wl, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
