#!/usr/bin/env python3
"""Headline metrics of one `ncu --set full` capture of ANY kernel as a small JSON for profiles/ (tools/ncu_summary.py is the
wf_render-specific one with the per-phase view).  Adds what showed the caustic gather's real problem: SM busy time, average
against maximum (sm__cycles_active) -- a kernel whose SMs are busy a quarter of its duration has a long pole, not a slow loop.
usage: ncu_metrics.py <report.ncu-rep> <out.json> "<what was captured>" """
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_active", "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out, what = sys.argv[1:4]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    metrics = {}
    for n, u, v in zip(names, units, vals):
        if n in WANT or (n.startswith(STALL) and n.endswith("_per_issue_active.ratio")):
            try:
                metrics[n] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                pass
    kernel = vals[names.index("Kernel Name")] if "Kernel Name" in names else "?"
    busy = None
    if "sm__cycles_active.avg" in metrics and "sm__cycles_elapsed.max" in metrics:
        busy = round(metrics["sm__cycles_active.avg"]["value"] / metrics["sm__cycles_elapsed.max"]["value"], 3)
    json.dump({"what": what, "kernel": kernel, "sm_busy_fraction_of_kernel_time": busy, "metrics": metrics}, open(out, "w"), indent=1)
    print(kernel, "busy", busy, {k: v["value"] for k, v in metrics.items() if not k.startswith(STALL)})


if __name__ == "__main__":
    main()
