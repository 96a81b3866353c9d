#!/usr/bin/env python3
"""Summary JSON of one `ncu --set full` capture of wf_render for profiles/: the headline metrics from
the raw page plus the per-phase stall view of prof_phases.py.
usage: ncu_summary.py <report.ncu-rep> <libbhrt.so> <out.json> [--traffic profiles/traffic.json]"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rep, lib, out = sys.argv[1:4]
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
    kernel = vals[names.index("Kernel Name")] if "Kernel Name" in names else "wf_render"
    here = os.path.dirname(os.path.abspath(__file__))
    phases = subprocess.check_output([sys.executable, os.path.join(here, "prof_phases.py"), rep, lib, "wf_render"]).decode().splitlines()
    summary = {
        "what": "ncu --set full --import-source on --clock-control none, kernel %s, command: python bench.py --steps 1 --warmup 0 --no-cpu "
                "(BASELINE config 2, 1920x1080x16spp, one frame in ONE launch); per-phase view by tools/prof_phases.py" % kernel,
        "metrics": metrics,
        "phases_by_stall_samples": phases[:30],
    }
    json.dump(summary, open(out, "w"), indent=1)
    if "--traffic" in sys.argv:
        t = sys.argv[sys.argv.index("--traffic") + 1]
        rd, wr = metrics["dram__bytes_read.sum"], metrics["dram__bytes_write.sum"]
        total = to_bytes(rd["value"], rd["unit"]) + to_bytes(wr["value"], wr["unit"])
        json.dump({"wf_render_dram_bytes_per_launch": int(total),
                   "source": "%s (dram__bytes_read.sum + dram__bytes_write.sum of one launch = one 1920x1080x16spp frame)" % out}, open(t, "w"), indent=1)
    print(json.dumps({k: v["value"] for k, v in metrics.items()}, indent=1))
    print("\n".join(phases[:14]))


if __name__ == "__main__":
    main()
