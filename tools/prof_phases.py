#!/usr/bin/env python3
"""Phase-level view of an ncu report of wf_render: share of executed warp instructions, stall samples
and the dominant stall reasons per phase / device function.
usage: prof_phases.py <report.ncu-rep> <libbhrt.so> [kernel-regex]"""
import collections
import csv
import io
import subprocess
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import prof_lines as pl


def _line_of(path, marker):
    for i, l in enumerate(open(path), 1):
        if marker in l:
            return i
    raise SystemExit("marker %r not found in %s" % (marker, path))


def core_ranges():
    """Line ranges of core.cuh's sections, found by their own markers (they move with every edit)."""
    core = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bhraytracer_b200", "csrc", "device", "core.cuh")
    return {
        "trav": (_line_of(core, "BHRT_HD bool intersect_box6("), _line_of(core, "// A ray query: the ONE place the estimator meets the geometry")),
        "graph": (_line_of(core, "BHRT_HD bool scene_trace_graph("), _line_of(core, "BHRT_HD Ray node_local_ray(")),
        "job": (_line_of(core, "BHRT_HD Ray node_local_ray("), _line_of(core, "// Winner-only part of the closest hit")),
        "est": (_line_of(core, "struct alignas(128) Frame {"), _line_of(core, "// Caustic photon emission: BuildCausticPhotonMap")),
    }


_R = None


def phase(fn, cur):
    global _R
    if _R is None:
        _R = core_ranges()
    if cur is None:
        return "other"
    (f, l), _ = cur
    if fn != "<kernel>":
        return "fn:" + fn.split("::")[-1]
    if f == "core.cuh":
        if _R["trav"][0] <= l < _R["trav"][1]:
            return "P2:traversal (box/tri/step)"
        if _R["job"][0] <= l < _R["job"][1]:
            return "P2:job control"
        if _R["graph"][0] <= l < _R["graph"][1]:
            return "P1:scene graph loop"
        if _R["est"][0] <= l < _R["est"][1]:
            return "P1:estimator step"
    if f == "wavefront.cuh":
        return "wf:%d" % (l // 10 * 10)
    return "inl:%s:%d" % (f, l // 25 * 25)


def main():
    rep, so = sys.argv[1], sys.argv[2]
    kern = sys.argv[3] if len(sys.argv) > 3 else "wf_render"
    table = pl.line_table(so, kern)
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]).decode(errors="replace")
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, smp, ie, it = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][ia], 16)
    tot = sum(int(r[smp] or 0) for r in data)
    agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in st}
    print("stalls:", {k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
    ph = collections.defaultdict(collections.Counter)
    for r in data:
        fn, cur = table.get(int(r[ia], 16) - base, ("?", None))
        k = phase(fn, cur)
        ph[k]["inst"] += int(r[ie] or 0)
        ph[k]["tinst"] += int(r[it] or 0)
        ph[k]["smp"] += int(r[smp] or 0)
        for i in st:
            ph[k][hdr[i]] += int(r[i] or 0)
    ti = sum(v["inst"] for v in ph.values())
    print("total warp inst %d, avg active %.1f" % (ti, sum(v["tinst"] for v in ph.values()) / ti))
    for k, v in sorted(ph.items(), key=lambda x: -x[1]["smp"])[:30]:
        top = sorted(((v[hdr[i]], hdr[i][6:]) for i in st), reverse=True)[:3]
        print("%-34s inst %5.1f%% act %4.1f  samples %5.1f%%  %s" % (k, 100 * v["inst"] / ti, v["tinst"] / max(v["inst"], 1), 100 * v["smp"] / tot,
                                                                     [(n, round(100 * c / max(v["smp"], 1))) for c, n in top]))


if __name__ == "__main__":
    main()
