#!/usr/bin/env python3
"""Join an ncu report (--set full --import-source on) of one kernel with the line table of the
cubin it ran, and print where the executed warp instructions and the stall samples are, per
device function and per source-line bucket.

usage: prof_lines.py <report.ncu-rep> <libbhrt.so> [kernel-substring] [bucket]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_table(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    dis = ""
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        dis += subprocess.check_output(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)]).decode(errors="replace") + "\n"
    table = {}
    in_kernel = False
    fn = None
    cur = None
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            in_kernel = kernel in m.group(1)
            fn = "<kernel>"
            continue
        if not in_kernel:
            continue
        m = re.match(r"\s*(\$?\S+?):\s*$", l)
        if m and "$" in m.group(1) and "bhrt" in m.group(1):
            name = m.group(1).split("$")[-1]
            try:
                name = subprocess.check_output(["c++filt", name]).decode().strip().split("(")[0]
            except Exception:
                pass
            fn = name
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inner = (os.path.basename(m.group(1)), int(m.group(2)))
            chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            outer = (os.path.basename(chain[-1][0]), int(chain[-1][1])) if chain else inner
            cur = (inner, outer)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            table[int(m.group(1), 16)] = (fn, cur)
    return table


def main():
    rep, so = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else "wf_render"
    bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    table = line_table(so, kernel)
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]).decode(errors="replace")
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ie, it, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    ini = hdr.index("stall_no_inst")
    base = None
    by_fn = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    by_line = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    by_outer = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    tot = [0, 0, 0, 0, 0]
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        fn, cur = table.get(a - base, ("?", None))
        inner, outer = cur if cur else (("?", 0), ("?", 0))
        v = [int(r[ie] or 0), int(r[it] or 0), int(r[ism] or 0), int(r[ini] or 0), 1]
        for d, k in ((by_fn, fn), (by_line, (inner[0], inner[1] // bucket * bucket)), (by_outer, (fn, outer[0], outer[1] // bucket * bucket))):
            for i in range(5):
                d[k][i] += v[i]
        for i in range(5):
            tot[i] += v[i]
    print("total: warp inst %d, avg active %.1f, samples %d (no_inst %d), static instr %d" % (tot[0], tot[1] / max(tot[0], 1), tot[2], tot[3], tot[4]))

    def show(title, d, n):
        print("\n== %s  [%%inst | active | %%samples | %%no_inst | static]" % title)
        for k, v in sorted(d.items(), key=lambda x: -x[1][2])[:n]:
            print("%6.2f%% %5.1f %6.2f%% %6.2f%% %6d  %s" % (100.0 * v[0] / tot[0], v[1] / max(v[0], 1), 100.0 * v[2] / max(tot[2], 1), 100.0 * v[3] / max(tot[3], 1), v[4], k))
    show("by device function", by_fn, 30)
    show("by innermost source line bucket", by_line, 40)
    show("by outermost (kernel/function body) line bucket", by_outer, 50)


if __name__ == "__main__":
    main()
