#!/bin/bash
# End-of-round measurements on one B200 (run under gpurun): the five BASELINE configs through bench.py (value, e2e, roofline,
# both CPU figures), config 2 on the exact build and with the wide traversal, the reference arm at the full frame.
# usage: tools/final_n1.sh <outdir>
out=${1:-gpurun_out/final}; mkdir -p $out
: > $out/bench_configs_n1.jsonl
for c in 1 2 3 4 5; do
  case $c in 1) st="--steps 20 --warmup 5";; 2) st="--steps 5 --warmup 3";; 3) st="--steps 3 --warmup 3";; 4) st="--steps 2 --warmup 3";; 5) st="--steps 2 --warmup 3";; esac
  timeout 900 python bench.py --config $c $st 2> $out/bench_c$c.err | tail -1 >> $out/bench_configs_n1.jsonl
done
timeout 300 python bench.py --exact --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $out/bench_c2_exact.json
timeout 300 python bench.py --wide --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $out/bench_c2_wide.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 2>/dev/null | tail -1 > $out/bench_reference_c2.json
python - <<'P' $out
import json,sys
for l in open(sys.argv[1]+"/bench_configs_n1.jsonl"):
    try:
        d=json.loads(l); print(d["config"]["workload"][:40], d["value"], d["ms_per_frame"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print("bad line", e, l[:100])
P
