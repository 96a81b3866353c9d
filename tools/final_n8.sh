#!/bin/bash
# End-of-round STRONG scaling on one N-GPU box (run under `gpurun --gpus N`): bench.py through torchrun for the configs given.
# usage: tools/final_n8.sh <outdir> <N> "<configs>"
out=${1:-gpurun_out/final8}; N=${2:-8}; cfgs=${3:-"2 5 3"}; mkdir -p $out
for c in $cfgs; do
  case $c in 1) st="--steps 20 --warmup 5";; 2) st="--steps 5 --warmup 3";; 3) st="--steps 3 --warmup 3";; 4) st="--steps 2 --warmup 3";; 5) st="--steps 2 --warmup 3";; esac
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + c)) bench.py --gpus $N --config $c $st --no-cpu 2> $out/n${N}_c$c.err | tail -1 >> $out/scaling_n$N.jsonl
done
python - <<'P' $out/scaling_n$N.jsonl
import json,sys
for l in open(sys.argv[1]):
    try:
        d=json.loads(l); print(d["n_gpus"], d["config"]["workload"][:32], d["value"], d["ms_per_frame"], "kernel_ms", d["roofline"].get("kernel_ms_per_launch"))
    except Exception as e: print("bad line", e, l[:100])
P
