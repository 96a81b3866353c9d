run() { echo "$1"; env $1 timeout 60 python bench.py --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E "metric" | cut -c1-45 | tail -1; }
run "BHRT_X=0"
run "BHRT_MESH_BUDGET=3"
run "BHRT_MESH_BUDGET=6"
run "BHRT_MESH_BATCH=512"
run "BHRT_MESH_BATCH=2048"
run "BHRT_TRAV=2,12,50"
run "BHRT_TRAV=3,16,40"
run "BHRT_TRAV=4,12,60"
