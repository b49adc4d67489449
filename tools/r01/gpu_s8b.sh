#!/bin/bash
mkdir -p gpurun_out
run() { # tag extra-args
  timeout 600 python bench.py --kind superpoint --steps 3 --warmup 2 --no-cpu-baseline --no-stages --no-e2e $2 > gpurun_out/s8_$1.json 2> gpurun_out/s8_$1.err
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/s8_$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1: value %.0f pairs/s ms/step %.1f | knn avg %.3f ms x %d share %.3f | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["launches"], r["share_of_step"], d["clocks"]))
except Exception as e: print("$1 parse fail", e); print(open("gpurun_out/s8_$1.err").read()[-600:])
PYEOF
}
run full64 "--images 64"
run nofilter64 "--images 64 --dev-no-filter"
run knnonly64 "--images 64 --dev-ratio 0.001"
run full100 "--images 100"
run knnonly100 "--images 100 --dev-ratio 0.001"
run full100_bp128 "--images 100 --batch-pairs 128"
run full100_bp512 "--images 100 --batch-pairs 512"
