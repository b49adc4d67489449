#!/bin/bash
# config #5 shapes: thousands of images, a seeded random sample of the all-pairs list is timed
mkdir -p gpurun_out
run() { # tag args
  timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-stages --no-e2e $2 > gpurun_out/sweep_$1.json 2> gpurun_out/sweep_$1.err
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/sweep_$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1: value %.0f pairs/s ms/step %.1f | knn avg %.3f ms share %.3f frac %.3f | %s | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["share_of_step"], r["frac"], d["config"].get("sample"), d["clocks"]))
except Exception as e: print("$1 parse fail", e); print(open("gpurun_out/sweep_$1.err").read()[-600:])
PYEOF
}
run sift_2000x4096 "--kind sift --images 2000 --kp 4096 --max-pairs 40000"
run orb_2000x8192 "--kind orb --images 2000 --kp 8192 --max-pairs 40000"
run sift_2000x8192 "--kind sift --images 2000 --kp 8192 --max-pairs 20000"
run orb_5000x4096 "--kind orb --images 5000 --kp 4096 --max-pairs 60000"
