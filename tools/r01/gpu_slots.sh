#!/bin/bash
mkdir -p gpurun_out
run() { # tag slots kind extra
  PM_SLOTS=$2 timeout 600 python bench.py --kind $3 --steps 3 --warmup 2 --no-cpu-baseline --no-stages --no-e2e $4 > gpurun_out/sl_$1.json 2> gpurun_out/sl_$1.err
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/sl_$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1: value %.0f pairs/s ms/step %.1f | knn avg %.3f ms x %d share %.3f | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["launches"], r["share_of_step"], d["clocks"]))
except Exception as e: print("$1 parse fail", e); print(open("gpurun_out/sl_$1.err").read()[-600:])
PYEOF
}
for S in 4 8 16; do run sp100_s$S $S superpoint "--images 100"; done
for S in 4 8; do run sift100_s$S $S sift "--images 100"; done
for S in 4 8; do run sift100_o50_s$S $S sift "--images 100 --outlier-frac 0.5"; done
run orb200_s4 4 orb "--images 200"; run orb200_s8 8 orb "--images 200"
