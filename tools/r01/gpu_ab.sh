#!/bin/bash
# A/B of two library builds on the same box: ab/libpm_old.so vs the in-tree build, interleaved
mkdir -p gpurun_out
for rep in 1 2; do
for tag in old new; do
  if [ $tag = old ]; then export PM_B200_LIB=$PWD/ab/libpm_old.so; else unset PM_B200_LIB; fi
  for spec in ${SPECS:-sift:100:0 sift:100:512 superpoint:64:0}; do
    IFS=: read kind images flags <<< "$spec"
    timeout 900 python bench.py --kind $kind --images $images --steps 4 --warmup 3 --no-cpu-baseline --debug-flags $flags > gpurun_out/ab_${tag}_${kind}_${flags}.json 2> gpurun_out/ab_${tag}_${kind}_${flags}.err
    python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/ab_${tag}_${kind}_${flags}.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$tag $kind $flags: value %.0f pairs/s ms/step %.1f | knn %.1f frac %.3f avg %.3f ms share %.3f | e2e %.0f | %s" % (d["value"], d["ms_per_step"], r["achieved"], r["frac"], r["avg_launch_ms"], r["share_of_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/ab_${tag}_${kind}_${flags}.err").read()[-1500:])
PYEOF
  done
done
done
