#!/bin/bash
# A/B of i8 kernel builds (ab/libpm_*.so vs in-tree) + probes; SPECS = "tag:lib:flags ..."
mkdir -p gpurun_out
run() { # tag lib flags [extra bench args via EXTRA_<tag>]
  if [ -n "$2" ]; then export PM_B200_LIB=$PWD/$2; else unset PM_B200_LIB; fi
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-stages --debug-flags $3 $4 > gpurun_out/i8ab_$1.json 2> gpurun_out/i8ab_$1.err
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/i8ab_$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1: value %.0f pairs/s ms/step %.1f | knn avg %.3f ms share %.3f frac %.3f | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["share_of_step"], r["frac"], d["clocks"]))
except Exception as e: print("$1 parse fail", e); print(open("gpurun_out/i8ab_$1.err").read()[-600:])
PYEOF
}
for spec in $SPECS; do
  IFS=: read tag lib flags extra <<< "$spec"
  run $tag "$lib" $flags "${extra//,/ }"
done
