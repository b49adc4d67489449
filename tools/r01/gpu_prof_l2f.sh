#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --kind superpoint --images 46 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-stages"
$CMD > gpurun_out/plain_sp46.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:l2f_fixup -s 4 -c 1 -f -o gpurun_out/prof_l2f_new $CMD > gpurun_out/ncu_full_l2f.log 2>&1; echo "ncu full l2f exit $?"
