#!/bin/bash
# per-kernel durations (ncu launch list, serialised): current build, SIFT and SuperPoint
mkdir -p gpurun_out
B="python bench.py --images 48 --steps 1 --warmup 1 --no-stages --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launch_sift.csv $B --kind sift > gpurun_out/r2_launch_sift.log 2>&1; echo "sift $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launch_keys3.csv $B --kind superpoint > gpurun_out/r2_launch_keys3.log 2>&1; echo "keys3 $?"
