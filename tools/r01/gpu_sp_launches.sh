#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_sp64.csv python bench.py --kind superpoint --images 64 --steps 1 --warmup 0 --no-e2e --no-stages --no-cpu-baseline > gpurun_out/ncu_launches_sp.log 2>&1; echo "ncu exit $?"
