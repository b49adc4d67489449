#!/bin/bash
mkdir -p gpurun_out
source tools/r01/gpu_misc_fn.sh
export PM_SLOTS=8
run sp100_pre_s8 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
export PM_SLOTS=12
run sp100_pre_s12 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
export PM_SLOTS=8
PM_TRACE=1 timeout 250 python bench.py --kind superpoint --images 64 --steps 1 --warmup 1 --no-e2e --no-stages --no-cpu-baseline > gpurun_out/trace_sp8.json 2> gpurun_out/trace_sp8.err; echo "trace exit $?"
