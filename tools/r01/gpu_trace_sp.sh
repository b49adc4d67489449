#!/bin/bash
mkdir -p gpurun_out
PM_TRACE=1 timeout 250 python bench.py --kind superpoint --images 64 --steps 1 --warmup 1 --no-e2e --no-stages --no-cpu-baseline > gpurun_out/trace_sp.json 2> gpurun_out/trace_sp.err; echo "trace exit $?"
PM_TRACE=1 timeout 250 python bench.py --kind sift --images 64 --steps 1 --warmup 1 --no-e2e --no-stages --no-cpu-baseline > gpurun_out/trace_sift.json 2> gpurun_out/trace_sift.err; echo "trace exit $?"
