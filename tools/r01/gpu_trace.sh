#!/bin/bash
mkdir -p gpurun_out
PM_TRACE=1 timeout 600 python bench.py --kind superpoint --images 64 --steps 1 --warmup 1 --no-cpu-baseline --no-stages --no-e2e > gpurun_out/trace_sp.json 2> gpurun_out/trace_sp.err
grep "pm trace" gpurun_out/trace_sp.err | tail -8
PM_TRACE=1 timeout 600 python bench.py --kind sift --images 64 --steps 1 --warmup 1 --no-cpu-baseline --no-stages --no-e2e > gpurun_out/trace_sift.json 2> gpurun_out/trace_sift.err
grep "pm trace" gpurun_out/trace_sift.err | tail -8
