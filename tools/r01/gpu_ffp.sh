#!/bin/bash
# s8-prefilter re-rank: 256-thread blocks (16 K registers, in-tree) vs 128-thread blocks (8 K, ab/libpm_ffp128.so)
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -x -k "superpoint or float or s8 or fuzz" > gpurun_out/tests_ffp.log 2>&1; echo "sp tests exit $?"; tail -3 gpurun_out/tests_ffp.log
source tools/r01/gpu_misc_fn.sh
run ffp256 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
export PM_B200_LIB=$PWD/ab/libpm_ffp128.so
run ffp128 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
unset PM_B200_LIB
PM_TRACE=1 timeout 250 python bench.py --kind superpoint --images 64 --steps 1 --warmup 1 --no-e2e --no-stages --no-cpu-baseline > gpurun_out/trace_sp.json 2> gpurun_out/trace_sp.err; echo "trace exit $?"
