#!/bin/bash
# after scan_counts went to 256 threads: every GPU test, then the three families (SuperPoint with both re-rank variants)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/tests_gpu.log
source tools/r01/gpu_misc_fn.sh
run sf_sp100_pre --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
export PM_L2F_STAGED=1
run sf_sp100_staged --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
unset PM_L2F_STAGED
run sf_sift --steps 5 --warmup 3 --no-stages
run sf_orb --kind orb --images 100 --steps 3 --warmup 2 --no-stages
PM_TRACE=1 timeout 250 python bench.py --kind superpoint --images 64 --steps 1 --warmup 1 --no-e2e --no-stages --no-cpu-baseline > gpurun_out/trace_sp.json 2> gpurun_out/trace_sp.err; echo "trace exit $?"
