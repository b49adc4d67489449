#!/bin/bash
# round 2: SuperPoint s8 candidates with two query row sets per cluster (l2_i8x2_kernel MODE 3): parity + A/B
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider"
timeout 1200 $PY -x -k "superpoint or float or s8 or fuzz or ragged or ratio_unique or edge_sizes or async" > gpurun_out/r2_tests_sp.log 2>&1; echo "sp tests exit $?"; tail -6 gpurun_out/r2_tests_sp.log
source tools/r01/gpu_misc_fn.sh
run r2_sp100_x2 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
run r2_sp100_one --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --debug-flags 524288
PM_L2F_PREFILTER=1 run r2_sp100_x2_pre --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
export PM_B200_LIB=$PWD/ab/libpm_s8st2.so
run r2_sp100_x2_st2 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
PM_L2F_PREFILTER=1 run r2_sp100_x2_st2_pre --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
unset PM_B200_LIB
# isolate the tensor kernel: a tiny ratio empties the tails
run r2_sp100_x2_knnonly --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e --dev-ratio 0.01 --dev-no-filter
run r2_sp100_one_knnonly --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e --dev-ratio 0.01 --dev-no-filter --debug-flags 524288
