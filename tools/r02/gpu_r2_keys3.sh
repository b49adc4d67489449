#!/bin/bash
# round 2: argmin epilogue (MODE 4) + one-column re-rank for SuperPoint: parity + speed; same-box A/B against the previous build
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider"
timeout 1200 $PY -x -k "superpoint or float or s8 or fuzz or ragged or ratio_unique or edge_sizes or async or fmat or fountain or golden" > gpurun_out/r2_tests_sp.log 2>&1; echo "sp tests exit $?"; tail -6 gpurun_out/r2_tests_sp.log
source tools/r01/gpu_misc_fn.sh
run r2k_sp100 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
run r2k_sp100_six --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e --debug-flags 1048576
run r2k_sp100_knnonly --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e --dev-ratio 0.01 --dev-no-filter
run r2k_sp100_out --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e --outlier-frac 0.3
PM_B200_LIB=$PWD/ab/libpm_w0.so run r2k_sp100_old --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
run r2k_sift100 --kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
