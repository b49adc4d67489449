#!/bin/bash
# every GPU test on the current build + the three families' benches (same box)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -15 gpurun_out/r2_tests_gpu.log
source tools/r01/gpu_misc_fn.sh
run r2f_sp100 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
run r2f_sp100_knnonly --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e --dev-ratio 0.01 --dev-no-filter
PM_B200_LIB=$PWD/ab/libpm_w0.so run r2f_sp100_old --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
run r2f_sift100 --kind sift --images 100 --steps 3 --warmup 2 --no-stages
run r2f_orb100 --kind orb --images 100 --steps 3 --warmup 2 --no-stages
