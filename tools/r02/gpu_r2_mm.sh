#!/bin/bash
# model-major fp32 scoring of the RANSAC kernel: paranoid cross-check, per-phase cycles, all GPU tests, A/B against the match-major build
source tools/r02/gpu_fn.sh
PM_B200_LIB=$PWD/ab/libpm_paranoid.so timeout 900 python tools/ransac_paranoid.py > gpurun_out/r2_mm_paranoid.log 2>&1; echo "paranoid exit $?"
echo "mismatch lines: $(grep -c MISMATCH gpurun_out/r2_mm_paranoid.log)  active: $(grep -c 'paranoid build active' gpurun_out/r2_mm_paranoid.log)"; grep -v "paranoid build active" gpurun_out/r2_mm_paranoid.log | tail -8 | cut -c1-200
PM_B200_LIB=$PWD/ab/libpm_prof.so python tools/ransac_prof.py 0.5 2>&1 | grep "RANSAC slot" | cut -c1-220
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -6 gpurun_out/r2_tests_gpu.log
A="--kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run mm_heavy $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_rsold.so run mm_old_heavy $A --outlier-frac 0.5
run mm_of03 $A --outlier-frac 0.3
PM_B200_LIB=$PWD/ab/libpm_rsold.so run mm_old_of03 $A --outlier-frac 0.3
run mm_of0 $A
PM_B200_LIB=$PWD/ab/libpm_rsold.so run mm_old_of0 $A
