#!/bin/bash
# fp32 certified classifier of the RANSAC scoring loop: paranoid cross-check, parity tests, speed A/B
source tools/r02/gpu_fn.sh
PM_B200_LIB=$PWD/ab/libpm_rspar.so timeout 900 python tools/ransac_paranoid.py > gpurun_out/r2_paranoid.log 2>&1; echo "paranoid exit $?"; grep -c MISMATCH gpurun_out/r2_paranoid.log; grep -c "paranoid build active" gpurun_out/r2_paranoid.log; tail -3 gpurun_out/r2_paranoid.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -x -k "fmat or philox or eight_point or pair_body or fountain or match_all_pairs_equals or outliers" > gpurun_out/r2_tests_rs.log 2>&1; echo "ransac tests exit $?"; tail -3 gpurun_out/r2_tests_rs.log
A="--kind sift --images 100 --steps 2 --warmup 1 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run rs2_new $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_rsppt2.so run rs2_ppt2 $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_rsold.so run rs2_old $A --outlier-frac 0.5
run rs2_new_b1024 $A --outlier-frac 0.5 --batch-pairs 1024
run rs2_new_of0 $A
PM_B200_LIB=$PWD/ab/libpm_rsold.so run rs2_old_of0 $A
run rs2_new_of03 $A --outlier-frac 0.3
PM_B200_LIB=$PWD/ab/libpm_rsold.so run rs2_old_of03 $A --outlier-frac 0.3
