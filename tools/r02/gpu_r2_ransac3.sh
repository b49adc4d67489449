#!/bin/bash
source tools/r02/gpu_fn.sh
PM_B200_LIB=$PWD/ab/libpm_prof.so python tools/ransac_prof.py 0.5 2>&1 | grep "RANSAC slot" | cut -c1-150
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -x -k "fmat or philox or eight_point or pair_body or fountain" > gpurun_out/r2_tests_rs.log 2>&1; echo "ransac tests exit $?"; tail -2 gpurun_out/r2_tests_rs.log
A="--kind sift --images 100 --steps 2 --warmup 1 --no-stages --no-configs --no-cpu-baseline --no-e2e"
PM_B200_LIB=$PWD/ab/libpm_rsp4.so run rs3_p4 $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_rsp8.so run rs3_p8 $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_rsold.so run rs3_old $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_rsp8.so run rs3_p8_of0 $A
