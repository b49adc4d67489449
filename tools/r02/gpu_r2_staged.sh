#!/bin/bash
# staged continuation of the epipolar filter (mega-rounds): parity, paranoid cross-check, A/B against the per-pair kernel
source tools/r02/gpu_fn.sh
timeout 1200 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -k "fmat or philox or eight_point or pair_body or fountain or staged or essential" > gpurun_out/r2_tests_rs.log 2>&1; echo "ransac tests exit $?"; tail -12 gpurun_out/r2_tests_rs.log
PM_B200_LIB=$PWD/ab/libpm_paranoid.so timeout 900 python tools/ransac_paranoid.py > gpurun_out/r2_staged_paranoid.log 2>&1; echo "paranoid exit $?"
echo "mismatch lines: $(grep -c MISMATCH gpurun_out/r2_staged_paranoid.log)  active: $(grep -c 'paranoid build active' gpurun_out/r2_staged_paranoid.log)"; grep -v "paranoid build active" gpurun_out/r2_staged_paranoid.log | tail -4 | cut -c1-200
A="--kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run st_heavy $A --outlier-frac 0.5
run st_heavy_1k $A --outlier-frac 0.5 --debug-flags 2097152
PM_B200_LIB=$PWD/ab/libpm_rs48.so run st_heavy_48 $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_rs80.so run st_heavy_80 $A --outlier-frac 0.5
run st_of03 $A --outlier-frac 0.3
run st_of03_1k $A --outlier-frac 0.3 --debug-flags 2097152
run st_of0 $A
run st_of0_1k $A --debug-flags 2097152
run st_heavy_b512 $A --outlier-frac 0.5 --batch-pairs 512
