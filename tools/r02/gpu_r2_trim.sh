#!/bin/bash
# trimmed fast path of rs_score_kernel: RANSAC tests, paranoid cross-check, heavy / mid workloads
source tools/r02/gpu_fn.sh
timeout 1200 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -k "fmat or philox or eight_point or pair_body or fountain or staged or essential" > gpurun_out/r2_tests_rs.log 2>&1; echo "ransac tests exit $?"; tail -3 gpurun_out/r2_tests_rs.log
PM_B200_LIB=$PWD/ab/libpm_paranoid.so timeout 900 python tools/ransac_paranoid.py > gpurun_out/r2_staged_paranoid.log 2>&1; echo "paranoid exit $?"
echo "mismatch lines: $(grep -c MISMATCH gpurun_out/r2_staged_paranoid.log)  active: $(grep -c 'paranoid build active' gpurun_out/r2_staged_paranoid.log)"
A="--kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run trim_heavy $A --outlier-frac 0.5
run trim_of03 $A --outlier-frac 0.3
run trim_heavy_b $A --outlier-frac 0.5
