#!/bin/bash
# staged filter with bounded grids over the list of handed-over pairs: every GPU test, paranoid cross-check, A/B in the easy case
source tools/r02/gpu_fn.sh
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -6 gpurun_out/r2_tests_gpu.log
PM_B200_LIB=$PWD/ab/libpm_paranoid.so timeout 900 python tools/ransac_paranoid.py > gpurun_out/r2_staged_paranoid.log 2>&1; echo "paranoid exit $?"
echo "mismatch lines: $(grep -c MISMATCH gpurun_out/r2_staged_paranoid.log)  active: $(grep -c 'paranoid build active' gpurun_out/r2_staged_paranoid.log)"
A="--kind sift --images 100 --steps 5 --warmup 3 --no-stages --no-configs --no-cpu-baseline"
run st3_of0_a $A
run st3_of0_1k_a $A --debug-flags 2097152
run st3_of0_b $A
run st3_of0_1k_b $A --debug-flags 2097152
run st3_heavy $A --outlier-frac 0.5 --no-e2e
run st3_of03 $A --outlier-frac 0.3 --no-e2e
run st3_of02 $A --outlier-frac 0.2 --no-e2e
run st3_of02_1k $A --outlier-frac 0.2 --no-e2e --debug-flags 2097152
