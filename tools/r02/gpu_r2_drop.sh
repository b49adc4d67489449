#!/bin/bash
# exact early drop in rs_score_kernel + one block per pair in the per-pair staged kernels: RANSAC tests, heavy / mid workloads
source tools/r02/gpu_fn.sh
timeout 1200 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -k "fmat or philox or eight_point or pair_body or fountain or staged or essential" > gpurun_out/r2_tests_rs.log 2>&1; echo "ransac tests exit $?"; tail -3 gpurun_out/r2_tests_rs.log
A="--kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run drop_heavy $A --outlier-frac 0.5
run drop_of03 $A --outlier-frac 0.3
run drop_of02 $A --outlier-frac 0.2
run drop_of0 $A
