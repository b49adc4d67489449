#!/bin/bash
# unit-norm form of the s8 search (no norm K-step) + staged-filter scheduling hint: all GPU tests, SuperPoint A/B, SIFT check
source tools/r02/gpu_fn.sh
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -6 gpurun_out/r2_tests_gpu.log
A="--steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline"
run nx_sp100 --kind superpoint --images 100 $A
run nx_sp100_norm --kind superpoint --images 100 $A --debug-flags 8388608
run nx_sift_of0 --kind sift --images 100 $A
run nx_sift_of0_1k --kind sift --images 100 $A --debug-flags 2097152
run nx_sift_heavy --kind sift --images 100 $A --outlier-frac 0.5 --no-e2e
run nx_sp100_b --kind superpoint --images 100 $A
