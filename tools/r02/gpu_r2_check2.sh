#!/bin/bash
# after the last scheduling changes (hand-over at 24 iterations, 6 slots, shrinking last batches): every GPU test + bench lines
source tools/r02/gpu_fn.sh
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/r2_tests_gpu.log
A="--images 100 --steps 5 --warmup 3 --no-stages --no-configs --no-cpu-baseline"
run c2_sift --kind sift $A
PM_RAMPDOWN=0 run c2_sift_noramp --kind sift $A
run c2_sift_b --kind sift $A
PM_RAMPDOWN=0 run c2_sift_noramp_b --kind sift $A
run c2_heavy --kind sift $A --outlier-frac 0.5 --no-e2e
run c2_sp --kind superpoint $A
run c2_orb --kind orb $A
