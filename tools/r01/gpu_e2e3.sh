#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/tests_gpu.log
timeout 250 python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "e2e probe exit $?"; tail -12 gpurun_out/e2e_probe.log
source tools/r01/gpu_misc_fn.sh
run sift_e2e --steps 5 --warmup 3 --no-stages
run orb_e2e --kind orb --images 100 --steps 3 --warmup 2 --no-stages
run sp_e2e --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
