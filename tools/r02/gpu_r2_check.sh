#!/bin/bash
# every GPU test on the current build (+ the shim), then a quick SP / SIFT sanity bench
source tools/r02/gpu_fn.sh
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -6 gpurun_out/r2_tests_gpu.log
run sp100 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline
run sift100 --kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline
