#!/bin/bash
# every GPU test on the current build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/tests_gpu.log
