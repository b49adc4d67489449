#!/bin/bash
mkdir -p gpurun_out
source tools/r01/gpu_misc_fn.sh
export CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_sp --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
run conn32_sift --steps 4 --warmup 3 --no-stages
