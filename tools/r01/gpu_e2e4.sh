#!/bin/bash
mkdir -p gpurun_out
source tools/r01/gpu_misc_fn.sh
for rep in 1 2; do
  export PM_RESULT_COPY=ce
  run ce_$rep --steps 5 --warmup 3 --no-stages
  export PM_RESULT_COPY=kernel
  run k_$rep --steps 5 --warmup 3 --no-stages
done
