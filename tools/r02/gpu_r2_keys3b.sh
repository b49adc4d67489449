#!/bin/bash
# same-box A/B: previous build (two row sets, six keys, chunk re-rank) vs this build with flag 1<<20 vs the argmin path
source tools/r01/gpu_misc_fn.sh
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv
export PM_B200_LIB=$PWD/ab/libpm_w0.so
run r2kb_old --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
unset PM_B200_LIB
run r2kb_six --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e --debug-flags 1048576
run r2kb_keys3 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
export PM_B200_LIB=$PWD/ab/libpm_w0.so
run r2kb_old2 --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
