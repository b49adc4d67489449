#!/bin/bash
# mbarrier wait back-off (nanosleep) of the epilogue / MMA warps: A/B on the three descriptor families
source tools/r02/gpu_fn.sh
A="--images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e"
for lib in "" w0 w32 w96m0; do
  if [ -n "$lib" ]; then export PM_B200_LIB=$PWD/ab/libpm_$lib.so; else unset PM_B200_LIB; fi
  run wait_orb_${lib:-def} --kind orb $A
  run wait_sift_${lib:-def} --kind sift $A
  run wait_sp_${lib:-def} --kind superpoint $A
done
