#!/bin/bash
# RANSAC under 50 % outliers: batch size, slots in flight, register cap of the RANSAC kernel
source tools/r02/gpu_fn.sh
A="--kind sift --images 100 --steps 2 --warmup 1 --no-stages --no-configs --no-cpu-baseline --no-e2e --outlier-frac 0.5"
run rs_base $A
run rs_b512 $A --batch-pairs 512
run rs_b1024 $A --batch-pairs 1024
PM_SLOTS=8 run rs_s8 $A
PM_SLOTS=8 run rs_s8_b512 $A --batch-pairs 512
PM_B200_LIB=$PWD/ab/libpm_rsmb6.so run rs_mb6 $A
PM_B200_LIB=$PWD/ab/libpm_rsmb8.so run rs_mb8 $A
PM_SLOTS=8 PM_B200_LIB=$PWD/ab/libpm_rsmb8.so run rs_mb8_s8 $A
