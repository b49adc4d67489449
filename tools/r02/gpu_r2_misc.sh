#!/bin/bash
# leftovers: batches in flight and the stage cut of the filter (A/B), config #5 sweep points with the final build
source tools/r02/gpu_fn.sh
A="--images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run misc_heavy --kind sift $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_cut24.so run misc_heavy_cut24 --kind sift $A --outlier-frac 0.5
PM_B200_LIB=$PWD/ab/libpm_cut24.so run misc_of0_cut24 --kind sift $A
PM_B200_LIB=$PWD/ab/libpm_cut24.so run misc_of02_cut24 --kind sift $A --outlier-frac 0.2
PM_SLOTS=6 run misc_heavy_s6 --kind sift $A --outlier-frac 0.5
PM_SLOTS=6 run misc_sift_s6 --kind sift $A
PM_SLOTS=6 run misc_sp_s6 --kind superpoint $A
PM_SLOTS=6 run misc_orb_s6 --kind orb $A
W="--steps 2 --warmup 1 --no-cpu-baseline --no-stages --no-e2e --no-configs"
run sweep_sift_64x16384 --kind sift --images 64 --kp 16384 $W
run sweep_orb_64x16384 --kind orb --images 64 --kp 16384 $W
run sweep_orb_2000x8192 --kind orb --images 2000 --kp 8192 --max-pairs 40000 $W
run sweep_sift_2000x4096 --kind sift --images 2000 --kp 4096 --max-pairs 40000 $W
