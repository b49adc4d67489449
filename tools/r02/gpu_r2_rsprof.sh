#!/bin/bash
# per-phase cycle counts of the RANSAC kernel (PM_RANSAC_PROFILE build) under 50 % outliers; variant without the literal path
mkdir -p gpurun_out
PM_B200_LIB=$PWD/ab/libpm_prof.so python tools/ransac_prof.py 0.5 2>&1 | grep "RANSAC slot" | cut -c1-150
PM_B200_LIB=$PWD/ab/libpm_prof1.so python tools/ransac_prof.py 0.5 2>&1 | grep "RANSAC slot" | cut -c1-150
