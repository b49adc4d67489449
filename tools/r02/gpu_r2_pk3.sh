#!/bin/bash
# packed-pair kernel, timing probes of the epilogue (wrong results): no field shift / XOR tree instead of the 16x2 minima
source tools/r02/gpu_fn.sh
A="--kind orb --images 100 --steps 3 --warmup 3 --no-stages --no-configs --no-cpu-baseline --no-e2e"
PM_B200_LIB=ab/libpm_pk_p1.so run pk3_noshift $A
PM_B200_LIB=ab/libpm_pk_p2.so run pk3_xortree $A
run pk3_default $A
