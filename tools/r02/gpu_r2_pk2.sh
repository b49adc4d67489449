#!/bin/bash
# packed-pair kernel: what bounds the epilogue?  tails off (no filter), epilogue polling interval, ncu --set full of one launch
source tools/r02/gpu_fn.sh
A="--kind orb --images 100 --steps 3 --warmup 3 --no-stages --no-configs --no-cpu-baseline --no-e2e"
run pk2_nofilter $A --dev-no-filter
run pk2_unpacked_nofilter $A --dev-no-filter --debug-flags 67108864
PM_B200_LIB=ab/libpm_pk_w0.so run pk2_w0 $A
PM_B200_LIB=ab/libpm_pk_w32.so run pk2_w32 $A
B="python bench.py --kind orb --images 48 --steps 1 --warmup 1 --no-stages --no-e2e --no-cpu-baseline --no-configs"
ncu --set full --clock-control none --import-source on -k regex:l2_i8x2_kernel -s 6 -c 1 -f -o gpurun_out/r2_prof_pk $B > gpurun_out/r2_prof_pk.log 2>&1; echo "ncu $?"
ls -la gpurun_out/r2_prof_pk.ncu-rep
