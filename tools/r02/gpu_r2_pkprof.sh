#!/bin/bash
# ncu --set full of ONE steady-state 256-pair launch of the packed-pair kind::mxf4 kernel (ORB, 100 images: batches 16/32/64/128/256..)
mkdir -p gpurun_out
B="python bench.py --kind orb --images 100 --steps 1 --warmup 1 --no-stages --no-e2e --no-cpu-baseline --no-configs"
$B > gpurun_out/r2_prof_pk256_plain.log 2>&1; echo "plain $?"; tail -c 600 gpurun_out/r2_prof_pk256_plain.log
ncu --set full --clock-control none --import-source on -k regex:l2_i8x2_kernel -s 6 -c 1 -f -o gpurun_out/r2_prof_pk256 $B > gpurun_out/r2_prof_pk256.log 2>&1; echo "ncu $?"
PM_TRACE=1 $B 2>&1 | grep -i "batch\|pairs" | head -12
