#!/bin/bash
mkdir -p gpurun_out
PM_TRACE=1 timeout 250 python tools/e2e_probe.py > gpurun_out/e2e_trace.out 2> gpurun_out/e2e_trace.err; echo "trace exit $?"
CMD="python bench.py --kind orb --images 46 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
$CMD > gpurun_out/plain_orb46.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:l2_i8x2 -s 4 -c 1 -f -o gpurun_out/prof_fp4 $CMD > gpurun_out/ncu_full_fp4.log 2>&1; echo "ncu full fp4 exit $?"
