#!/bin/bash
mkdir -p gpurun_out
PM_TRACE=1 timeout 250 python tools/e2e_probe.py > gpurun_out/e2e_trace.out 2> gpurun_out/e2e_trace.err; echo "trace exit $?"
