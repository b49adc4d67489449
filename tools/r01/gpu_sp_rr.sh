#!/bin/bash
# re-rank with long segments: SuperPoint parity tests + bench
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -x -k "superpoint or float or s8 or fuzz or ragged or ratio_unique" > gpurun_out/tests_sp_rr.log 2>&1; echo "sp tests exit $?"; tail -6 gpurun_out/tests_sp_rr.log
source tools/r01/gpu_misc_fn.sh
run sp64 --kind superpoint --images 64 --steps 3 --warmup 2
run sp100 --kind superpoint --images 100 --steps 3 --warmup 2
