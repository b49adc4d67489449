#!/bin/bash
# re-rank with the s8 prefilter (three blocks next to the tensor kernel): parity tests + A/B against the staged variant
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -x -k "superpoint or float or s8 or fuzz or ragged or ratio_unique" > gpurun_out/tests_sp_pre.log 2>&1; echo "sp tests exit $?"; tail -6 gpurun_out/tests_sp_pre.log
source tools/r01/gpu_misc_fn.sh
run sp100_pre --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
export PM_L2F_STAGED=1
run sp100_staged --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages
unset PM_L2F_STAGED
run sp64_pre --kind superpoint --images 64 --steps 3 --warmup 2 --no-stages
