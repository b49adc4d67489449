#!/bin/bash
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -k "hamming or orb or ties or ragged" > gpurun_out/tests_ham.log 2>&1; echo "ham tests exit $?"; tail -8 gpurun_out/tests_ham.log
source tools/r01/gpu_misc_fn.sh
run orb_tensor --kind orb --images 100 --steps 3 --warmup 2
run orb_popc --kind orb --images 60 --steps 2 --warmup 1 --debug-flags 1024
