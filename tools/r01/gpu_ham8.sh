#!/bin/bash
# 256-bit rows on kind::i8 (two query row sets per cluster) vs the kind::f8f6f4 kernel: parity tests + A/B bench + probe
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -x -k "hamming or orb" > gpurun_out/tests_ham8.log 2>&1; echo "ham tests exit $?"; tail -12 gpurun_out/tests_ham8.log
source tools/r01/gpu_misc_fn.sh
run orb_fp4 --kind orb --images 100 --steps 3 --warmup 2
run orb_i8 --kind orb --images 100 --steps 3 --warmup 2 --debug-flags 262144
run orb_fp4_probe --kind orb --images 100 --steps 3 --warmup 2 --debug-flags 131072 --no-e2e --no-stages

