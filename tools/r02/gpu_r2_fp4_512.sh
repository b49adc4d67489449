#!/bin/bash
# 512-bit binary rows on kind::mxf4 + block-cyclic shares: every GPU test, the 512-bit A/B, default configs
source tools/r02/gpu_fn.sh
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -6 gpurun_out/r2_tests_gpu.log
python tools/r02/orb512_probe.py 2>&1 | tail -4
run f4_orb100 --kind orb --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline
