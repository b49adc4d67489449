#!/bin/bash
# packed pairs of train rows in the kind::mxf4 Hamming kernel (l2_i8x2_kernel PK): binary-row GPU tests, ORB-100 A/B
source tools/r02/gpu_fn.sh
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -k "hamming or orb or bits" > gpurun_out/r2_pk_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/r2_pk_tests.log
run pk_orb100 --kind orb --images 100 --steps 3 --warmup 3 --no-stages --no-configs --no-cpu-baseline
run pk_orb100_unpacked --kind orb --images 100 --steps 3 --warmup 3 --no-stages --no-configs --no-cpu-baseline --debug-flags 67108864
run pk_orb100_probe --kind orb --images 100 --steps 3 --warmup 3 --no-stages --no-configs --no-cpu-baseline --no-e2e --debug-flags 131072
