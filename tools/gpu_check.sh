#!/bin/bash
# Runs on the GPU box via gpurun: parity tests in isolated processes (a faulting kernel poisons its
# CUDA context), then a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider"
echo "=== stage A: SIMT/popc/ransac (no tensor path)" 
timeout 900 $PY -k "hamming or superpoint or fmat or ratio_unique or errors or min_matches or match_descriptors or (force_simt and 1)" > gpurun_out/tests_A.log 2>&1; echo "A exit $?"; tail -5 gpurun_out/tests_A.log
echo "=== stage B: tensor path"
timeout 600 $PY -k "tc_accumulators" > gpurun_out/tests_B1.log 2>&1; echo "B1 exit $?"; tail -15 gpurun_out/tests_B1.log
timeout 900 $PY -k "sift and not full_size" > gpurun_out/tests_B2.log 2>&1; echo "B2 exit $?"; tail -15 gpurun_out/tests_B2.log
echo "=== stage C: everything else"
timeout 1200 $PY -k "not (hamming or superpoint or fmat or ratio_unique or errors or min_matches or match_descriptors or sift)" > gpurun_out/tests_C.log 2>&1; echo "C exit $?"; tail -15 gpurun_out/tests_C.log
timeout 900 $PY -k "full_size" > gpurun_out/tests_D.log 2>&1; echo "D exit $?"; tail -15 gpurun_out/tests_D.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "=== bench"
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
