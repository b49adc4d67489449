#!/bin/bash
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -k "fmat or pair_body or all_pairs or fountain or shim or min_matches" > gpurun_out/tests_ransac.log 2>&1; echo "ransac tests exit $?"; tail -5 gpurun_out/tests_ransac.log
source tools/r01/gpu_misc_fn.sh
for tag in old new; do
  if [ $tag = old ]; then export PM_B200_LIB=$PWD/ab/libpm_old.so; else unset PM_B200_LIB; fi
  run ${tag}_sift_out50 --kind sift --images 100 --steps 3 --warmup 2 --outlier-frac 0.5
  run ${tag}_sift_out29 --kind sift --images 100 --steps 3 --warmup 2 --outlier-frac 0.29
  run ${tag}_sift_out0 --kind sift --images 100 --steps 3 --warmup 2
done
