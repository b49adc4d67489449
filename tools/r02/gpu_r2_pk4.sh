#!/bin/bash
# packed-pair kernel with ONE norm K-step for both rows (pair norm blocks, per-K-block scale factors): where does
# (historical: the -DPM_PK_SFMODE build switch behind ab/libpm_pk_sf*.so was removed once layout 0 -- bytes 0 | 1 of a column -- proved to be the one)
# scale_vec::2X read the two scales of a row?  tests with each candidate layout, then the ORB-100 A/B with the one that passes
source tools/r02/gpu_fn.sh
T="timeout 600 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x -k test_hamming_i8_two_set_kernel_sizes_around_row_sets"
ok=""
for m in 0 1 2 3; do
  if [ $m -eq 0 ]; then unset PM_B200_LIB; else export PM_B200_LIB=ab/libpm_pk_sf$m.so; fi
  $T > gpurun_out/r2_pk4_tests_m$m.log 2>&1; rc=$?; echo "scale layout $m exit $rc"; tail -2 gpurun_out/r2_pk4_tests_m$m.log
  if [ $rc -eq 0 ]; then ok=$m; break; fi
done
echo "passing layout: ${ok:-none}"
[ -z "$ok" ] && unset PM_B200_LIB
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -k "hamming or orb or bits" > gpurun_out/r2_pk4_tests.log 2>&1; echo "binary-row tests exit $? (lib ${PM_B200_LIB:-default})"; tail -4 gpurun_out/r2_pk4_tests.log
A="--kind orb --images 100 --steps 3 --warmup 3 --no-stages --no-configs --no-cpu-baseline"
run pk4_orb100 $A
