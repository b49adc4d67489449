#!/bin/bash
# round 2: sleeping mbarrier waits in the persistent tensor kernels -- do the tail kernels hide now?  A/B of sleep lengths
source tools/r01/gpu_misc_fn.sh
for v in main w0 w1 w2 w3; do
  if [ $v = main ]; then unset PM_B200_LIB; else export PM_B200_LIB=$PWD/ab/libpm_$v.so; fi
  run r2w_sp_$v --kind superpoint --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
  run r2w_sift_$v --kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
  run r2w_orb_$v --kind orb --images 100 --steps 3 --warmup 2 --no-stages --no-e2e
done
