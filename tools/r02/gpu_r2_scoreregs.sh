#!/bin/bash
# register budget of rs_score_kernel (blocks next to the persistent kNN kernel) + compute-sanitizer on the staged filter
source tools/r02/gpu_fn.sh
A="--kind sift --images 100 --steps 3 --warmup 2 --no-stages --no-configs --no-cpu-baseline --no-e2e --outlier-frac 0.5"
run sr_64 $A
PM_B200_LIB=$PWD/ab/libpm_rs48.so run sr_48 $A
PM_B200_LIB=$PWD/ab/libpm_rs40.so run sr_40 $A
PM_B200_LIB=$PWD/ab/libpm_rs80.so run sr_80 $A
PM_SLOTS=8 run sr_64_slots8 $A
run sr_64_b512 $A --batch-pairs 512
cat > /tmp/san.py <<'PYEOF'
import sys, numpy as np
sys.path.insert(0, ".")
from reconstructor_b200 import api
g = np.load("tests/golden/fmat_scenes.npz")
for sampler in (0, 1):
    for resid in (0, 1):
        with api.PairMatcher(sampler=sampler, residual_mode=resid, seed=3) as pm:
            for k in (22, 30, 38, 46, 51):
                F, mask, st, it = pm.estimate_fundamental(g[f"s{k}_p1"], g[f"s{k}_p2"])
                print(sampler, resid, k, st, it, int(mask.sum()))
PYEOF
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/r2_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r2_sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"; tail -3 gpurun_out/r2_sanitizer_synccheck.log
