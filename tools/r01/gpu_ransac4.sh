#!/bin/bash
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -k "fmat or pair_body or all_pairs or fountain or shim or min_matches or full_size or fast_path" > gpurun_out/tests_ransac.log 2>&1; echo "ransac tests exit $?"; tail -4 gpurun_out/tests_ransac.log
PM_B200_LIB=$PWD/ab/libpm_paranoid.so timeout 900 python tools/ransac_paranoid.py > gpurun_out/paranoid.log 2>&1; echo "paranoid exit $?"; grep -c MISMATCH gpurun_out/paranoid.log; tail -2 gpurun_out/paranoid.log
run() { # tag extra
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-stages --no-e2e $2 > gpurun_out/rs_$1.json 2> gpurun_out/rs_$1.err
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/rs_$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1: value %.0f pairs/s ms/step %.1f | knn avg %.3f ms share %.3f | inliers/step %d | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["share_of_step"], d["inliers_per_step"], d["clocks"]))
except Exception as e: print("$1 parse fail", e); print(open("gpurun_out/rs_$1.err").read()[-600:])
PYEOF
}
run o50 "--outlier-frac 0.5"
run o30 "--outlier-frac 0.3"
run o0 ""
run orb_o50 "--kind orb --images 100 --outlier-frac 0.5"
