#!/bin/bash
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider"
timeout 900 $PY -k "sift or full_size or all_pairs" > gpurun_out/tests_tc.log 2>&1; echo "tc tests exit $?"; tail -4 gpurun_out/tests_tc.log
for f in ${FLAG_LIST:-12}; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --debug-flags $f > gpurun_out/bench_f$f.log 2>gpurun_out/bench_f$f.err; echo "bench flags=$f exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_f$f.log").read().strip().splitlines()[-1])
    print("flags $f: value %.0f pairs/s  ms/step %.1f  roofline frac %.3f  avg knn launch %.2f ms share %.2f e2e %s clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["share_of_step"], d["e2e"], d["clocks"]))
except Exception as e: print("parse fail", e)
PYEOF
done
BEST=${BEST_FLAGS:-12}
CMD2="python bench.py --images 23 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --debug-flags $BEST"
$CMD2 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc -s 1 -c 1 -o gpurun_out/prof_tc $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log
