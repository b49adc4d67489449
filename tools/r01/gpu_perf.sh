#!/bin/bash
# Perf round on the GPU box: TC parity with every epilogue variant, bench per variant, ncu launch
# list of the bench and one ncu --set full capture of the tensor kernel.
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider"
timeout 900 $PY > gpurun_out/tests_tc.log 2>&1; echo "tc tests exit $?"; tail -4 gpurun_out/tests_tc.log
for f in ${FLAG_LIST:-4 12}; do
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --debug-flags $f > gpurun_out/bench_f$f.log 2>gpurun_out/bench_f$f.err; echo "bench flags=$f exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_f$f.log").read().strip().splitlines()[-1])
    print("flags $f: value %.0f pairs/s  ms/step %.1f  roofline frac %.3f  avg knn launch %.2f ms share %.2f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["share_of_step"], d["clocks"]))
except Exception as e: print("parse fail", e)
PYEOF
done
BEST=${BEST_FLAGS:-12}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --debug-flags $BEST"
$CMD > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
CMD2="python bench.py --images 23 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --debug-flags $BEST"
$CMD2 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc -s 1 -c 1 -o gpurun_out/prof_tc $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
