#!/bin/bash
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 120 -p no:cacheprovider"
timeout 300 $PY -k "sift_knn_bit_exact and 28" > gpurun_out/tests_pair1.log 2>&1; echo "pair test1 exit $?"; tail -15 gpurun_out/tests_pair1.log
timeout 600 $PY -k "sift or full_size" > gpurun_out/tests_pair2.log 2>&1; echo "pair test2 exit $?"; tail -8 gpurun_out/tests_pair2.log
for f in ${FLAG_LIST:-12 20 28 60}; do
  timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --debug-flags $f > gpurun_out/bench_f$f.log 2>gpurun_out/bench_f$f.err; echo "bench flags=$f exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_f$f.log").read().strip().splitlines()[-1])
    print("flags $f: value %.0f pairs/s  ms/step %.1f  roofline frac %.3f  avg knn launch %.3f ms share %.2f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["share_of_step"], d["clocks"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_f$f.err").read()[-800:])
PYEOF
done
