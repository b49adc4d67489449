#!/bin/bash
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider"
timeout 600 $PY -k "fast_path or all_pairs or fountain or full_size or pair_body" > gpurun_out/tests_fast.log 2>&1; echo "fast tests exit $?"; tail -12 gpurun_out/tests_fast.log
for f in ${FLAG_LIST:-0 64}; do
  timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --debug-flags $f > gpurun_out/bench_f$f.log 2>gpurun_out/bench_f$f.err; echo "bench flags=$f exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_f$f.log").read().strip().splitlines()[-1])
    print("flags $f: value %.0f pairs/s  ms/step %.1f  roofline frac %.3f  avg knn launch %.3f ms share %.2f e2e %.0f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["share_of_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_f$f.err").read()[-800:])
PYEOF
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
