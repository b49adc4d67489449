#!/bin/bash
# real-valued tensor path: parity tests, SuperPoint bench, launch list
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -s"
timeout 900 $PY -k "float_tensor or superpoint" > gpurun_out/tests_float.log 2>&1; echo "float tests exit $?"; tail -25 gpurun_out/tests_float.log
for spec in "superpoint ${SP_IMAGES:-40} 0"; do
  set -- $spec
  timeout 900 python bench.py --kind $1 --images $2 --steps 3 --warmup 2 --no-cpu-baseline --debug-flags $3 > gpurun_out/bench_$1_$3.json 2> gpurun_out/bench_$1_$3.err; echo "bench $1 flags=$3 exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_$1_$3.json").read().strip().splitlines()[-1])
    print("$1 flags $3: value %.0f pairs/s ms/step %.1f roofline %s e2e %.0f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"], d["e2e"]["value"], d["clocks"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_$1_$3.err").read()[-1500:])
PYEOF
done
CMD="python bench.py --kind superpoint --images 24 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_launches_sp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sp.csv $CMD > gpurun_out/ncu_launches_sp.log 2>&1
echo "ncu launches exit $?"
