#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/tests_gpu.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/tests_gpu.log
for mode in "" "--sync-ingest"; do
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-stages $mode > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_e2e.json").read().strip().splitlines()[-1])
    print("mode '$mode': value %.0f ms/step %.1f | e2e %s" % (d["value"], d["ms_per_step"], d["e2e"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_e2e.err").read()[-800:])
PYEOF
done
