#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider -x -k "s8_ or float_tensor or superpoint or async" > gpurun_out/tests_s8.log 2>&1; echo "s8 tests exit $?"; tail -15 gpurun_out/tests_s8.log
for spec in s8:0 fp16:32768 s8b:0 fp16b:32768; do
  IFS=: read tag flags <<< "$spec"
  timeout 600 python bench.py --kind superpoint --images 64 --steps 3 --warmup 2 --no-cpu-baseline --no-stages --debug-flags $flags > gpurun_out/s8_$tag.json 2> gpurun_out/s8_$tag.err
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/s8_$tag.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$tag: value %.0f pairs/s ms/step %.1f | knn avg %.3f ms share %.3f frac %.3f | e2e %.0f | rerank %s | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["share_of_step"], r["frac"], d["e2e"]["value"], r.get("rerank"), d["clocks"]))
except Exception as e: print("$tag parse fail", e); print(open("gpurun_out/s8_$tag.err").read()[-600:])
PYEOF
done
