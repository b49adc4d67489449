#!/bin/bash
# quick iteration: tensor-path parity tests (integer + real-valued) and both L2 benches
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider"
timeout 900 $PY -k "${TEST_K:-float_tensor or superpoint or fast_path or full_size}" > gpurun_out/tests_iter.log 2>&1; echo "iter tests exit $?"; tail -6 gpurun_out/tests_iter.log
for spec in ${SPECS:-superpoint:40:0 sift:100:0}; do
  IFS=: read kind images flags <<< "$spec"
  timeout 900 python bench.py --kind $kind --images $images --steps 3 --warmup 2 --no-cpu-baseline --debug-flags $flags > gpurun_out/bench_${kind}_$flags.json 2> gpurun_out/bench_${kind}_$flags.err; echo "bench $kind flags=$flags exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_${kind}_$flags.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$kind flags $flags: value %.0f pairs/s ms/step %.1f | knn %.1f %s frac %.3f avg %.3f ms share %s | e2e %.0f | %s %s" % (d["value"], d["ms_per_step"], r["achieved"], r["unit"], r["frac"], r["avg_launch_ms"], r["share_of_step"], d["e2e"]["value"], d["clocks"], r.get("rerank")))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_${kind}_$flags.err").read()[-1500:])
PYEOF
done
