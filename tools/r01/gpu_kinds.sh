#!/bin/bash
mkdir -p gpurun_out
for spec in "orb 60 0" "orb 60 2" "superpoint 16 0"; do
  set -- $spec
  timeout 900 python bench.py --kind $1 --images $2 --steps 2 --warmup 1 --cpu-seconds 6 --debug-flags $3 > gpurun_out/bench_$1_$3.json 2> gpurun_out/bench_$1_$3.err; echo "bench $1 flags=$3 exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/bench_$1_$3.json").read().strip().splitlines()[-1])
    print("$1 flags $3: value %.0f pairs/s ms/step %.1f roofline %s cpu %s e2e %.0f" % (d["value"], d["ms_per_step"], {k:d["roofline"][k] for k in ("achieved","peak","frac","unit","avg_launch_ms","share_of_step")}, d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1), d["e2e"]["value"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_$1_$3.err").read()[-600:])
PYEOF
done
