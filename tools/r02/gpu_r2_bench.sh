#!/bin/bash
# the default bench line (with configs) + the GPU tests that failed / are new
source tools/r02/gpu_fn.sh


T0=$(date +%s); python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"; tail -5 gpurun_out/r2_bench_default.err
python - <<'PYEOF'
import json
d=json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("main: value %.0f ms/step %.2f e2e %.0f | roof achieved %.0f peak %.0f frac %.3f bf16frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["achieved"], r["peak"], r["frac"], r["frac_of_bf16_sustained"], r["share_of_step"]))
print("stages:", json.dumps(d["stages"])[:600])
for k,c in d["configs"].items():
    rr=c["roofline"]
    print(k, "value %.0f ms/step %.1f e2e %.0f | achieved %.0f peak %.0f frac %.3f share %.3f setup %s" % (c["value"], c["ms_per_step"], c["e2e"]["value"], rr["achieved"], rr["peak"], rr["frac"], rr["share_of_step"], c["setup_s"]), c["clocks"])
print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PYEOF
