#!/bin/bash
# state check after the container was re-created: every GPU test, then the default bench line (both arms)
source tools/r02/gpu_fn.sh
nvidia-smi -L; nproc; free -g | head -2
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -6 gpurun_out/r2_tests_gpu.log
T0=$(date +%s); timeout 1500 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"; tail -5 gpurun_out/r2_bench_default.err
python - <<'PYEOF'
import json
d=json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("main: value %.0f ms/step %.2f e2e %.0f | roof achieved %.0f peak %.0f frac %.3f bf16frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["achieved"], r["peak"], r["frac"], r["frac_of_bf16_sustained"], r["share_of_step"]))
print("stages:", json.dumps(d["stages"])[:900])
for k,c in d["configs"].items():
    rr=c["roofline"]
    print(k, "value %.0f ms/step %.1f e2e %.0f | achieved %.0f peak %.0f frac %.3f share %.3f setup %s" % (c["value"], c["ms_per_step"], c["e2e"]["value"], rr["achieved"], rr["peak"], rr["frac"], rr["share_of_step"], c["setup_s"]), c["clocks"])
print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PYEOF
T0=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref exit $? in $(( $(date +%s) - T0 )) s"; cat gpurun_out/r2_bench_reference.json | cut -c1-400
