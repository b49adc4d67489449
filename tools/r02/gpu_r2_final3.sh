#!/bin/bash
# closing evidence, last build of the round (packed-pair ORB kernel): every GPU test, smoke, both bench arms with the default command
source tools/r02/gpu_fn.sh
nvidia-smi -L; nproc
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/r2_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
T0=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref exit $? in $(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 1500 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/r2_bench_default.err
python - <<'PYEOF'
import json
d=json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("main: value %.0f ms/step %.2f e2e %.0f | roof achieved %.0f peak %.0f frac %.3f bf16frac %.3f share %.3f launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["achieved"], r["peak"], r["frac"], r["frac_of_bf16_sustained"], r["share_of_step"], d["gpu_launches"]))
print("stages:", json.dumps(d["stages"])[:1000])
for k,c in d["configs"].items():
    rr=c["roofline"]
    print(k, "value %.0f ms/step %.1f e2e %.0f | achieved %.0f peak %.0f frac %.3f share %.3f setup %s" % (c["value"], c["ms_per_step"], c["e2e"]["value"], rr["achieved"], rr["peak"], rr["frac"], rr["share_of_step"], c["setup_s"]), c["clocks"])
print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PYEOF
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2_launches_bench_default.log 2>&1; echo "ncu list exit $?"
run orb100 --kind orb --images 100 --steps 5 --warmup 3 --no-configs --no-cpu-baseline
B="python bench.py --kind orb --images 48 --steps 1 --warmup 1 --no-stages --no-e2e --no-cpu-baseline --no-configs"
ncu --set full --clock-control none --import-source on -k regex:l2_i8x2_kernel -s 6 -c 1 -f -o gpurun_out/r2_prof_pk $B > gpurun_out/r2_prof_pk.log 2>&1; echo "ncu pk $?"
