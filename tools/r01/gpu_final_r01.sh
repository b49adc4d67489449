#!/bin/bash
# Round-end evidence: every GPU test, smoke, default bench (both arms), launch lists, ncu --set full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" > gpurun_out/cpu_info.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/bench_default.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sift_i8.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches sift exit $?"
CMDS="python bench.py --kind superpoint --images 24 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_superpoint_s8.csv $CMDS > gpurun_out/ncu_launches_sp.log 2>&1; echo "ncu launches superpoint exit $?"
CMDS2="python bench.py --kind superpoint --images 23 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
$CMDS2 > gpurun_out/plain_sp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc2 -s 1 -c 1 -f -o gpurun_out/prof_s8 $CMDS2 > gpurun_out/ncu_full_s8.log 2>&1; echo "ncu full s8 exit $?"
ncu --set full --clock-control none --import-source on -k regex:l2f_fixup -s 1 -c 1 -f -o gpurun_out/prof_l2f_s8 $CMDS2 > gpurun_out/ncu_full_l2f.log 2>&1; echo "ncu full l2f exit $?"
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1]); r=d["roofline"]
    print("$2: value %.0f pairs/s ms/step %.1f | knn %.3f ms frac %.3f share %.2f | e2e %.0f | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["share_of_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e: print("$2 parse fail", e)
PYEOF
}
show gpurun_out/bench_default.json "default (sift 100)"
timeout 900 python bench.py --kind superpoint --images 100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_sp100.json 2>gpurun_out/bench_sp100.err; show gpurun_out/bench_sp100.json "superpoint 100"
timeout 900 python bench.py --kind orb --images 500 --steps 2 --warmup 1 --no-cpu-baseline --no-stages > gpurun_out/bench_orb500.json 2>gpurun_out/bench_orb500.err; show gpurun_out/bench_orb500.json "orb 500 (config 3)"
