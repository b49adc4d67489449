#!/bin/bash
# Round-1 evidence after the fp4 Hamming kernel: every GPU test, smoke, both bench arms, ORB config #3, launch list + ncu --set full
# of the fp4 kernel (a steady-state launch), e2e probe.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1]); r=d["roofline"]
    print("$2: value %.0f pairs/s ms/step %.1f | knn %.3f ms frac %.3f share %.2f | e2e %.0f | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["share_of_step"], d["e2e"]["value"] if d.get("e2e") else -1, d["clocks"]))
except Exception as e: print("$2 parse fail", e)
PYEOF
}
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; show gpurun_out/bench_default.json "default (sift 100)"
timeout 900 python bench.py --kind orb --images 500 --steps 2 --warmup 1 --no-cpu-baseline --no-stages > gpurun_out/bench_orb500.json 2>gpurun_out/bench_orb500.err; show gpurun_out/bench_orb500.json "orb 500 (config 3)"
timeout 900 python bench.py --kind orb --images 100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_orb100.json 2>gpurun_out/bench_orb100.err; show gpurun_out/bench_orb100.json "orb 100"
CMD="python bench.py --kind orb --images 46 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_orb_fp4.csv $CMD > gpurun_out/ncu_launches_orb.log 2>&1; echo "ncu launches orb exit $?"
$CMD > gpurun_out/plain_orb46.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:l2_i8x2 -s 6 -c 1 -f -o gpurun_out/prof_fp4 $CMD > gpurun_out/ncu_full_fp4.log 2>&1; echo "ncu full fp4 exit $?"
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "e2e probe exit $?"; tail -12 gpurun_out/e2e_probe.log
